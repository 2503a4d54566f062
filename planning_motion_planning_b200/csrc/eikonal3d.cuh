// eikonal3d.cuh -- 3D Eikonal solve (replaces FastMarching3D.py:19-101,126-145).
//
// Volumes are [ny][nx][nz] with z contiguous (the planner's layout,
// Coupled_motion_planner.py:1627).  A tile is 4 (y) x 8 (x) x TZ (z) cells and
// belongs to one warp: lane (ly*8 + lx) owns the z-column (ly, lx) and keeps a
// TZ-bit mask of its armed cells, so z-rows are 128/256-byte coalesced in HBM and
// the in-tile scheme is the 2D one with two more neighbour directions (x and y
// neighbours are other lanes, reached by shuffle; z neighbours are mask bits).
#pragma once
#include "fm_common.cuh"
#include "pow2_glibc.cuh"

namespace fmb {

constexpr int T3Y = 4, T3X = 8;     // tile cross-section: 32 columns == 32 lanes

template <typename real>
struct Problem3D {
    const real *cost;
    long long cost_qstride;
    real *T;
    long long T_qstride;
    int ny, nx, nz, nty, ntx, ntz, nq;
    const int *seeds;        // [nq][3] = x,y,z
    int *tile_state;
    Queue q;
    int step_cap;
    int arm_all;             // 1: every visit re-arms all cells of its tile (polish pass over an existing field)
    // sweep engine (eikonal3d_sweep.cuh)
    unsigned long long *tile_prio;   // [ntiles] ordered bits of the lowest value that activated the tile since its last visit began
    unsigned long long *run_prio;    // [ntiles] priority the tile had when its current / last visit started
    int causal;              // 1: local causal order (a tile waits for queued / running neighbours of lower priority)
    int check_passes;        // Jacobi check passes tried before another round of sweeps (>= 1)
    double *slack;           // tolerance of the causal wait rule in T units, written by the seed kernel:
    double slack_frac;       //   slack_frac x (T3Y cells x cost at the seed of query 0)
    double hop_frac;         // second-ring wait rule: slack[1] = hop_frac x the same scale (<= 0: off)
    int variant;             // bit 0: straight-line sweep step
    const int *enable;       // device flag or nullptr: when it reads 0 the init and solve kernels of this launch return at once
                             // (a solve that a preceding kernel found unnecessary: fmb_solve3d_until_f64)
};

// FastMarching3D.py:59-75 -- descending-dimension quadratic solver, in the
// reference's operation order (Tarray = [Tx,Ty,Tz]; sumlist() is right-associated).
template <typename real>
__device__ __forceinline__ real solve3d_update(real t0, real t1, real t2, real C) {
    using N = num<real>;
    const real C2 = N::mul(C, C);
    // ---- n = 3
    {
        int im = 0; real mx = t0;
        if (t1 > mx) { mx = t1; im = 1; }
        if (t2 > mx) { mx = t2; im = 2; }
        real d0 = N::sub(mx, t0), d1 = N::sub(mx, t1), d2 = N::sub(mx, t2);
        real sumT = N::add(N::add(N::mul(d0, d0), N::mul(d1, d1)), N::mul(d2, d2));
        if (C2 > sumT) {
            real S = N::add(t0, N::add(t1, t2));
            real Q = N::add(N::mul(t0, t0), N::add(N::mul(t1, t1), N::mul(t2, t2)));
            real disc = N::sub(N::add(N::mul((real)3, C2), N::mul(S, S)), N::mul((real)3, Q));
            return N::div3(N::add(S, N::sqrt(disc)));
        }
        // Tarray.remove(Tmax): keep the other two in order
        if (im == 0) { t0 = t1; t1 = t2; } else if (im == 1) { t1 = t2; }
    }
    // ---- n = 2
    {
        const bool first_is_max = !(t1 > t0);
        const real mx = first_is_max ? t0 : t1;
        real d0 = N::sub(mx, t0), d1 = N::sub(mx, t1);
        real sumT = N::add(N::mul(d0, d0), N::mul(d1, d1));
        if (C2 > sumT) {
            real S = N::add(t0, t1);
            real Q = N::add(N::mul(t0, t0), N::mul(t1, t1));
            real disc = N::sub(N::add(N::mul((real)2, C2), N::mul(S, S)), N::mul((real)2, Q));
            return N::mul(N::add(S, N::sqrt(disc)), (real)0.5);
        }
        if (first_is_max) t0 = t1;
    }
    // ---- n = 1
    {
        real d0 = N::sub(t0, t0);                     // NaN when t0 is +inf: the test below fails like the reference's
        if (C2 > N::mul(d0, d0)) {
            real disc = N::sub(N::add(C2, N::mul(t0, t0)), N::mul(t0, t0));
            return N::add(t0, N::sqrt(disc));
        }
    }
    return num<real>::inf();                          // no finite neighbour (reference: max([]) raises)
}

// The same solver with the reference's OWN rounding of the squares it takes on NumPy scalars
// (FastMarching3D.py:68-71: (Tmax-Tarray[a])**2, C**2 and (sumlist(Tarray))**2 go through libm pow, which is not
// correctly rounded; array(Tarray)**2 is an exact-rounded product): bit-for-bit the value the reference assigns.
// ~5x the arithmetic of solve3d_update, so it runs as a polish pass over the converged field (fmb_polish3d_f64):
// the two fixed points differ by an ulp in a fraction of a percent of the cells -- enough to decide exact ties of
// the pop order, and with them the accepted set of the early exit on uniform-cost volumes.
__device__ __forceinline__ double solve3d_update_exact(double t0, double t1, double t2, double C) {
    using N = num<double>;
    const double C2 = pow2_glibc(C);
    double a[3] = {t0, t1, t2};
#pragma unroll
    for (int n = 3; n >= 1; --n) {
        int im = 0;
#pragma unroll
        for (int i = 1; i < 3; ++i) if (i < n && a[i] > a[im]) im = i;          // max(): first maximum
        const double mx = a[im];
        double sumT = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) if (i < n) sumT = N::add(sumT, pow2_glibc(N::sub(mx, a[i])));
        if (C2 > sumT) {
            double S, Q;
            if (n == 3) { S = N::add(a[0], N::add(a[1], a[2])); Q = N::add(N::mul(a[0], a[0]), N::add(N::mul(a[1], a[1]), N::mul(a[2], a[2]))); }
            else if (n == 2) { S = N::add(a[0], a[1]); Q = N::add(N::mul(a[0], a[0]), N::mul(a[1], a[1])); }
            else { S = a[0]; Q = N::mul(a[0], a[0]); }
            const double disc = N::sub(N::add(N::mul((double)n, C2), pow2_glibc(S)), N::mul((double)n, Q));
            const double num_ = N::add(S, N::sqrt(disc));
            return n == 3 ? N::div3(num_) : n == 2 ? N::mul(num_, 0.5) : num_;
        }
        // Tarray.remove(Tmax)
        if (im == 0) { a[0] = a[1]; a[1] = a[2]; } else if (im == 1) { a[1] = a[2]; }
    }
    return N::inf();                                  // no finite neighbour (reference: max([]) raises)
}

// Branch-free form of the two solvers above (sweep engine): the level n of the descending-dimension scheme is decided
// by its tests alone (subtractions, squares, sums -- no square root), then ONE quadratic is evaluated with operands
// chosen by selects, in the reference's operation order for that level, so the value is bit for bit the one the
// branching form returns.  A warp whose lanes take different levels no longer serialises three square roots
// (measured: ~1150 cycles per sweep step with the branching form).  `slow` is raised where the fast paths do not
// apply (square root outside sqrt_rn_fast's range, quotient outside div3's): the caller re-evaluates those lanes
// with the branching form after a vote.
template <bool EXACT>
__device__ __forceinline__ double solve3d_update_sel(double t0, double t1, double t2, double C, bool &slow) {
    using N = num<double>;
    const double C2 = EXACT ? pow2_glibc(C) : N::mul(C, C);
    // n = 3: max() keeps the first maximum.  (Tmax - Tmax)**2 is an exact zero and adding it is exact, so the sum of the
    // three squares in the reference's order equals the sum of the squares of the two OTHER differences in their
    // order: two squarings instead of three (and one instead of two at n = 2) -- what counts when a squaring is libm pow.
    const bool g1 = t1 > t0;
    const double m01 = g1 ? t1 : t0;
    const bool g2 = t2 > m01;
    const double mx3 = g2 ? t2 : m01;
    // Tarray.remove(Tmax): the other two, in order
    const bool rm0 = !g1 && !g2, rm2 = g2;
    const double a = rm0 ? t1 : t0, b = rm2 ? t1 : t2;
    const double ea = N::sub(mx3, a), eb = N::sub(mx3, b);
    const double pa = EXACT ? pow2_glibc(ea) : N::mul(ea, ea), pb = EXACT ? pow2_glibc(eb) : N::mul(eb, eb);
    // (inf - inf = NaN at the maximum itself when it is +inf: the reference's sum is NaN then and its test fails)
    const bool ok3 = mx3 < N::inf() && C2 > N::add(pa, pb);
    const bool fmax = !(b > a);
    const double mx2 = fmax ? a : b;
    const double c1 = fmax ? b : a;                      // the last one standing
    const double f1 = N::sub(mx2, c1);
    const double r1 = EXACT ? pow2_glibc(f1) : N::mul(f1, f1);
    const bool ok2 = mx2 < N::inf() && C2 > r1;
    const bool ok1 = c1 < N::inf() && C2 > 0.0;          // (c1 - c1)**2 = 0, NaN when c1 is +inf
    const int n = ok3 ? 3 : ok2 ? 2 : ok1 ? 1 : 0;
    const double q0 = N::mul(t0, t0), q1 = N::mul(t1, t1), q2 = N::mul(t2, t2);
    const double qa = rm0 ? q1 : q0, qb = rm2 ? q1 : q2;
    const double S = n == 3 ? N::add(t0, N::add(t1, t2)) : n == 2 ? N::add(a, b) : c1;
    const double Q = n == 3 ? N::add(q0, N::add(q1, q2)) : n == 2 ? N::add(qa, qb) : (fmax ? qb : qa);
    const double nf = (double)n;
    const double disc = N::sub(N::add(N::mul(nf, C2), EXACT ? pow2_glibc(S) : N::mul(S, S)), N::mul(nf, Q));
    const bool rootable = n > 0 && sqrt_fast_ok(disc);
    const double num_ = N::add(S, sqrt_rn_fast(rootable ? disc : 1.0));
    // x / 3 as in num<double>::div3, fast path only
    const double third = __longlong_as_double(0x3FD5555555555555LL);
    const double qq = N::mul(num_, third);
    const double q3 = __fma_rn(__fma_rn(-3.0, qq, num_), third, qq);
    const bool div_ok = fabs(num_) < 1e300 && fabs(num_) > 1e-290;
    slow = n > 0 && (!rootable || (n == 3 && !div_ok));
    const double r = n == 3 ? q3 : n == 2 ? N::mul(num_, 0.5) : num_;
    return n > 0 ? r : N::inf();
}

// fp32 variant: the reference expression n*C^2 + S^2 - n*Q cancels catastrophically in single
// precision once T >> C (hard part 3), so the float kernel solves the same quadratic in
// shifted variables u_i = t_i - min(t): T = m + (Su + sqrt(n*C^2 + Su^2 - n*Qu)) / n.  Same
// descending-dimension logic; only the fp64 path has to reproduce the reference's rounding.
template <>
__device__ __forceinline__ float solve3d_update<float>(float t0, float t1, float t2, float C) {
    const float INF = num<float>::inf();
    float lo = fminf(t0, fminf(t1, t2));
    if (!(lo < INF)) return INF;
    float u[3] = {t0 - lo, t1 - lo, t2 - lo};
    // sort ascending (3 elements); +inf entries end up last and are dropped by the test below
    if (u[0] > u[1]) { float x = u[0]; u[0] = u[1]; u[1] = x; }
    if (u[1] > u[2]) { float x = u[1]; u[1] = u[2]; u[2] = x; }
    if (u[0] > u[1]) { float x = u[0]; u[0] = u[1]; u[1] = x; }
    const float C2 = C * C;
    for (int n = 3; n >= 1; --n) {
        const float mx = u[n - 1];
        float sumT = 0.f, S = 0.f, Q = 0.f;
        for (int i = 0; i < n; ++i) { const float d = mx - u[i]; sumT += d * d; S += u[i]; Q += u[i] * u[i]; }
        if (C2 > sumT) return lo + (S + sqrtf(fmaxf((float)n * C2 + S * S - (float)n * Q, 0.f))) / (float)n;
    }
    return INF;
}

template <typename real, int TZ>
struct Tile3D {
    static constexpr int PZ = TZ + 2;                          // z pitch incl. halo
    static constexpr int PS = (T3X + 2) * PZ;                  // y-slab pitch
    // column (y,x) starts at (y+1)*PS + (x+1)*PZ: [+1] low z halo, [+2 .. +TZ+1] interior (16-byte
    // aligned for cp.async), [+TZ+2] high z halo (= slot 0 of the next column, otherwise unused)
    static constexpr int T_ELEMS = (T3Y + 2) * PS + 2;
    static constexpr int C_ELEMS = T3Y * T3X * PZ;
    static constexpr int WARP_ELEMS = T_ELEMS + C_ELEMS;
    static constexpr size_t WARP_BYTES = sizeof(real) * WARP_ELEMS;
    // element index of cell (y,x,z) with y in [-1,T3Y], x in [-1,T3X], z in [-1,TZ]
    static __device__ __forceinline__ int at(int y, int x, int z) { return (y + 1) * PS + (x + 1) * PZ + z + 2; }
};

template <typename real>
__global__ void init_fill3d_kernel(Problem3D<real> P, int ring_slots) {
    if (P.enable && !*P.enable) return;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const real INF = num<real>::inf();
    const long long per_q = (long long)P.ny * P.nx * P.nz;
    for (int q = 0; q < P.nq; ++q) {
        real *Tq = P.T + (long long)q * P.T_qstride;
        for (long long i = tid; i < per_q; i += nth) Tq[i] = INF;
    }
    const long long ntiles = (long long)P.nq * P.nty * P.ntx * P.ntz;
    for (long long i = tid; i < ntiles; i += nth) {
        P.tile_state[i] = ST_IDLE;
        if (P.causal) { P.tile_prio[i] = 0x7ff0000000000000ULL; P.run_prio[i] = 0x7ff0000000000000ULL; }
    }
    for (long long i = tid; i < ring_slots; i += nth) P.q.ring[i] = -1;
    if (tid == 0) ctl_reset(P.q.ctl);
}

template <typename real, int TZ>
__global__ void init_seed3d_kernel(Problem3D<real> P) {
    if (P.enable && !*P.enable) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nq) return;
    const int sx = P.seeds[3 * q], sy = P.seeds[3 * q + 1], sz = P.seeds[3 * q + 2];
    if (sx < 0 || sy < 0 || sz < 0 || sx >= P.nx || sy >= P.ny || sz >= P.nz) return;
    P.T[(long long)q * P.T_qstride + ((long long)sy * P.nx + sx) * P.nz + sz] = (real)0;
    const int tx = sx / T3X, ty = sy / T3Y, tz = sz / TZ;
    const int base = q * P.nty * P.ntx * P.ntz;
    if (P.causal && q == 0) {
        const real c0 = P.cost[((long long)sy * P.nx + sx) * P.nz + sz];
        const double scale = (c0 > (real)0 && c0 < num<real>::inf()) ? (double)T3Y * (double)c0 : 0.0;
        P.slack[0] = P.slack_frac * scale;
        P.slack[1] = P.hop_frac * scale;
    }
    // the seed's own tile and every face-neighbour tile that sees it in its halo
    for (int k = 0; k < 7; ++k) {
        int cx = tx, cy = ty, cz = tz;
        bool ok = true;
        switch (k) {
            case 1: ok = (sx % T3X == 0) && tx > 0; cx = tx - 1; break;
            case 2: ok = (sx % T3X == T3X - 1) && tx < P.ntx - 1; cx = tx + 1; break;
            case 3: ok = (sy % T3Y == 0) && ty > 0; cy = ty - 1; break;
            case 4: ok = (sy % T3Y == T3Y - 1) && ty < P.nty - 1; cy = ty + 1; break;
            case 5: ok = (sz % TZ == 0) && tz > 0; cz = tz - 1; break;
            case 6: ok = (sz % TZ == TZ - 1) && tz < P.ntz - 1; cz = tz + 1; break;
            default: break;
        }
        if (!ok) continue;
        const int item = base + (cy * P.ntx + cx) * P.ntz + cz;
        if (tile_activate(P.tile_state, P.q.ctl, item)) {
            if (P.causal) P.tile_prio[item] = 0ULL;
            q_push(P.q, item);
            atomicAdd(&P.q.ctl->pushes, 1ULL);
        }
    }
}

// resume: keep T, reset the scheduler state; activate_all3d then queues every tile (polish pass)
template <typename real>
__global__ void init_resume3d_kernel(Problem3D<real> P, int ring_slots) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const long long ntiles = (long long)P.nq * P.nty * P.ntx * P.ntz;
    for (long long i = tid; i < ntiles; i += nth) {
        P.tile_state[i] = ST_IDLE;
        if (P.causal) { P.tile_prio[i] = 0x7ff0000000000000ULL; P.run_prio[i] = 0x7ff0000000000000ULL; }
    }
    for (long long i = tid; i < ring_slots; i += nth) P.q.ring[i] = -1;
    if (tid == 0) ctl_reset(P.q.ctl);
}
template <typename real>
__global__ void activate_all3d_kernel(Problem3D<real> P) {
    const long long ntiles = (long long)P.nq * P.nty * P.ntx * P.ntz;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles && tile_activate(P.tile_state, P.q.ctl, (int)t)) q_push(P.q, (int)t);
}

template <typename real, int TZ, int WARPS, bool EXACT = false>
__global__ void __launch_bounds__(WARPS * 32) solve3d_kernel(Problem3D<real> P) {
    if (P.enable && !*P.enable) return;
    using TL = Tile3D<real, TZ>;
    constexpr int PZ = TL::PZ, PS = TL::PS;
    FMB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    real *sT = reinterpret_cast<real *>(smem_raw) + (size_t)warp * TL::WARP_ELEMS;
    real *sC = sT + TL::T_ELEMS;
    const real INF = num<real>::inf();
    const int tiles_per_q = P.nty * P.ntx * P.ntz;
    const int ly = lane >> 3, lx = lane & 7;          // my column
    const long long sy_ = (long long)P.nx * P.nz, sx_ = P.nz;   // global strides

    unsigned long long n_visits = 0, n_steps = 0, n_evals = 0, n_pushes = 0, n_written = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0;

    for (;;) {
        const long long tc0 = clock64();
        int item;
        {
            int it = -1;
            if (lane == 0) it = q_pop_lane0(P.q);
            item = __shfl_sync(FULL, it, 0);
            if (item < 0) break;
        }
        if (lane == 0) { atomicExch(&P.tile_state[item], ST_RUNNING); __threadfence(); }
        __syncwarp();
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        int t = item - q * tiles_per_q;
        const int tz = t % P.ntz; t /= P.ntz;
        const int tx = t % P.ntx; const int ty = t / P.ntx;
        const int x0 = tx * T3X, y0 = ty * T3Y, z0 = tz * TZ;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- stage T (56 z-columns incl. lateral halos) and cost (32 columns) ------------------
        unsigned cmask = 0;
        {
            const int y = y0 + ly, x = x0 + lx;       // z halos of my own column
            real zlo = INF, zhi = INF;
            if (y < P.ny && x < P.nx) {
                if (z0 > 0) zlo = ld_T(&Tq[y * sy_ + x * sx_ + z0 - 1]);
                if (z0 + TZ < P.nz) zhi = ld_T(&Tq[y * sy_ + x * sx_ + z0 + TZ]);
            }
            constexpr int NCOL = (T3Y + 2) * (T3X + 2);
            constexpr int EPC = 16 / (int)sizeof(real);
            constexpr int CPC = TZ / EPC;                              // 16-byte chunks per column
            const bool fast = sizeof(real) == 8 && y0 >= 1 && y0 + T3Y < P.ny && x0 >= 1 && x0 + T3X < P.nx &&
                              z0 + TZ <= P.nz && (P.nz % EPC) == 0 && ((size_t)Tq % 16) == 0 && ((size_t)cq % 16) == 0;
            if (fast) {
                // every 16-byte chunk of every column in flight at once (cp.async.cg: L2 only)
#pragma unroll 4
                for (int c = lane; c < NCOL * CPC; c += 32) {
                    const int col = c / CPC, zz = (c % CPC) * EPC;
                    const int yy = col / (T3X + 2) - 1, xx = col % (T3X + 2) - 1;
                    const bool corner = (yy < 0 || yy >= T3Y) && (xx < 0 || xx >= T3X);
                    if (!corner) cp_async16_cg(&sT[TL::at(yy, xx, zz)], &Tq[(y0 + yy) * sy_ + (x0 + xx) * sx_ + z0 + zz]);
                }
#pragma unroll 4
                for (int c = lane; c < 32 * CPC; c += 32) {
                    const int col = c / CPC, zz = (c % CPC) * EPC;
                    cp_async16_cg(&sC[col * PZ + zz], &cq[(y0 + (col >> 3)) * sy_ + (x0 + (col & 7)) * sx_ + z0 + zz]);
                }
                cp_async_wait_all();
                __syncwarp();
            } else {
                constexpr int BT = 8;
                for (int b0 = 0; b0 < NCOL; b0 += BT) {
                    real v[BT];
#pragma unroll
                    for (int u = 0; u < BT; ++u) {
                        const int col = b0 + u;
                        const int yy = col / (T3X + 2) - 1, xx = col % (T3X + 2) - 1;
                        const int gy = y0 + yy, gx = x0 + xx, gz = z0 + lane;
                        v[u] = INF;
                        if (col < NCOL && lane < TZ && gy >= 0 && gy < P.ny && gx >= 0 && gx < P.nx && gz < P.nz)
                            v[u] = ld_T(&Tq[gy * sy_ + gx * sx_ + gz]);
                    }
#pragma unroll
                    for (int u = 0; u < BT; ++u) {
                        const int col = b0 + u;
                        const int yy = col / (T3X + 2) - 1, xx = col % (T3X + 2) - 1;
                        const bool corner = (yy < 0 || yy >= T3Y) && (xx < 0 || xx >= T3X);
                        if (col < NCOL && lane < TZ && !corner) sT[TL::at(yy, xx, lane)] = v[u];
                    }
                }
                for (int b0 = 0; b0 < 32; b0 += BT) {
                    real v[BT];
#pragma unroll
                    for (int u = 0; u < BT; ++u) {
                        const int col = b0 + u;
                        const int gy = y0 + (col >> 3), gx = x0 + (col & 7), gz = z0 + lane;
                        v[u] = INF;
                        if (lane < TZ && gy < P.ny && gx < P.nx && gz < P.nz) v[u] = __ldg(&cq[gy * sy_ + gx * sx_ + gz]);
                    }
#pragma unroll
                    for (int u = 0; u < BT; ++u) if (lane < TZ) sC[(b0 + u) * PZ + lane] = v[u];
                }
                __syncwarp();
            }
            sT[TL::at(ly, lx, -1)] = zlo;
            sT[TL::at(ly, lx, TZ)] = zhi;
            for (int c = 0; c < 32; ++c) {           // finite-cost mask per column (lane == z for the vote)
                const unsigned b = __ballot_sync(FULL, lane < TZ && sC[c * PZ + (lane < TZ ? lane : 0)] < INF);
                if (lane == c) cmask = b;
            }
        }
        __syncwarp();

        // ---- arm the cells next to a lower halo value -------------------------
        real *col = sT + TL::at(ly, lx, 0);          // col[k] = T(my column, z=k)
        const real *colC = sC + lane * PZ;
        unsigned mask = 0;
        if (col[-1] < col[0]) mask |= 1u;
        if (col[TZ] < col[TZ - 1]) mask |= 1u << (TZ - 1);
        for (int c = 0; c < 32; ++c) {               // lateral faces, lane == z
            const int yy = c >> 3, xx = c & 7;
            const bool fxm = xx == 0, fxp = xx == T3X - 1, fym = yy == 0, fyp = yy == T3Y - 1;
            if (!(fxm || fxp || fym || fyp)) continue;
            const int zz = lane < TZ ? lane : 0;
            const real own = sT[TL::at(yy, xx, zz)];
            bool lower = false;
            if (fxm) lower |= sT[TL::at(yy, xx - 1, zz)] < own;
            if (fxp) lower |= sT[TL::at(yy, xx + 1, zz)] < own;
            if (fym) lower |= sT[TL::at(yy - 1, xx, zz)] < own;
            if (fyp) lower |= sT[TL::at(yy + 1, xx, zz)] < own;
            const unsigned b = __ballot_sync(FULL, lane < TZ && lower);
            if (lane == c) mask |= b;
        }
        {   // a source inside this tile arms its six neighbours
            const int sx = P.seeds[3 * q] - x0, sy = P.seeds[3 * q + 1] - y0, sz = P.seeds[3 * q + 2] - z0;
            if (sx >= 0 && sx < T3X && sy >= 0 && sy < T3Y && sz >= 0 && sz < TZ) {
                if (lx == sx && ly == sy) {
                    if (sz > 0) mask |= 1u << (sz - 1);
                    if (sz < TZ - 1) mask |= 1u << (sz + 1);
                }
                if (ly == sy && (lx == sx - 1 || lx == sx + 1)) mask |= 1u << sz;
                if (lx == sx && (ly == sy - 1 || ly == sy + 1)) mask |= 1u << sz;
            }
        }
        if (P.arm_all) mask = cmask;
        mask &= cmask;

        // ---- relax to the fixed point (same lock-step scheme as 2D; see eikonal2d.cuh) --------
        const long long tc2 = clock64();
        unsigned dirty = 0, last = 0;
        bool up = true;
        int steps = 0;
        unsigned active;
        while ((active = __ballot_sync(FULL, mask != 0)) != 0) {
            unsigned m_xm = 0, m_xp = 0, m_ym = 0, m_yp = 0;
            if (mask) {
                const unsigned hi = mask & (~0u << last);
                const unsigned lo = mask & ((2u << last) - 1u);
                up = up ? (hi != 0) : (lo == 0);
                const unsigned k = up ? (unsigned)(__ffs(hi) - 1) : (unsigned)(31 - __clz(lo));
                last = k;
                const unsigned bit = 1u << k;
                mask &= ~bit;
                real *p = col + k;
                const real zm = p[-1], zp = p[1], xm = p[-PZ], xp = p[PZ], ym = p[-PS], yp = p[PS], cur = p[0];
                // FastMarching3D.py:44-57: per-axis minimum, Tarray = [Tx, Ty, Tz]
                real v;
                if (EXACT) v = (real)solve3d_update_exact((double)(xm < xp ? xm : xp), (double)(ym < yp ? ym : yp), (double)(zm < zp ? zm : zp), (double)colC[k]);
                else v = solve3d_update<real>(xm < xp ? xm : xp, ym < yp ? ym : yp, zm < zp ? zm : zp, colC[k]);
                // lower values always win; a value a few ulp higher also replaces the stored one, so that the field
                // ends as an exact fixed point of the update instead of the minimum over a history of roundings
                // (see eikonal2d.cuh)
                // (polish pass: the reference's 3D expression is ill-conditioned -- a 1-ulp change of an input moves the
                // result by ~200 ulp -- so the exact-arithmetic value may lie well above the fast one: accept within 1e-11)
                if (v != cur && v <= num<real>::mul(cur, EXACT ? (real)(1.0 + 1e-11) : (real)(1.0 + 8.0 / 4503599627370496.0))) {      // lower, or at most ~4 ulp higher
                    *p = v;
                    dirty |= bit;
                    mask |= (zm > v ? bit >> 1 : 0u) | (zp > v ? bit << 1 : 0u);
                    m_xm = xm > v ? bit : 0u;
                    m_xp = xp > v ? bit : 0u;
                    m_ym = ym > v ? bit : 0u;
                    m_yp = yp > v ? bit : 0u;
                }
            }
            unsigned r_xp = __shfl_down_sync(FULL, m_xm, 1);   // my x+ neighbour improved and I am its x-
            unsigned r_xm = __shfl_up_sync(FULL, m_xp, 1);
            unsigned r_yp = __shfl_down_sync(FULL, m_ym, 8);
            unsigned r_ym = __shfl_up_sync(FULL, m_yp, 8);
            if (lx == T3X - 1) r_xp = 0;
            if (lx == 0) r_xm = 0;
            if (ly == T3Y - 1) r_yp = 0;
            if (ly == 0) r_ym = 0;
            mask = (mask | r_xp | r_xm | r_yp | r_ym) & cmask;
            __syncwarp();
            n_evals += __popc(active);
            if (++steps > P.step_cap) break;
        }
        n_steps += steps;
        ++n_visits;
        if (steps > P.step_cap) {
            if (lane == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_STEPCAP);
            break;
        }
        const long long tc3 = clock64();

        // ---- write back dirty columns + face tests (lane == z) --------------------------------
        bool f_xm = false, f_xp = false, f_ym = false, f_yp = false;
        {
            unsigned cols_dirty = __ballot_sync(FULL, dirty != 0);
            while (cols_dirty) {
                const int c = __ffs(cols_dirty) - 1;
                cols_dirty &= cols_dirty - 1;
                const unsigned dj = __shfl_sync(FULL, dirty, c);
                const int yy = c >> 3, xx = c & 7;
                if (lane < TZ && ((dj >> lane) & 1u)) {
                    const real nv = sT[TL::at(yy, xx, lane)];
                    st_T(&Tq[(y0 + yy) * sy_ + (x0 + xx) * sx_ + z0 + lane], nv);
                    if (xx == 0 && nv < sT[TL::at(yy, -1, lane)]) f_xm = true;
                    if (xx == T3X - 1 && nv < sT[TL::at(yy, T3X, lane)]) f_xp = true;
                    if (yy == 0 && nv < sT[TL::at(-1, xx, lane)]) f_ym = true;
                    if (yy == T3Y - 1 && nv < sT[TL::at(T3Y, xx, lane)]) f_yp = true;
                }
                n_written += __popc(dj);
            }
        }
        const bool f_zm = (dirty & 1u) && col[0] < col[-1];
        const bool f_zp = ((dirty >> (TZ - 1)) & 1u) && col[TZ - 1] < col[TZ];
        unsigned act = 0;                     // bit s: neighbour s = z-, z+, x-, x+, y-, y+
        if (__any_sync(FULL, f_zm) && tz > 0) act |= 1u;
        if (__any_sync(FULL, f_zp) && tz < P.ntz - 1) act |= 2u;
        if (__any_sync(FULL, f_xm) && tx > 0) act |= 4u;
        if (__any_sync(FULL, f_xp) && tx < P.ntx - 1) act |= 8u;
        if (__any_sync(FULL, f_ym) && ty > 0) act |= 16u;
        if (__any_sync(FULL, f_yp) && ty < P.nty - 1) act |= 32u;
        const int nact = __popc(act);
        if (lane == 0 && nact) atomicAdd(&P.q.ctl->pending, nact);
        __threadfence();
        __syncwarp();
        bool pushed = false, newly = false, requeue = false;
        {
            const bool is_nbr = lane < 6 && ((act >> lane) & 1u);
            const bool is_self = lane == 6;
            const int off = lane == 0 ? -1 : lane == 1 ? 1 : lane == 2 ? -P.ntz : lane == 3 ? P.ntz
                            : lane == 4 ? -P.ntx * P.ntz : P.ntx * P.ntz;
            const int tgt = is_self ? item : item + off;
            if (is_nbr || is_self) {
                int *st = &P.tile_state[tgt];
                int old = atomicCAS(st, is_self ? ST_RUNNING : ST_IDLE, is_self ? ST_IDLE : ST_QUEUED);
                if (is_self) {
                    if (old != ST_RUNNING) { atomicExch(st, ST_QUEUED); requeue = true; }
                } else {
                    for (;;) {
                        if (old == ST_IDLE) { newly = true; break; }
                        if (old == ST_QUEUED || old == ST_DIRTY) break;
                        if (atomicCAS(st, ST_RUNNING, ST_DIRTY) == ST_RUNNING) break;
                        old = atomicCAS(st, ST_IDLE, ST_QUEUED);
                    }
                }
                if (newly || requeue) { q_push(P.q, tgt); pushed = true; }
            }
        }
        const int n_new = __popc(__ballot_sync(FULL, newly));
        const int n_req = __popc(__ballot_sync(FULL, requeue));
        n_pushes += __popc(__ballot_sync(FULL, pushed));
        int stop = 0;
        if (lane == 0) {
            const int drop = (nact - n_new) + (n_req ? 0 : 1);
            if (drop) atomicSub(&P.q.ctl->pending, drop);
            stop = ld_volatile(&P.q.ctl->abort);
        }
        stop = __shfl_sync(FULL, stop, 0);
        const long long tc4 = clock64();
        c_wait += tc1 - tc0; c_load += tc2 - tc1; c_relax += tc3 - tc2; c_store += tc4 - tc3;
        if (stop) break;
    }
    if (lane == 0) {
        atomicAdd(&P.q.ctl->cyc_wait, (unsigned long long)c_wait);
        atomicAdd(&P.q.ctl->cyc_load, (unsigned long long)c_load);
        atomicAdd(&P.q.ctl->cyc_relax, (unsigned long long)c_relax);
        atomicAdd(&P.q.ctl->cyc_store, (unsigned long long)c_store);
        atomicAdd(&P.q.ctl->visits, n_visits);
        atomicAdd(&P.q.ctl->steps, n_steps);
        atomicAdd(&P.q.ctl->evals, n_evals);
        atomicAdd(&P.q.ctl->pushes, n_pushes);
        atomicAdd(&P.q.ctl->cells_written, n_written);
    }
}

}  // namespace fmb
