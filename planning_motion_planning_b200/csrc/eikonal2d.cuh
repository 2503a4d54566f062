// eikonal2d.cuh -- 2D Eikonal solve (replaces FastMarching.py:17-29,44-112).
//
// One warp owns one 32-row x TW-column tile at a time: lane r owns row r.
// The tile and its one-cell halo live in shared memory.  Inside the tile the warp
// runs a cell-granular Fast Iterative Method: every lane keeps a TW-bit mask of
// the cells of its row whose inputs changed; each lock-step iteration every lane
// with a non-empty mask relaxes one of them with the reference's two-branch
// upwind update and, on improvement, re-arms the neighbours that can still
// benefit (left/right in its own mask, up/down by warp shuffle).  The warp-wide
// vote "all masks empty" is the tile's convergence test, so the tile leaves at an
// exact fixed point of the update (epsilon = 0).
#pragma once
#include "fm_common.cuh"

namespace fmb {

constexpr int TILE_H = 32;   // rows per tile == lanes per warp

template <typename real>
struct Problem2D {
    const real *cost;
    long long cost_pitch, cost_qstride;
    real *T;
    long long T_pitch, T_qstride;
    int rows, cols, ntx, nty, nq;
    const int *seeds;        // [nq][2] = x,y
    int *tile_state;         // [nq*ntx*nty]
    Queue q;
    int step_cap;            // in-tile iteration cap (DEV_STEPCAP beyond)
    unsigned long long *tile_prio;   // [nq*ntx*nty] ordered bits of the lowest activating value (best-first mode)
    int best_first;          // 1: ring carries query ids, workers claim the lowest-priority queued tile
    int arm_rows;            // bit 0 / bit 1: the first / last tile row holds a halo row written from outside
                             // (domain decomposition): every visit of those tiles re-arms all their cells
    // windowed order (one large map): a popped tile whose level (priority / delta) is more than
    // `win_window` levels above the lowest queued level is put back at the tail instead of being run
    int windowed, win_window;
    int *lev_count;          // [WIN_LEVELS] queued tiles per level
    int *win_hint;           // lowest level that may be non-empty
    int *tile_level;         // [ntiles] level recorded when the tile was queued
    double *win_inv_delta;   // 1 / (T units per level), set by the seed kernel
    unsigned long long *run_prio;   // [ntiles] priority a tile had when its current / last visit started (windowed == 2)
    int pipeline;            // sweep engine, local causal order: pipelined visits (streaming halos, early activation)
    int precheck;            // sweep engine: 1 = a visit opens with one check pass (recognises no-op visits)
    int check_passes;        // sweep engine: Jacobi check passes tried before another round of sweeps (>= 1)
    int win_div;             // levels per tile crossing at the source's cost (1 in the warp engine)
    double *slack;           // local causal order: tolerance of the wait rule in T units, written by the seed kernel
    double slack_frac;       //   = slack_frac x (tile width x cost at the seed of query 0)
    int variant;             // sweep engine: bit 0 = straight-line sweep step (predication instead of vote + branches)
    double hop_frac;         // second-ring wait rule: slack[1] = hop_frac x the same scale (<= 0: rule off)
    int win_running;         // 1: a tile keeps its level count while it RUNS (released when it finishes), so the
                             //    window is measured from the lowest queued-or-running level
    // sweep engine, one map, cost map still being uploaded while the solve runs (fmb_solve2d_h2d_f64): band b of
    // 2^band_shift rows has arrived once band_ready[b] != 0 (written by the copy stream after the band's DMA)
    const int *band_ready = nullptr;
    int band_shift = 0;
};
constexpr int WIN_LEVELS = 8192;
__device__ __forceinline__ int win_level(unsigned long long pbits, double inv_delta) {
    const double v = __longlong_as_double((long long)pbits) * inv_delta;
    return v < (double)(WIN_LEVELS - 1) ? (int)v : WIN_LEVELS - 1;        // +inf / NaN land on the last level
}
template <typename real>
__device__ __forceinline__ void win_count_push(const Problem2D<real> &P, int item) {
    const int L = win_level(*reinterpret_cast<const volatile unsigned long long *>(&P.tile_prio[item]), *P.win_inv_delta);
    *reinterpret_cast<volatile int *>(&P.tile_level[item]) = L;
    atomicAdd(&P.lev_count[L], 1);
    atomicMin(P.win_hint, L);
    __threadfence();          // level and count are visible before the ring slot is
}

// FastMarching.py:17-29 getEikonal, written branch-for-branch on the values
// a = min(left,right), b = min(up,down).  Products and sums are individually
// rounded (no FMA contraction) so a cell relaxed from the same inputs gives the
// same bits as the reference.
template <typename real>
__device__ __forceinline__ real eikonal_update(real a, real b, real c) {
    using N = num<real>;
    real m = a < b ? a : b;          // inputs are never NaN here; cheaper than fmin()
    real d = N::sub(a, b);
    // one-sided when the other side is too far (or +inf): covers the reference's
    // isinf() branches and `cost < |Thor - Tver|`; (inf - inf) = NaN lands here too
    if (!(fabs(d) <= c)) return N::add(m, c);
    real disc = N::sub(N::mul((real)2, N::mul(c, c)), N::mul(d, d));
    return N::mul((real)0.5, N::add(N::add(a, b), N::sqrt(disc)));
}

// The same update without divergent branches (both branches evaluated, one selected): the sweep engine runs it
// inside a one-warp dependent chain where a divergent branch costs more than the spare arithmetic, and ptxas
// can interleave independent evaluations only when there is no branch between them.  Same operations in the
// same order on the selected path, hence the same bits.  The square root sees 1.0 instead of the discriminant
// of the one-sided case (negative / NaN / inf).  Precondition: finite costs lie in [COST_MIN, COST_MAX] (the
// range in which the discriminant stays inside the branch-free square root's domain); the sweep engine checks
// it on every tile it visits and fails the solve loudly otherwise (DEV_COSTRANGE).
template <typename real> struct cost_range;
template <> struct cost_range<double> { static constexpr double lo = 1e-140, hi = 1e140; };
template <> struct cost_range<float> { static constexpr float lo = 1e-15f, hi = 1e15f; };
template <typename real>
__device__ __forceinline__ real eikonal_update_sel(real a, real b, real c) {
    using N = num<real>;
    const real m = a < b ? a : b;
    const real d = N::sub(a, b);
    const bool two_sided = fabs(d) <= c && c < N::inf();
    const real one = N::add(m, c);
    const real disc = N::sub(N::mul((real)2, N::mul(c, c)), N::mul(d, d));
    const real two = N::mul((real)0.5, N::add(N::add(a, b), sqrt_rn_fast(two_sided ? disc : (real)1)));
    return two_sided ? two : one;
}

template <typename real, int TW>
struct Tile2D {
    static constexpr int PT = TW + 2;                         // smem row pitch (even, PT-1 odd: conflict-free skews)
    // row j (-1..32) starts at (j+1)*PT: [+1] left halo, [+2 .. +TW+1] interior (16-byte aligned
    // for cp.async), [+TW+2] right halo (= slot 0 of the next row, which is otherwise unused)
    static constexpr int T_ELEMS = (TILE_H + 2) * PT + 2;
    static constexpr int C_ELEMS = TILE_H * PT;
    static constexpr int WARP_ELEMS = T_ELEMS + C_ELEMS;
    static constexpr size_t WARP_BYTES = sizeof(real) * WARP_ELEMS;
};

// ---------------------------------------------------------------------------
// init: T = +inf, tile states idle, ring empty, counters zero
template <typename real>
__global__ void init_fill2d_kernel(Problem2D<real> P, int ring_slots) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const real INF = num<real>::inf();
    const long long per_q = (long long)P.rows * P.cols;
    const long long total = per_q * P.nq;
    for (long long i = tid; i < total; i += nth) {
        long long q = i / per_q, r = i - q * per_q;
        long long y = r / P.cols, x = r - y * P.cols;
        P.T[q * P.T_qstride + y * P.T_pitch + x] = INF;
    }
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    for (long long i = tid; i < ntiles; i += nth) {
        P.tile_state[i] = ST_IDLE;
        if (P.best_first || P.windowed) P.tile_prio[i] = 0x7ff0000000000000ULL;
        if (P.windowed == 2) P.run_prio[i] = 0x7ff0000000000000ULL;
    }
    for (long long i = tid; i < ring_slots; i += nth) P.q.ring[i] = -1;
    if (P.windowed == 1)
        for (long long i = tid; i < WIN_LEVELS; i += nth) P.lev_count[i] = 0;
    if (tid == 0) {
        ctl_reset(P.q.ctl);
        if (P.windowed == 1) *P.win_hint = WIN_LEVELS - 1;
    }
}

// seeds: T[seed] = 0 and the tiles that see the seed (its own tile, plus the
// neighbour tile(s) when the seed sits on a tile edge) are queued.
template <typename real, int TW>
__global__ void init_seed2d_kernel(Problem2D<real> P) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nq) return;
    const int sx = P.seeds[2 * q], sy = P.seeds[2 * q + 1];
    if (sx < 0 || sy < 0 || sx >= P.cols || sy >= P.rows) return;   // host validates; nothing to solve
    P.T[q * P.T_qstride + (long long)sy * P.T_pitch + sx] = (real)0;
    const int tx = sx / TW, ty = sy / TILE_H;
    const int base = q * P.ntx * P.nty;
    if (P.windowed == 1) {        // one level = the time to cross one tile at the source's cost
        const real c0 = P.cost[q * P.cost_qstride + (long long)sy * P.cost_pitch + sx];
        const double div = P.win_div > 0 ? (double)P.win_div : 1.0;
        *P.win_inv_delta = (c0 > (real)0 && c0 < num<real>::inf()) ? div / ((double)TW * (double)c0) : div / (double)TW;
    }
    if (P.windowed == 2 && q == 0) {
        const real c0 = P.cost[(long long)sy * P.cost_pitch + sx];
        const double scale = (c0 > (real)0 && c0 < num<real>::inf()) ? (double)TW * (double)c0 : 0.0;
        P.slack[0] = P.slack_frac * scale;
        P.slack[1] = P.hop_frac * scale;
    }
    int cand[5][2] = {{tx, ty}, {-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
    if (sx % TW == 0 && tx > 0) { cand[1][0] = tx - 1; cand[1][1] = ty; }
    if (sx % TW == TW - 1 && tx < P.ntx - 1) { cand[2][0] = tx + 1; cand[2][1] = ty; }
    if (sy % TILE_H == 0 && ty > 0) { cand[3][0] = tx; cand[3][1] = ty - 1; }
    if (sy % TILE_H == TILE_H - 1 && ty < P.nty - 1) { cand[4][0] = tx; cand[4][1] = ty + 1; }
    for (int k = 0; k < 5; ++k) {
        if (cand[k][0] < 0) continue;
        int item = base + cand[k][1] * P.ntx + cand[k][0];
        if (tile_activate(P.tile_state, P.q.ctl, item)) {
            if (P.best_first || P.windowed) P.tile_prio[item] = 0ULL;
            if (P.windowed == 1) win_count_push<real>(P, item);
            q_push(P.q, P.best_first ? q : item);
            atomicAdd(&P.q.ctl->pushes, 1ULL);
        }
    }
}

// resume: keep T as it is (another solve, or halo rows received from a neighbour slab, already
// live in it), reset the scheduler state and queue the tiles selected by `activate`
// (bit 0: first tile row, bit 1: last tile row, bit 2: every tile) plus the seed's tiles.
template <typename real>
__global__ void init_resume2d_kernel(Problem2D<real> P, int ring_slots) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    for (long long i = tid; i < ntiles; i += nth) { P.tile_state[i] = ST_IDLE; P.tile_prio[i] = 0x7ff0000000000000ULL; }
    for (long long i = tid; i < ring_slots; i += nth) P.q.ring[i] = -1;
    if (tid == 0) ctl_reset(P.q.ctl);
}
template <typename real>
__global__ void activate_rows2d_kernel(Problem2D<real> P, int activate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.ntx * P.nty) return;
    const int ty = t / P.ntx;
    // the last array row may sit alone in the last tile row: then the tile row above it (which sees
    // that row through its shared-memory halo) is the one that has work
    const bool last_alone = P.nty >= 2 && (P.rows - 1) % TILE_H == 0;
    const bool on = ((activate & 1) && ty == 0) || (activate & 4) ||
                    ((activate & 2) && (ty == P.nty - 1 || (last_alone && ty == P.nty - 2)));
    if (on && tile_activate(P.tile_state, P.q.ctl, t)) { q_push(P.q, t); atomicAdd(&P.q.ctl->pushes, 1ULL); }
}

// ---------------------------------------------------------------------------
// persistent solver
//
// BEST = false: tiles are popped in FIFO order (best for ONE large map: speculative early
//               visits hide the serial chain of tile visits, tools/sched_model.c).
// BEST = true : the ring carries query ids; a worker that holds a ticket for query q claims
//               the QUEUED tile of q with the lowest priority (smallest value that activated
//               it).  Per-query best-first order cuts re-visits several-fold when many
//               independent queries share the GPU (batched planning).
#ifndef FMB_CG_MINB
#define FMB_CG_MINB 5
#endif
// CG = true : the cost tile is NOT staged in shared memory; the relaxation reads the cost of its cell from global memory
//              (L1-cached, the tile is touched once while its finite-cost masks are built).  Halves the shared memory
//              of a tile in flight, i.e. more resident warps for the latency-bound batch regime.
template <typename real, int TW, int WARPS, bool BEST, bool CG = false>
__global__ void __launch_bounds__(WARPS * 32, CG ? FMB_CG_MINB : 0) solve2d_kernel(Problem2D<real> P) {
    using TL = Tile2D<real, TW>;
    constexpr int PT = TL::PT;
    constexpr unsigned ROWMASK = (TW == 32) ? 0xffffffffu : ((1u << TW) - 1u);
    constexpr int ROWS_PER_IT = 32 / TW;                  // tile rows covered by one 32-lane load
    constexpr int NIT_T = (TILE_H + 2) * TW / 32;         // 32-lane loads for T rows -1..32
    constexpr int NIT_C = TILE_H * TW / 32;
    FMB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    real *sT = reinterpret_cast<real *>(smem_raw) + (size_t)warp * (CG ? TL::T_ELEMS : TL::WARP_ELEMS);
    real *sC = sT + TL::T_ELEMS;            // (unused when CG)
    const real INF = num<real>::inf();
    const int tiles_per_q = P.ntx * P.nty;
    const unsigned long long PRIO_INF = 0x7ff0000000000000ULL;

    unsigned long long n_visits = 0, n_steps = 0, n_evals = 0, n_pushes = 0, n_written = 0, n_defer = 0;
    int streak = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0;

    for (;;) {
        const long long tc0 = clock64();
        int item;
        for (;;) {
            int it = -1;
            if (lane == 0) it = q_pop_lane0(P.q);
            item = __shfl_sync(FULL, it, 0);
            if (item < 0 || BEST || !P.windowed) break;
            // windowed order: run the tile only if it is within win_window levels of the lowest queued
            // level; otherwise put it back at the tail (the lowest queued level itself always runs)
            int defer = 0;
            if (lane == 0 && streak < 32) {            // a worker never defers more than 32 times in a row
                __threadfence();
                const int L = ld_volatile(&P.tile_level[item]);
                atomicSub(&P.lev_count[L], 1);
                const int h0 = ld_volatile(P.win_hint);
                int h = h0;
                const int stop_at = min(L - P.win_window, h0 + 64);          // bounded scan
                while (h < stop_at && ld_volatile(&P.lev_count[h]) <= 0) ++h;
                if (h != h0) atomicCAS(P.win_hint, h0, h);
                if (h < L - P.win_window && h < stop_at) {                    // a queued tile sits more than a window below this one
                    atomicAdd(&P.lev_count[L], 1);
                    q_push(P.q, item);
                    defer = 1;
                }
            }
            else if (lane == 0) atomicSub(&P.lev_count[ld_volatile(&P.tile_level[item])], 1);
            defer = __shfl_sync(FULL, defer, 0);
            if (!defer) { streak = 0; break; }
            ++streak;
            n_defer += 1;
            __nanosleep(streak < 4 ? 200u : streak < 12 ? 800u : 2000u);      // back off: the front is elsewhere
        }
        if (item < 0) break;
        if (BEST) {
            // `item` is a query id: claim its lowest-priority queued tile (QUEUED -> RUNNING)
            const int base = item * tiles_per_q;
            const long long t0 = clock64();
            int claimed = -1;
            while (claimed < 0) {
                unsigned long long bp = ~0ULL;
                int bt = -1;
                // four tiles per lane per batch, all eight loads in flight before any is used
                for (int tb = 0; tb < tiles_per_q; tb += 128) {
                    int stv[4];
                    unsigned long long prv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = tb + u * 32 + lane;
                        const bool in = t < tiles_per_q;
                        stv[u] = in ? ld_volatile(&P.tile_state[base + t]) : ST_IDLE;
                        prv[u] = in ? *reinterpret_cast<const volatile unsigned long long *>(&P.tile_prio[base + t]) : ~0ULL;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (stv[u] == ST_QUEUED && prv[u] < bp) { bp = prv[u]; bt = tb + u * 32 + lane; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long op = __shfl_xor_sync(FULL, bp, o);
                    const int ot = __shfl_xor_sync(FULL, bt, o);
                    if (op < bp || (op == bp && ot >= 0 && (bt < 0 || ot < bt))) { bp = op; bt = ot; }
                }
                int ok = 0;
                if (lane == 0 && bt >= 0) ok = atomicCAS(&P.tile_state[base + bt], ST_QUEUED, ST_RUNNING) == ST_QUEUED;
                ok = __shfl_sync(FULL, ok, 0);
                if (ok) { claimed = base + bt; break; }
                int bail = 0;
                if (lane == 0) {
                    if (ld_volatile(&P.q.ctl->abort)) bail = 1;
                    else if (clock64() - t0 > P.q.watchdog_cycles) { atomicCAS(&P.q.ctl->abort, 0, DEV_WATCHDOG); bail = 1; }
                }
                if (__shfl_sync(FULL, bail, 0)) break;
            }
            if (claimed < 0) break;
            item = claimed;
            if (lane == 0) { atomicExch(&P.tile_prio[item], PRIO_INF); __threadfence(); }
        } else {
            // QUEUED -> RUNNING *before* sampling T: anything published after this point flips
            // the state to DIRTY and the tile is run again.
            if (lane == 0) {
                atomicExch(&P.tile_state[item], ST_RUNNING);
                if (P.windowed) atomicExch(&P.tile_prio[item], PRIO_INF);
                __threadfence();
            }
        }
        __syncwarp();
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        const int t = item - q * tiles_per_q;
        const int ty = t / P.ntx, tx = t - ty * P.ntx;
        const int x0 = tx * TW, y0 = ty * TILE_H;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- stage tile + halo ------------------------------------------------------------
        // interior tiles with 16-byte aligned rows: every 16-byte chunk of the 34 T rows and the 32
        // cost rows is put in flight with cp.async.cg (L2 only) and waited for once; edge / unaligned
        // tiles take the register path (loads issued in batches before their smem stores).
        unsigned cmask = 0;                              // cells that can ever be relaxed (finite cost)
        {
            const int y = y0 + lane;                     // left / right halo columns (lane == row)
            real vl = INF, vr = INF;
            if (y < P.rows) {
                if (x0 > 0) vl = ld_T(&Tq[(long long)y * P.T_pitch + x0 - 1]);
                if (x0 + TW < P.cols) vr = ld_T(&Tq[(long long)y * P.T_pitch + x0 + TW]);
            }
            constexpr int EPC = 16 / (int)sizeof(real);                 // elements per 16-byte chunk
            constexpr int CPR = TW / EPC;                               // chunks per tile row
            // (fp32 rows are only 8-byte aligned in this smem layout: they take the register path)
            const bool fast = sizeof(real) == 8 && y0 >= 1 && y0 + TILE_H < P.rows && x0 + TW <= P.cols &&
                              (P.T_pitch % EPC) == 0 && (P.cost_pitch % EPC) == 0 &&
                              ((size_t)Tq % 16) == 0 && ((size_t)cq % 16) == 0;
            if (fast) {
#pragma unroll
                for (int c = lane; c < (TILE_H + 2) * CPR; c += 32) {
                    const int row = c / CPR, col = (c % CPR) * EPC;
                    cp_async16_cg(&sT[row * PT + 2 + col], &Tq[(long long)(y0 - 1 + row) * P.T_pitch + x0 + col]);
                }
                if (!CG) {
#pragma unroll
                    for (int c = lane; c < TILE_H * CPR; c += 32) {
                        const int row = c / CPR, col = (c % CPR) * EPC;
                        cp_async16_cg(&sC[row * PT + col], &cq[(long long)(y0 + row) * P.cost_pitch + x0 + col]);
                    }
                    cp_async_wait_all();
                    __syncwarp();
#pragma unroll 4
                    for (int j = 0; j < TILE_H; ++j) {
                        const unsigned bal = __ballot_sync(FULL, lane < TW && sC[j * PT + (lane < TW ? lane : 0)] < INF);
                        if (lane == j) cmask = bal;
                    }
                } else {
                    // one coalesced row per load, eight rows in flight; the lines stay in L1 for the relaxation
#pragma unroll
                    for (int j0 = 0; j0 < TILE_H; j0 += 8) {
                        real cv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            cv[u] = lane < TW ? __ldg(&cq[(long long)(y0 + j0 + u) * P.cost_pitch + x0 + lane]) : INF;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const unsigned bal = __ballot_sync(FULL, cv[u] < INF);
                            if (lane == j0 + u) cmask = bal;
                        }
                    }
                    cp_async_wait_all();
                    __syncwarp();
                }
            } else {
                constexpr int BT = 9;
#pragma unroll
                for (int b0 = 0; b0 < NIT_T; b0 += BT) {
                    real v[BT];
#pragma unroll
                    for (int u = 0; u < BT; ++u) {
                        const int idx = (b0 + u) * 32 + lane;
                        const int j = idx / TW - 1, i = idx % TW;
                        const int yy = y0 + j, xx = x0 + i;
                        v[u] = INF;
                        if (b0 + u < NIT_T && yy >= 0 && yy < P.rows && xx < P.cols) v[u] = ld_T(&Tq[(long long)yy * P.T_pitch + xx]);
                    }
#pragma unroll
                    for (int u = 0; u < BT; ++u) {
                        const int idx = (b0 + u) * 32 + lane;
                        if (b0 + u < NIT_T) sT[(idx / TW) * PT + idx % TW + 2] = v[u];
                    }
                }
                constexpr int BC = 8;
#pragma unroll
                for (int b0 = 0; b0 < NIT_C; b0 += BC) {
                    real c[BC];
#pragma unroll
                    for (int u = 0; u < BC; ++u) {
                        const int idx = (b0 + u) * 32 + lane;
                        const int j = idx / TW, i = idx % TW;
                        const int yy = y0 + j, xx = x0 + i;
                        c[u] = INF;
                        if (yy < P.rows && xx < P.cols) c[u] = __ldg(&cq[(long long)yy * P.cost_pitch + xx]);
                    }
#pragma unroll
                    for (int u = 0; u < BC; ++u) {
                        const int idx = (b0 + u) * 32 + lane;
                        if (!CG) sC[(idx / TW) * PT + idx % TW] = c[u];
                        const unsigned bal = __ballot_sync(FULL, c[u] < INF);
#pragma unroll
                        for (int sft = 0; sft < ROWS_PER_IT; ++sft)
                            if (lane == (b0 + u) * ROWS_PER_IT + sft) cmask = (bal >> (sft * TW)) & ROWMASK;
                    }
                }
            }
            sT[(lane + 1) * PT + 1] = vl;
            sT[(lane + 1) * PT + TW + 2] = vr;
        }
        __syncwarp();

        // ---- arm the cells next to a lower halo value -------------------------
        real *rowT = sT + (lane + 1) * PT + 2;      // rowT[k] = T(row lane, col k); rowT[-1], rowT[TW] halos
        const real *rowC = CG ? cq + (long long)(y0 + lane) * P.cost_pitch + x0 : sC + lane * PT;   // only cells of cmask are read
        unsigned mask = 0;
        if (rowT[-1] < rowT[0]) mask |= 1u;
        if (rowT[TW] < rowT[TW - 1]) mask |= 1u << (TW - 1);
        {
            const bool in = lane < TW;
            const unsigned bt = __ballot_sync(FULL, in && sT[lane + 2] < sT[PT + lane + 2]);
            const unsigned bb = __ballot_sync(FULL, in && sT[(TILE_H + 1) * PT + lane + 2] < sT[TILE_H * PT + lane + 2]);
            if (lane == 0) mask |= bt;
            if (lane == TILE_H - 1) mask |= bb;
        }
        {   // a source inside this tile arms its four neighbours
            const int lx = P.seeds[2 * q] - x0, ly = P.seeds[2 * q + 1] - y0;
            if (lx >= 0 && lx < TW && ly >= 0 && ly < TILE_H) {
                if (lane == ly) {
                    if (lx > 0) mask |= 1u << (lx - 1);
                    if (lx < TW - 1) mask |= 1u << (lx + 1);
                }
                if (lane == ly - 1 || lane == ly + 1) mask |= 1u << lx;
            }
        }
        if (((P.arm_rows & 1) && ty == 0) || ((P.arm_rows & 2) && ty == P.nty - 1)) mask = cmask;
        mask &= cmask;

        // ---- relax to the fixed point ---------------------------------------
        // One armed cell per lane per lock-step iteration.  Lanes that evaluate run converged, so
        // every lane's shared-memory reads of an iteration precede that iteration's writes (a
        // Jacobi step) and a single __syncwarp per iteration publishes the writes.
        const long long tc2 = clock64();
        unsigned dirty = 0, last = 0;
        bool up = true;
        int steps = 0;
        real vmin = INF;
        unsigned active;
        while ((active = __ballot_sync(FULL, mask != 0)) != 0) {
            unsigned up_msg = 0, dn_msg = 0;
            if (mask) {
                // keep sweeping in the current direction along the row, turn round at its end
                const unsigned hi = mask & (~0u << last);
                const unsigned lo = mask & ((2u << last) - 1u);
                up = up ? (hi != 0) : (lo == 0);
                const unsigned k = up ? (unsigned)(__ffs(hi) - 1) : (unsigned)(31 - __clz(lo));
                last = k;
                const unsigned bit = 1u << k;
                mask &= ~bit;
                real *p = rowT + k;
                const real ck = CG ? __ldg(rowC + k) : rowC[k];
                const real l = p[-1], r = p[1], u = p[-PT], d = p[PT], cur = p[0];
                const real v = eikonal_update<real>(l < r ? l : r, u < d ? u : d, ck);
                // Lower values always win.  A value up to a few ulp HIGHER also replaces the stored one: a cell
                // keeps the minimum over its history of updates, rounding is not monotone, and without this a
                // cell can stay an ulp below update(final neighbours) -- the reference's field is an exact fixed
                // point of the update, and exact ties between mirror-image cells depend on it.  Upwind
                // dependencies are acyclic, so once a cell's inputs are final it is written once and rests.
                if (v != cur && v <= num<real>::mul(cur, (real)(1.0 + 8.0 / 4503599627370496.0))) {      // lower, or at most ~4 ulp higher
                    *p = v;
                    dirty |= bit;
                    mask |= (l > v ? bit >> 1 : 0u) | (r > v ? bit << 1 : 0u);   // only neighbours that can still improve
                    up_msg = u > v ? bit : 0u;
                    dn_msg = d > v ? bit : 0u;
                    if (BEST || P.windowed) vmin = v < vmin ? v : vmin;
                }
            }
            unsigned from_below = __shfl_down_sync(FULL, up_msg, 1);   // row+1 improved and my cell above it is larger
            unsigned from_above = __shfl_up_sync(FULL, dn_msg, 1);
            if (lane == TILE_H - 1) from_below = 0;
            if (lane == 0) from_above = 0;
            mask = (mask | from_below | from_above) & cmask;
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

        // ---- write back changed cells (dirty rows only, coalesced) ---------------
        {
            unsigned rows_dirty = __ballot_sync(FULL, dirty != 0);
            while (rows_dirty) {
                const int j = __ffs(rows_dirty) - 1;
                rows_dirty &= rows_dirty - 1;
                const unsigned dj = __shfl_sync(FULL, dirty, j);
                if (lane < TW && ((dj >> lane) & 1u))
                    st_T(&Tq[(long long)(y0 + j) * P.T_pitch + x0 + lane], sT[(j + 1) * PT + lane + 2]);
                n_written += __popc(dj);
            }
        }
        // ---- which neighbours can still improve? ----------------------------
        // only an edge cell that changed in this visit AND undercuts the value across the edge
        // can lower anything in the neighbour (causality).
        const bool nl = (dirty & 1u) && rowT[0] < rowT[-1];
        const bool nr = ((dirty >> (TW - 1)) & 1u) && rowT[TW - 1] < rowT[TW];
        const unsigned d_top = __shfl_sync(FULL, dirty, 0), d_bot = __shfl_sync(FULL, dirty, TILE_H - 1);
        const bool nt = lane < TW && ((d_top >> lane) & 1u) && sT[PT + lane + 2] < sT[lane + 2];
        const bool nb = lane < TW && ((d_bot >> lane) & 1u) && sT[TILE_H * PT + lane + 2] < sT[(TILE_H + 1) * PT + lane + 2];
        unsigned act = 0;
        if (__any_sync(FULL, nl) && tx > 0) act |= 1u;
        if (__any_sync(FULL, nr) && tx < P.ntx - 1) act |= 2u;
        if (__any_sync(FULL, nt) && ty > 0) act |= 4u;
        if (__any_sync(FULL, nb) && ty < P.nty - 1) act |= 8u;
        unsigned long long pbits = PRIO_INF;
        if (BEST || P.windowed) {     // priority handed to the neighbours: the lowest value this visit produced
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const real ov = __shfl_xor_sync(FULL, vmin, o); vmin = ov < vmin ? ov : vmin; }
            pbits = (unsigned long long)__double_as_longlong((double)vmin);
        }
        // Publishing.  `pending` is raised for every candidate first (fire-and-forget, ordered by the
        // same fence as the T stores) and corrected afterwards, so it never under-counts.  Lanes 0..3
        // then try one neighbour each and lane 4 retires this tile IN THE SAME compare-and-swap
        // instruction, so the common case costs one atomic round trip instead of five.
        const int nact = __popc(act);
        if (lane == 0 && nact) atomicAdd(&P.q.ctl->pending, nact);
        if ((BEST || P.windowed) && lane < 4 && ((act >> lane) & 1u))
            atomicMin(&P.tile_prio[item + (lane == 0 ? -1 : lane == 1 ? 1 : lane == 2 ? -P.ntx : P.ntx)], pbits);
        __threadfence();          // T stores (+ pending, priorities) are device-visible ...
        __syncwarp();             // ... before any state transition is published
        bool pushed = false, newly = false, requeue = false;
        {
            const bool is_nbr = lane < 4 && ((act >> lane) & 1u);
            const bool is_self = lane == 4;
            const int tgt = is_self ? item : item + (lane == 0 ? -1 : lane == 1 ? 1 : lane == 2 ? -P.ntx : P.ntx);
            if (is_nbr || is_self) {
                int *st = &P.tile_state[tgt];
                int old = atomicCAS(st, is_self ? ST_RUNNING : ST_IDLE, is_self ? ST_IDLE : ST_QUEUED);
                if (is_self) {
                    if (old != ST_RUNNING) { atomicExch(st, ST_QUEUED); requeue = true; }     // was DIRTY: run again
                } else {
                    for (;;) {
                        if (old == ST_IDLE) { newly = true; break; }
                        if (old == ST_QUEUED || old == ST_DIRTY) break;
                        if (atomicCAS(st, ST_RUNNING, ST_DIRTY) == ST_RUNNING) break;          // ask the runner to go again
                        old = atomicCAS(st, ST_IDLE, ST_QUEUED);
                    }
                }
                if (newly || requeue) {
                    if (!BEST && P.windowed == 1) win_count_push<real>(P, tgt);
                    q_push(P.q, BEST ? q : tgt);
                    pushed = true;
                }
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
    // per-warp counters (warp-uniform values; lane 0 reports)
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
        if (n_defer) atomicAdd(&P.q.ctl->pad[0], n_defer);
    }
}

}  // namespace fmb
