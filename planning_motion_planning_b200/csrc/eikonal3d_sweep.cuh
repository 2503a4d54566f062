// eikonal3d_sweep.cuh -- 3D Eikonal solve, sweep engine (replaces FastMarching3D.py:19-101,126-145).
//
// One tile visit = one CTA of 8 warps; each warp runs ONE of the eight plane-wavefront Gauss-Seidel sweeps
// (+-y, +-x, +-z) over the same 4 x 8 x 16 shared tile at the same time -- the 3D form of eikonal2d_sweep.cuh.
// Lane (ly, lx) owns one z-column in the warp's sweep orientation and relaxes cell kz = d - ly - lx at step d:
// its three sweep-upwind neighbours were relaxed one step earlier -- the z one is the lane's own previous result,
// the x one comes from lane - 1 and the y one from lane - 8 by shuffle -- so a front that crosses the tile within one
// octant of directions is final after ONE sweep of 26 steps with one evaluation per cell.  Sweeps that run against
// the front find no cell with a lower sweep-upwind neighbour and skip their steps after one vote.  Check passes
// (256 threads x 2 cells, Jacobi) decide convergence: the visit ends when a pass changes nothing -- the same
// epsilon = 0 fixed point as the warp engine of eikonal3d.cuh.
//
// Work order: the local causal order of the 2D sweep engine, with six neighbours.  A popped tile waits while a
// neighbour tile that is queued or running carries a LOWER priority (= lowest value that activated it): that
// neighbour is upwind and will still lower this tile's halo.  The warp engine's plain FIFO order visited a tile 31
// times at 256^3 (70 evaluations per cell, 35x the algorithmic DRAM traffic).
#pragma once
#include "eikonal3d.cuh"

namespace fmb {

__device__ __forceinline__ int nbr_off3d(int s, int ntx, int ntz) {       // s = z-, z+, x-, x+, y-, y+
    return s == 0 ? -1 : s == 1 ? 1 : s == 2 ? -ntz : s == 3 ? ntz : s == 4 ? -ntx * ntz : ntx * ntz;
}

// Take the next tile for this CTA (all 32 lanes of warp 0).  Returns the tile index (already RUNNING, fenced) or -1.
template <typename real>
__device__ __forceinline__ int cta_acquire3d(const Problem3D<real> &P, int lane, int &streak, unsigned long long &n_defer) {
    const int tiles_per_q = P.nty * P.ntx * P.ntz;
    const unsigned long long PRIO_INF = 0x7ff0000000000000ULL;
    int item;
    for (;;) {
        int it = -1;
        if (lane == 0) it = q_pop_lane0(P.q);
        item = __shfl_sync(FULL, it, 0);
        if (item < 0 || !P.causal) break;
        int t = item % tiles_per_q;
        const int tz = t % P.ntz; t /= P.ntz;
        const int tx = t % P.ntx, ty = t / P.ntx;
        const unsigned long long mine = *reinterpret_cast<const volatile unsigned long long *>(&P.tile_prio[item]);
        const double slack = *reinterpret_cast<const volatile double *>(P.slack);
        bool blocked = false;
        if (lane < 24) {
            // lanes 0-5: the six face neighbours.  lanes 6-23: the second ring (12 edge diagonals, 6 tiles two steps away
            // in a straight line), counted only when more than `hop` below my priority (see eikonal2d_cta.cuh)
            const int dz_[24] = {-1, 1, 0, 0, 0, 0, /*xz*/ -1, 1, -1, 1, /*yz*/ -1, 1, -1, 1, /*xy*/ 0, 0, 0, 0, -2, 2, 0, 0, 0, 0};
            const int dx_[24] = {0, 0, -1, 1, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, -1, 1, -1, 1, 0, 0, -2, 2, 0, 0};
            const int dy_[24] = {0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, -1, 1, 1, -1, -1, 1, 1, 0, 0, 0, 0, -2, 2};
            const int nz = tz + dz_[lane], nx = tx + dx_[lane], ny = ty + dy_[lane];
            const bool ring2 = lane >= 6;
            if (nz >= 0 && nz < P.ntz && nx >= 0 && nx < P.ntx && ny >= 0 && ny < P.nty && !(ring2 && P.hop_frac <= 0.0)) {
                const int n = item + (dy_[lane] * P.ntx + dx_[lane]) * P.ntz + dz_[lane];
                const int st = ld_volatile(&P.tile_state[n]);
                unsigned long long key = ~0ULL;
                if (st == ST_QUEUED || st == ST_DIRTY) key = *reinterpret_cast<const volatile unsigned long long *>(&P.tile_prio[n]);
                if (st == ST_RUNNING || st == ST_DIRTY) {
                    const unsigned long long rk = *reinterpret_cast<const volatile unsigned long long *>(&P.run_prio[n]);
                    key = rk < key ? rk : key;
                }
                // (~0 reads as NaN: never blocks)
                const double tol = ring2 ? *reinterpret_cast<const volatile double *>(P.slack + 1) : slack;
                blocked = __longlong_as_double((long long)key) + tol < __longlong_as_double((long long)mine);
            }
        }
        const bool any_blocked = __any_sync(FULL, blocked) && streak < 100000;
        if (!any_blocked) { streak = 0; break; }
        if (lane == 0) q_push(P.q, item);
        ++streak;
        n_defer += 1;
        __nanosleep(streak < 4 ? 100u : streak < 12 ? 400u : 1000u);
    }
    if (item < 0) return -1;
    // QUEUED -> RUNNING *before* T is sampled: anything published after this point flips the state to DIRTY
    if (lane == 0) {
        if (P.causal) {
            const unsigned long long key = atomicExch(&P.tile_prio[item], PRIO_INF);
            *reinterpret_cast<volatile unsigned long long *>(&P.run_prio[item]) = key;
        }
        atomicExch(&P.tile_state[item], ST_RUNNING);
        __threadfence();
    }
    __syncwarp();
    return item;
}

template <typename real, bool EXACT>
__device__ __forceinline__ real solve3d_update_any(real tx, real ty, real tz, real c) {
    if (EXACT) return (real)solve3d_update_exact((double)tx, (double)ty, (double)tz, (double)c);
    return solve3d_update<real>(tx, ty, tz, c);
}
// warp-collective: every lane evaluates (branch-free for fp64), lanes outside the fast paths are redone after a vote
template <typename real, bool EXACT>
__device__ __forceinline__ real solve3d_update_warp(real tx, real ty, real tz, real c, bool want) {
    if (sizeof(real) == 8) {
        bool slow = false;
        real v = (real)solve3d_update_sel<EXACT>((double)tx, (double)ty, (double)tz, (double)c, slow);
        if (__any_sync(FULL, want && slow)) {
            if (want && slow) v = solve3d_update_any<real, EXACT>(tx, ty, tz, c);
            __syncwarp();
        }
        return v;
    }
    return want ? solve3d_update_any<real, EXACT>(tx, ty, tz, c) : num<real>::inf();
}

template <typename real, bool EXACT>
__global__ void __launch_bounds__(256) solve3d_sweep_kernel(Problem3D<real> P) {
    if (P.enable && !*P.enable) return;
    constexpr int TZ = 16;
    using TL = Tile3D<real, TZ>;
    constexpr int PZ = TL::PZ, PS = TL::PS, NSTEP = T3Y + T3X + TZ - 2;
    FMB_DYN_SMEM(smem_raw);
    real *sT = reinterpret_cast<real *>(smem_raw);
    real *sC = sT + TL::T_ELEMS;
    real *sFace = sC + TL::C_ELEMS;                                          // [8] lowest changed value per face that undercuts the halo
    unsigned *sDirty = reinterpret_cast<unsigned *>(sFace + 8);              // [32] changed cells per column since the last write-back
    int *sCtl = reinterpret_cast<int *>(sDirty + 32);                        // [0] tile, [1] stop, [3] continue in place
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const real INF = num<real>::inf();
    const int tiles_per_q = P.nty * P.ntx * P.ntz;
    const real UP = EXACT ? (real)(1.0 + 1e-11) : (real)(1.0 + 8.0 / 4503599627370496.0);
    const long long sy_ = (long long)P.nx * P.nz, sx_ = P.nz;               // global strides

    unsigned long long n_visits = 0, n_steps = 0, n_pushes = 0, n_defer = 0, n_rounds = 0, n_cont = 0;
    unsigned my_evals = 0, my_written = 0;
    int streak = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0, c_check = 0;

    for (;;) {
        const long long tc0 = clock64();
        if (warp == 0) {
            const int it = cta_acquire3d<real>(P, lane, streak, n_defer);
            if (lane == 0) sCtl[0] = it;
        }
        if (tid < 32) sDirty[tid] = 0;
        __syncthreads();
        const int item = sCtl[0];
        if (item < 0) break;
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        int t = item - q * tiles_per_q;
        const int tz = t % P.ntz; t /= P.ntz;
        const int tx = t % P.ntx; const int ty = t / P.ntx;
        const int x0 = tx * T3X, y0 = ty * T3Y, z0 = tz * TZ;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- stage T (60 z-columns incl. lateral halos, minus the 4 corner columns), the z halos and the costs ----
        {
            constexpr int NCOL = (T3Y + 2) * (T3X + 2);
            constexpr int EPC = 16 / (int)sizeof(real);
            constexpr int CPC = TZ / EPC;
            const bool fast = sizeof(real) == 8 && y0 >= 1 && y0 + T3Y < P.ny && x0 >= 1 && x0 + T3X < P.nx &&
                              z0 + TZ <= P.nz && (P.nz % EPC) == 0 && ((size_t)Tq % 16) == 0 && ((size_t)cq % 16) == 0;
            if (fast) {
                for (int c = tid; c < NCOL * CPC; c += 256) {
                    const int col = c / CPC, zz = (c % CPC) * EPC;
                    const int yy = col / (T3X + 2) - 1, xx = col % (T3X + 2) - 1;
                    const bool corner = (yy < 0 || yy >= T3Y) && (xx < 0 || xx >= T3X);
                    if (!corner) cp_async16_cg(&sT[TL::at(yy, xx, zz)], &Tq[(y0 + yy) * sy_ + (x0 + xx) * sx_ + z0 + zz]);
                }
                {
                    const int c = tid;                     // 32 columns x 8 chunks = 256
                    const int col = c / CPC, zz = (c % CPC) * EPC;
                    if (col < 32) cp_async16_cg(&sC[col * PZ + zz], &cq[(y0 + (col >> 3)) * sy_ + (x0 + (col & 7)) * sx_ + z0 + zz]);
                }
                cp_async_wait_all();
            } else {
                for (int idx = tid; idx < NCOL * TZ; idx += 256) {
                    const int col = idx / TZ, zz = idx % TZ;
                    const int yy = col / (T3X + 2) - 1, xx = col % (T3X + 2) - 1;
                    const bool corner = (yy < 0 || yy >= T3Y) && (xx < 0 || xx >= T3X);
                    if (corner) continue;
                    const int gy = y0 + yy, gx = x0 + xx, gz = z0 + zz;
                    real v = INF;
                    if (gy >= 0 && gy < P.ny && gx >= 0 && gx < P.nx && gz < P.nz) v = ld_T(&Tq[gy * sy_ + gx * sx_ + gz]);
                    sT[TL::at(yy, xx, zz)] = v;
                }
                for (int idx = tid; idx < 32 * TZ; idx += 256) {
                    const int col = idx / TZ, zz = idx % TZ;
                    const int gy = y0 + (col >> 3), gx = x0 + (col & 7), gz = z0 + zz;
                    real c = INF;
                    if (gy < P.ny && gx < P.nx && gz < P.nz) c = __ldg(&cq[gy * sy_ + gx * sx_ + gz]);
                    sC[col * PZ + zz] = c;
                }
            }
            if (tid < 64) {                                // z halos of the 32 interior columns
                const int col = tid & 31, yy = col >> 3, xx = col & 7;
                const int gy = y0 + yy, gx = x0 + xx, gz = tid < 32 ? z0 - 1 : z0 + TZ;
                real v = INF;
                if (gy < P.ny && gx < P.nx && gz >= 0 && gz < P.nz) v = ld_T(&Tq[gy * sy_ + gx * sx_ + gz]);
                sT[TL::at(yy, xx, tid < 32 ? -1 : TZ)] = v;
            }
        }
        __syncthreads();
        const long long tc2 = clock64();

        // ---- per-warp sweep geometry: lane (ly, lx) in sweep order -> column (jy, jx) ----
        const int sz = (warp & 1) ? -1 : 1, sx = (warp & 2) ? -1 : 1, sy = (warp & 4) ? -1 : 1;
        const int ly = lane >> 3, lx = lane & 7;
        const int jy = sy > 0 ? ly : T3Y - 1 - ly, jx = sx > 0 ? lx : T3X - 1 - lx;
        volatile real *colT = sT + TL::at(jy, jx, 0);
        const real *colC = sC + (jy * T3X + jx) * PZ;
        const int dvx = sx > 0 ? PZ : -PZ, dvy = sy > 0 ? PS : -PS;      // towards the sweep-downwind column
        int steps = 0, round = 0;
        long long t_rounds = 0;
        // Octant rule (variant bit 3).  With the loose rule "a sweep-upwind neighbour lies below me" seven of the eight
        // sweeps re-evaluate nearly every cell to no effect, and with four sweep warps per scheduler those evaluations are
        // what a step costs.  A sweep is the right one for a cell when, on EVERY axis, the lower of the cell's two
        // neighbours lies on the sweep's upwind side (ties, incl. two +inf neighbours, go to the positive orientation):
        // then the sweep's own chain carries every input of the cell's update.  Each cell is evaluated by one sweep per
        // round; the others skip the step after one vote.  (The check passes ignore the rule: it costs rounds at worst.)
        const bool octant = (P.variant & 8) != 0;
        auto ax_ok = [](real up, real dw, int sgn) -> bool { return up < dw || (up == dw && sgn > 0); };
        for (;;) {      // continuation loop: one pass per (re-)activation served in place
            const long long tr0 = clock64();
            for (;; ++round) {
                // ---- check passes: one Jacobi relaxation of every cell (thread = column tid >> 3, cells 2 * (tid & 7) .. +1,
                // all inputs loaded before the first store); converged when a pass changes nothing.  Round 0 of a fresh
                // visit goes straight to the sweeps.
                const long long tk0 = clock64();
                int again = 1;
                const int npass = round == 0 ? 0 : P.check_passes;
                for (int pass = 0; pass < npass; ++pass) {
                    const int col = tid >> 3, zb = 2 * (tid & 7);
                    real *p = sT + TL::at(col >> 3, col & 7, zb);
                    const real *pc = sC + col * PZ + zb;
                    const real m0 = p[-1], m1 = p[0], m2 = p[1], m3 = p[2];
                    const real xa0 = p[-PZ], xb0 = p[PZ], ya0 = p[-PS], yb0 = p[PS];
                    const real xa1 = p[1 - PZ], xb1 = p[1 + PZ], ya1 = p[1 - PS], yb1 = p[1 + PS];
                    const real c0 = pc[0], c1 = pc[1];
                    unsigned dbits = 0;
                    {
                        const real ax = xa0 < xb0 ? xa0 : xb0, ay = ya0 < yb0 ? ya0 : yb0, az = m0 < m2 ? m0 : m2;
                        const bool want = (ax < m1 || ay < m1 || az < m1) && c0 < INF;
                        const real v = solve3d_update_warp<real, EXACT>(ax, ay, az, c0, want);
                        if (want) {
                            ++my_evals;
                            if (v != m1 && v <= num<real>::mul(m1, UP)) { p[0] = v; dbits |= 1u << zb; }
                        }
                    }
                    {
                        const real ax = xa1 < xb1 ? xa1 : xb1, ay = ya1 < yb1 ? ya1 : yb1, az = m1 < m3 ? m1 : m3;
                        const bool want = (ax < m2 || ay < m2 || az < m2) && c1 < INF;
                        const real v = solve3d_update_warp<real, EXACT>(ax, ay, az, c1, want);
                        if (want) {
                            ++my_evals;
                            if (v != m2 && v <= num<real>::mul(m2, UP)) { p[1] = v; dbits |= 2u << zb; }
                        }
                    }
                    if (dbits) atomicOr(&sDirty[col], dbits);
                    steps += 2;
                    again = __syncthreads_or(dbits != 0);
                    if (!again) break;
                }
                c_check += clock64() - tk0;
                if (!again) break;
                if (steps > P.step_cap) break;
                ++n_rounds;

                // ---- octant gate: which of the eight sweeps have any cell to evaluate?  One pass of compares (no update):
                // a cell belongs to the sweep whose upwind side carries the lower neighbour on every axis, if any neighbour
                // lies below it.  A tile far from the source holds one to three octants; the other sweep warps skip the
                // round instead of competing for issue slots (eight warps x ~200 instructions per step otherwise).
                if (octant) {
                    if (tid == 0) sCtl[2] = 0;
                    __syncthreads();
                    const int col = tid >> 3, zb = 2 * (tid & 7);
                    const real *p = sT + TL::at(col >> 3, col & 7, zb);
                    const real *pc = sC + col * PZ + zb;
                    unsigned om = 0;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const real cur = p[k], zm = p[k - 1], zp = p[k + 1], xm = p[k - PZ], xp = p[k + PZ], ym = p[k - PS], yp = p[k + PS];
                        const real lz = zm < zp ? zm : zp, lxv = xm < xp ? xm : xp, lyv = ym < yp ? ym : yp;
                        if ((lz < cur || lxv < cur || lyv < cur) && pc[k] < INF)
                            om |= 1u << ((zp < zm ? 1 : 0) | (xp < xm ? 2 : 0) | (yp < ym ? 4 : 0));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) om |= __shfl_xor_sync(FULL, om, o);
                    if (lane == 0 && om) atomicOr(reinterpret_cast<unsigned *>(&sCtl[2]), om);
                    __syncthreads();
                    if (!((sCtl[2] >> warp) & 1)) { __syncthreads(); steps += NSTEP; continue; }
                }

                // ---- one sweep per warp.  Lane (ly, lx) relaxes cell kz = d - ly - lx of its column at step d.
                int kz = -ly - lx;
                real res = colT[sz > 0 ? -1 : TZ];             // before my first cell: the z halo
                unsigned dirty = 0;
                bool hot = false;
                int zc = sz > 0 ? min(max(kz, 0), TZ - 1) : TZ - 1 - min(max(kz, 0), TZ - 1);
                real n_cur = colT[zc], n_c = colC[zc], n_dwz = colT[zc + sz], n_dwx = colT[zc + dvx], n_dwy = colT[zc + dvy];
                real n_upx0 = colT[zc - dvx], n_upy0 = colT[zc - dvy];
                if (!EXACT && (P.variant & 1)) {
                    // straight-line step (see eikonal2d_sweep.cuh): the update is evaluated in every step and applied by
                    // predication, the cell's value "as it is now" is read as late as possible
                    for (int d = 0; d < NSTEP; ++d, ++kz) {
                        const bool valid = (unsigned)kz < (unsigned)TZ;
                        const int z = zc;
                        const real cur = n_cur, c = n_c, dwz = n_dwz, dwx = n_dwx, dwy = n_dwy, upx0 = n_upx0, upy0 = n_upy0;
                        zc = sz > 0 ? min(max(kz + 1, 0), TZ - 1) : TZ - 1 - min(max(kz + 1, 0), TZ - 1);
                        n_cur = colT[zc]; n_c = colC[zc]; n_dwz = colT[zc + sz]; n_dwx = colT[zc + dvx]; n_dwy = colT[zc + dvy];
                        if (lx == 0) n_upx0 = colT[zc - dvx];
                        if (ly == 0) n_upy0 = colT[zc - dvy];
                        real upx = __shfl_up_sync(FULL, res, 1);
                        real upy = __shfl_up_sync(FULL, res, 8);
                        if (lx == 0) upx = upx0;
                        if (ly == 0) upy = upy0;
                        const bool go = valid && (res < cur || upx < cur || upy < cur) && c < INF &&
                                        (!octant || (ax_ok(res, dwz, sz) && ax_ok(upx, dwx, sx) && ax_ok(upy, dwy, sy)));
                        const real v = solve3d_update_warp<real, EXACT>(upx < dwx ? upx : dwx, upy < dwy ? upy : dwy, res < dwz ? res : dwz, c, go);
                        const bool acc = go && v != cur && v <= num<real>::mul(cur, UP);
                        const real now = colT[z];
                        const bool st = acc && v != now && v <= num<real>::mul(now, UP);
                        if (st) colT[z] = v;
                        dirty |= st ? 1u << z : 0u;
                        my_evals += go;
                        res = valid ? (acc ? v : cur) : res;
                    }
                } else
                for (int d = 0; d < NSTEP; ++d, ++kz) {
                    const bool valid = (unsigned)kz < (unsigned)TZ;
                    const int z = zc;
                    const real cur = n_cur, c = n_c, dwz = n_dwz, dwx = n_dwx, dwy = n_dwy, upx0 = n_upx0, upy0 = n_upy0;
                    zc = sz > 0 ? min(max(kz + 1, 0), TZ - 1) : TZ - 1 - min(max(kz + 1, 0), TZ - 1);
                    n_cur = colT[zc]; n_c = colC[zc]; n_dwz = colT[zc + sz]; n_dwx = colT[zc + dvx]; n_dwy = colT[zc + dvy];
                    if (lx == 0) n_upx0 = colT[zc - dvx];
                    if (ly == 0) n_upy0 = colT[zc - dvy];
                    real upx = __shfl_up_sync(FULL, res, 1);
                    real upy = __shfl_up_sync(FULL, res, 8);
                    if (lx == 0) upx = upx0;
                    if (ly == 0) upy = upy0;
                    const bool go = valid && (res < cur || upx < cur || upy < cur) && c < INF &&
                                    (!octant || (ax_ok(res, dwz, sz) && ax_ok(upx, dwx, sx) && ax_ok(upy, dwy, sy)));
                    real out = cur;
                    if ((hot && !octant) || __any_sync(FULL, go)) {
                        const real v = solve3d_update_warp<real, EXACT>(upx < dwx ? upx : dwx, upy < dwy ? upy : dwy, res < dwz ? res : dwz, c, go);
                        if (go) {
                            ++my_evals;
                            if (v != cur && v <= num<real>::mul(cur, UP)) {
                                out = v;
                                // `cur` was loaded a step ago: weigh the value against the cell as it is NOW, so that a lower
                                // value another sweep stored meanwhile is not overwritten
                                const real now = colT[z];
                                if (v != now && v <= num<real>::mul(now, UP)) {
                                    colT[z] = v;
                                    dirty |= 1u << z;
                                }
                            }
                        }
                        __syncwarp();
                    }
                    hot = __any_sync(FULL, go);
                    if (valid) res = out;
                }
                if (dirty) atomicOr(&sDirty[jy * T3X + jx], dirty);
                __syncthreads();
                steps += NSTEP;
            }
            t_rounds += clock64() - tr0;
            if (steps > P.step_cap) break;

            // ---- write back changed cells (half warp = one column, lane & 15 = z: 128-byte rows) ----
            {
                const int zz = tid & 15;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int col = (tid >> 4) + 16 * rr;
                    if ((sDirty[col] >> zz) & 1u) {
                        st_T(&Tq[(y0 + (col >> 3)) * sy_ + (x0 + (col & 7)) * sx_ + z0 + zz], sT[TL::at(col >> 3, col & 7, zz)]);
                        ++my_written;
                    }
                }
            }
            // ---- per face: the lowest changed value that undercuts the halo (warp s = face s: z-, z+, x-, x+, y-, y+) ----
            if (warp < 6) {
                real m = INF;
                if (warp < 2) {                                // z faces: 32 columns, one cell each
                    const int zz = warp == 0 ? 0 : TZ - 1, zh = warp == 0 ? -1 : TZ;
                    const int yy = lane >> 3, xx = lane & 7;
                    const real v = sT[TL::at(yy, xx, zz)];
                    if (((sDirty[lane] >> zz) & 1u) && v < sT[TL::at(yy, xx, zh)]) m = v;
                } else if (warp < 4) {                         // x faces: 4 columns x 16 cells
                    const int xx = warp == 2 ? 0 : T3X - 1, xh = warp == 2 ? -1 : T3X;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int idx = lane + 32 * k, yy = idx >> 4, zz = idx & 15;
                        const real v = sT[TL::at(yy, xx, zz)];
                        if (((sDirty[yy * T3X + xx] >> zz) & 1u) && v < sT[TL::at(yy, xh, zz)] && v < m) m = v;
                    }
                } else {                                       // y faces: 8 columns x 16 cells
                    const int yy = warp == 4 ? 0 : T3Y - 1, yh = warp == 4 ? -1 : T3Y;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = lane + 32 * k, xx = idx >> 4, zz = idx & 15;
                        const real v = sT[TL::at(yy, xx, zz)];
                        if (((sDirty[yy * T3X + xx] >> zz) & 1u) && v < sT[TL::at(yh, xx, zz)] && v < m) m = v;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const real v = __shfl_xor_sync(FULL, m, o);
                    m = v < m ? v : m;
                }
                if (lane == 0) sFace[warp] = m;
            }
            __syncthreads();          // every T store of the CTA and the face minima precede warp 0's fence below

            // ---- publish: activate the neighbours across undercut faces; retire, or continue in place ----
            if (warp == 0) {
                const real mine = lane < 6 ? sFace[lane] : INF;
                bool ex = false;
                if (lane < 6) ex = lane == 0 ? tz > 0 : lane == 1 ? tz < P.ntz - 1 : lane == 2 ? tx > 0 : lane == 3 ? tx < P.ntx - 1
                                   : lane == 4 ? ty > 0 : ty < P.nty - 1;
                const bool is_nbr = lane < 6 && ex && mine < INF;
                const unsigned act = __ballot_sync(FULL, is_nbr);
                const int nact = __popc(act);
                if (lane == 0 && nact) atomicAdd(&P.q.ctl->pending, nact);
                const int tgt = lane < 6 ? item + nbr_off3d(lane, P.ntx, P.ntz) : item;
                if (P.causal && is_nbr) atomicMin(&P.tile_prio[tgt], (unsigned long long)__double_as_longlong((double)mine));
                __threadfence();          // the CTA's T stores (+ pending, priorities) are device-visible ...
                __syncwarp();             // ... before any state transition is published
                bool pushed = false, newly = false, cont = false;
                {
                    const bool is_self = lane == 6;
                    if (is_nbr || is_self) {
                        int *st = &P.tile_state[tgt];
                        int old = atomicCAS(st, is_self ? ST_RUNNING : ST_IDLE, is_self ? ST_IDLE : ST_QUEUED);
                        if (is_self) {
                            if (old != ST_RUNNING) { atomicExch(st, ST_RUNNING); cont = true; }   // re-activated while it ran: serve that here
                        } else {
                            for (;;) {
                                if (old == ST_IDLE) { newly = true; break; }
                                if (old == ST_QUEUED || old == ST_DIRTY) break;
                                if (atomicCAS(st, ST_RUNNING, ST_DIRTY) == ST_RUNNING) break;          // ask the runner to look again
                                old = atomicCAS(st, ST_IDLE, ST_QUEUED);
                            }
                        }
                        if (newly) { q_push(P.q, tgt); pushed = true; }
                    }
                }
                const int n_new = __popc(__ballot_sync(FULL, newly));
                const int n_cnt = __popc(__ballot_sync(FULL, cont));
                n_pushes += __popc(__ballot_sync(FULL, pushed));
                if (lane == 0) {
                    const int drop = (nact - n_new) + (n_cnt ? 0 : 1);
                    if (drop) atomicSub(&P.q.ctl->pending, drop);
                    sCtl[1] = ld_volatile(&P.q.ctl->abort);
                    sCtl[3] = n_cnt;
                    if (n_cnt) __threadfence();        // RUNNING again before the halo is sampled
                }
            }
            __syncthreads();
            if (!sCtl[3] || sCtl[1]) break;
            if (tid < 32) sDirty[tid] = 0;                         // what was written back is published
            // ---- continue in place: fresh halo shell, then a check pass decides whether anything is left to do ----
            if (round == 0) round = 1;
            ++n_cont;
            {
                // 2 x 32 z-halo cells, 2 x 64 x-halo cells, 2 x 128 y-halo cells = 448 loads over 256 threads
                for (int idx = tid; idx < 448; idx += 256) {
                    int yy, xx, zz;
                    if (idx < 64) { const int c = idx & 31; yy = c >> 3; xx = c & 7; zz = idx < 32 ? -1 : TZ; }
                    else if (idx < 192) { const int k = idx - 64, f = k >> 6, c = k & 63; yy = c >> 4; zz = c & 15; xx = f ? T3X : -1; }
                    else { const int k = idx - 192, f = k >> 7, c = k & 127; xx = c >> 4; zz = c & 15; yy = f ? T3Y : -1; }
                    const int gy = y0 + yy, gx = x0 + xx, gz = z0 + zz;
                    real v = INF;
                    if (gy >= 0 && gy < P.ny && gx >= 0 && gx < P.nx && gz >= 0 && gz < P.nz) v = ld_T(&Tq[gy * sy_ + gx * sx_ + gz]);
                    sT[TL::at(yy, xx, zz)] = v;
                }
            }
            __syncthreads();
        }
        n_steps += steps;
        ++n_visits;
        if (steps > P.step_cap) {
            if (tid == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_STEPCAP);
            break;
        }
        const int stop = sCtl[1];
        const long long tc4 = clock64();
        c_wait += tc1 - tc0; c_load += tc2 - tc1; c_relax += t_rounds; c_store += tc4 - tc2 - t_rounds;
        if (stop) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_evals += __shfl_xor_sync(FULL, my_evals, o);
        my_written += __shfl_xor_sync(FULL, my_written, o);
    }
    if (lane == 0) {
        atomicAdd(&P.q.ctl->evals, (unsigned long long)my_evals);
        atomicAdd(&P.q.ctl->cells_written, (unsigned long long)my_written);
    }
    if (tid == 0) {
        atomicAdd(&P.q.ctl->cyc_wait, (unsigned long long)c_wait);
        atomicAdd(&P.q.ctl->cyc_load, (unsigned long long)c_load);
        atomicAdd(&P.q.ctl->cyc_relax, (unsigned long long)c_relax);
        atomicAdd(&P.q.ctl->cyc_store, (unsigned long long)c_store);
        atomicAdd(&P.q.ctl->visits, n_visits);
        atomicAdd(&P.q.ctl->steps, n_steps);
        atomicAdd(&P.q.ctl->pushes, n_pushes);
        if (n_defer) atomicAdd(&P.q.ctl->pad[0], n_defer);
        atomicAdd(&P.q.ctl->pad[1], (unsigned long long)c_check);
        atomicAdd(&P.q.ctl->rounds, n_rounds);
        atomicAdd(&P.q.ctl->continuations, n_cont);
    }
}

}  // namespace fmb
