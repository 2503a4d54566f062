// eikonal2d_wsweep.cuh -- 2D Eikonal solve, warp-per-tile sweep engine for BATCHES of queries
// (replaces FastMarching.py:17-29,44-112, many queries at once).
//
// A batch of independent queries is a throughput problem: thousands of tiles are runnable at any time, so a tile
// visit should cost as few instruction slots and as little shared memory as possible, and latency is hidden by
// having many visits resident per SM -- the opposite trade of the four-warp sweep visit of eikonal2d_sweep.cuh
// (shortest possible visit, for the chain of dependent visits of ONE map).
//
//   * one WARP per tile visit (32 x 32 tile, lane = row); the four diagonal-wavefront Gauss-Seidel sweeps of the
//     sweep engine run one after another on that warp, the one that runs WITH the front first: its orientation is
//     read off the halo (which side carries the lowest value);
//   * after that first sweep a cell can only be off its fixed point if a neighbour on the sweep's downwind side
//     ended below it (those are the only inputs that changed after the cell was evaluated): one pass of two
//     compares per cell decides, with no evaluation of the update, whether the visit is over -- the common case
//     when a front crosses the tile within one quadrant of directions.  Otherwise the other three sweeps run and
//     full Jacobi check passes (lane = row, 8 cells at a time) iterate to the epsilon = 0 fixed point as in the
//     other engines;
//   * STAGE_C = false keeps only the T tile in shared memory (9.3 KB per warp instead of 17.9 KB -> twice the
//     resident warps): the cost of a cell is read from global memory one sweep step ahead of its use (a batch on one
//     map shares a cost array that lives in L2);
//   * the per-query best-first order of eikonal2d.cuh (BEST) picks the tile.
//
// Measured against the armed-cell warp visit of eikonal2d.cuh (one armed cell per lane per lock-step iteration,
// 119 instructions per iteration, ~8 useful lanes): see DESIGN.md 5.
#pragma once
#include "eikonal2d_cta.cuh"

namespace fmb {

// One diagonal-wavefront sweep of a 32 x 32 shared tile by one warp (lane = row in sweep order).
// Returns the bit mask of the cells of this lane's row that changed.
template <typename real, bool STAGE_C>
__device__ __forceinline__ unsigned wsweep_dir(real *sT, const real *sC, const real *cq, long long cost_pitch, int x0, int y0,
                                               int rows, int cols, int sx, int sy, int lane, unsigned &evals, bool &bad_cost) {
    using TL = Tile2D<real, 32>;
    constexpr int PT = TL::PT, TW = 32, NSTEP = TILE_H + TW - 1;
    const real INF = num<real>::inf();
    const real UP = (real)(1.0 + 8.0 / 4503599627370496.0);
    const int jrow = sy > 0 ? lane : TILE_H - 1 - lane;
    real *rowT = sT + (jrow + 1) * PT + 2;
    const real *rowC = STAGE_C ? sC + jrow * PT : cq + (long long)(y0 + jrow) * cost_pitch + x0;
    const bool row_in = y0 + jrow < rows;
    const int cmax = cols - x0;                    // columns of this tile inside the map
    const int dv = sy > 0 ? PT : -PT;              // towards the sweep-downwind row
    const int hcol = sx > 0 ? -1 : TW;             // my upwind halo column (tile coordinates)
    int i = sx > 0 ? -lane : TW - 1 + lane;
    real res = rowT[hcol];
    unsigned dirty = 0;
    bool hot = false;
    int ic = min(max(i, 0), TW - 1);
    auto cost_at = [&](int k) -> real {
        if (STAGE_C) return rowC[k];
        return (row_in && k < cmax) ? __ldg(&rowC[k]) : INF;
    };
    real n_cur = rowT[ic], n_c = cost_at(ic), n_dwh = rowT[ic + sx], n_dwv = rowT[ic + dv], n_up0 = rowT[ic - dv];
    for (int d = 0; d < NSTEP; ++d, i += sx) {
        const bool valid = (unsigned)i < (unsigned)TW;
        const real cur = n_cur, c = n_c, dwh = n_dwh, dwv = n_dwv, up0 = n_up0;
        const int iw = ic;                         // == i when valid
        ic = min(max(i + sx, 0), TW - 1);
        n_cur = rowT[ic]; n_c = cost_at(ic); n_dwh = rowT[ic + sx]; n_dwv = rowT[ic + dv];
        if (lane == 0) n_up0 = rowT[ic - dv];
        real up = __shfl_up_sync(FULL, res, 1);
        if (lane == 0) up = up0;
        const bool go = valid && (res < cur || up < cur) && c < INF;
        real out = cur;
        if (hot || __any_sync(FULL, go)) {
            const real v = eikonal_update_sel<real>(res < dwh ? res : dwh, up < dwv ? up : dwv, c);
            evals += go;
            bad_cost |= go && !(c >= cost_range<real>::lo && c <= cost_range<real>::hi);
            if (go && v != cur && v <= num<real>::mul(cur, UP)) {
                out = v;
                rowT[iw] = v;
                dirty |= 1u << iw;
            }
        }
        hot = __any_sync(FULL, go);
        if (valid) res = out;
    }
    __syncwarp();                                  // the sweep's stores are visible to every lane from here on
    return dirty;
}

template <typename real, bool BEST, bool STAGE_C>
__global__ void __launch_bounds__(128, STAGE_C ? 3 : 4) solve2d_wsweep_kernel(Problem2D<real> P) {
    using TL = Tile2D<real, 32>;
    constexpr int PT = TL::PT, TW = 32, NSTEP = TILE_H + TW - 1;
    constexpr int WELEMS = TL::T_ELEMS + (STAGE_C ? TL::C_ELEMS : 0);
    constexpr int CB = 4;                              // cells per lane per block of the check passes
    FMB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    real *sT = reinterpret_cast<real *>(smem_raw) + (size_t)warp * WELEMS;
    real *sC = sT + TL::T_ELEMS;                       // valid only when STAGE_C
    const real INF = num<real>::inf();
    const int tiles_per_q = P.ntx * P.nty;
    const real UP = (real)(1.0 + 8.0 / 4503599627370496.0);

    unsigned long long n_visits = 0, n_steps = 0, n_pushes = 0, n_defer = 0, n_noop = 0, n_rounds = 0, n_written = 0;
    unsigned my_evals = 0;
    int streak = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0, c_check = 0;

    for (;;) {
        const long long tc0 = clock64();
        const int item = cta_acquire<real, BEST>(P, lane, streak, n_defer);
        if (item < 0) break;
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        const int t = item - q * tiles_per_q;
        const int ty = t / P.ntx, tx = t - ty * P.ntx;
        const int x0 = tx * TW, y0 = ty * TILE_H;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- stage T (+ halo) and, when STAGE_C, the costs ----
        unsigned cmask = 0;
        {
            constexpr int EPC = 16 / (int)sizeof(real);
            constexpr int CPR = TW / EPC;
            const int y = y0 + lane;
            real vl = INF, vr = INF;
            if (y < P.rows) {
                if (x0 > 0) vl = ld_T(&Tq[(long long)y * P.T_pitch + x0 - 1]);
                if (x0 + TW < P.cols) vr = ld_T(&Tq[(long long)y * P.T_pitch + x0 + TW]);
            }
            const bool fast = sizeof(real) == 8 && y0 >= 1 && y0 + TILE_H < P.rows && x0 + TW <= P.cols &&
                              (P.T_pitch % EPC) == 0 && (P.cost_pitch % EPC) == 0 &&
                              ((size_t)Tq % 16) == 0 && ((size_t)cq % 16) == 0;
            if (fast) {
#pragma unroll
                for (int c = lane; c < (TILE_H + 2) * CPR; c += 32) {
                    const int row = c / CPR, col = (c % CPR) * EPC;
                    cp_async16_cg(&sT[row * PT + 2 + col], &Tq[(long long)(y0 - 1 + row) * P.T_pitch + x0 + col]);
                }
                if (STAGE_C) {
#pragma unroll
                    for (int c = lane; c < TILE_H * CPR; c += 32) {
                        const int row = c / CPR, col = (c % CPR) * EPC;
                        cp_async16_cg(&sC[row * PT + col], &cq[(long long)(y0 + row) * P.cost_pitch + x0 + col]);
                    }
                }
                cp_async_wait_all();
            } else {
                for (int j = -1; j <= TILE_H; ++j) {
                    const int yy = y0 + j, xx = x0 + lane;
                    real v = INF;
                    if (yy >= 0 && yy < P.rows && xx < P.cols) v = ld_T(&Tq[(long long)yy * P.T_pitch + xx]);
                    sT[(j + 1) * PT + lane + 2] = v;
                }
                if (STAGE_C) {
                    for (int j = 0; j < TILE_H; ++j) {
                        const int yy = y0 + j, xx = x0 + lane;
                        real c = INF;
                        if (yy < P.rows && xx < P.cols) c = __ldg(&cq[(long long)yy * P.cost_pitch + xx]);
                        sC[j * PT + lane] = c;
                    }
                }
            }
            sT[(lane + 1) * PT + 1] = vl;
            sT[(lane + 1) * PT + TW + 2] = vr;
            __syncwarp();
            // cells that can ever be relaxed (finite cost), one bit mask per row
#pragma unroll 8
            for (int j = 0; j < TILE_H; ++j) {
                real c;
                if (STAGE_C) c = sC[j * PT + lane];
                else c = (y0 + j < P.rows && x0 + lane < P.cols) ? __ldg(&cq[(long long)(y0 + j) * P.cost_pitch + x0 + lane]) : INF;
                const unsigned bal = __ballot_sync(FULL, c < INF);
                if (lane == j) cmask = bal;
            }
        }
        __syncwarp();
        const long long tc2 = clock64();

        // ---- orientation of the first sweep: with the front, i.e. away from the side that carries the lowest halo value
        real *rowT = sT + (lane + 1) * PT + 2;
        int sx0, sy0;
        {
            real mL = rowT[-1], mR = rowT[TW], mT = sT[lane + 2], mB = sT[(TILE_H + 1) * PT + lane + 2];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                real v;
                v = __shfl_xor_sync(FULL, mL, o); mL = v < mL ? v : mL;
                v = __shfl_xor_sync(FULL, mR, o); mR = v < mR ? v : mR;
                v = __shfl_xor_sync(FULL, mT, o); mT = v < mT ? v : mT;
                v = __shfl_xor_sync(FULL, mB, o); mB = v < mB ? v : mB;
            }
            sx0 = mR < mL ? -1 : 1;
            sy0 = mB < mT ? -1 : 1;
        }
        unsigned dirty = 0;                    // lane = row: cells of that row changed in this visit
        bool bad_cost = false;
        int steps = 0;
        // ---- round 0: the sweep that runs with the front, then the two-compare test of its downwind sides;
        //      further rounds: the other three sweeps (and the first again), Jacobi check passes to the fixed point ----
        bool more = true;
        for (int round = 0; more; ++round) {
            const int k0 = round == 1 ? 1 : 0, k1 = round == 0 ? 1 : 4;
#pragma unroll 1
            for (int k = k0; k < k1; ++k) {
                const int sx = (k & 1) ? -sx0 : sx0, sy = (k & 2) ? -sy0 : sy0;
                const unsigned d = wsweep_dir<real, STAGE_C>(sT, sC, cq, P.cost_pitch, x0, y0, P.rows, P.cols, sx, sy, lane, my_evals, bad_cost);
                const unsigned dflip = __shfl_sync(FULL, d, TILE_H - 1 - lane);
                dirty |= sy > 0 ? d : dflip;
                steps += NSTEP;
            }
            if (round == 0) {
                const int dv = sy0 > 0 ? PT : -PT;
                unsigned fm = 0;
#pragma unroll 1
                for (int cb = 0; cb < TW; cb += CB) {
                    const real *rT = rowT + cb;
                    real m[CB + 2], dn[CB];
#pragma unroll
                    for (int k = 0; k < CB; ++k) { m[k + 1] = rT[k]; dn[k] = rT[k + dv]; }
                    m[0] = rT[-1]; m[CB + 1] = rT[CB];
#pragma unroll
                    for (int k = 0; k < CB; ++k) {
                        const real cur = m[k + 1], nb = sx0 > 0 ? m[k + 2] : m[k];       // downwind-x neighbour (the halo at the end)
                        if (nb < cur || dn[k] < cur) fm |= 1u << (cb + k);
                    }
                }
                more = __any_sync(FULL, (fm & cmask) != 0);
                n_noop += !more && !__any_sync(FULL, dirty != 0);
                continue;
            }
            ++n_rounds;
            const long long tk0 = clock64();
            bool again = true;
            for (int pass = 0; pass < P.check_passes && again; ++pass) {
                bool changed = false;
#pragma unroll 1
                for (int cb = 0; cb < TW; cb += CB) {
                    const real *rT = rowT + cb;
                    real m[CB + 2], u[CB], dn[CB], cc[CB];
#pragma unroll
                    for (int k = 0; k < CB; ++k) {
                        m[k + 1] = rT[k]; u[k] = rT[k - PT]; dn[k] = rT[k + PT];
                        if (STAGE_C) cc[k] = sC[lane * PT + cb + k];
                        else cc[k] = ((cmask >> (cb + k)) & 1u) ? __ldg(&cq[(long long)(y0 + lane) * P.cost_pitch + x0 + cb + k]) : INF;
                    }
                    m[0] = rT[-1]; m[CB + 1] = rT[CB];
                    __syncwarp();                              // every load of this block precedes its stores (Jacobi)
                    unsigned dbits = 0;
#pragma unroll
                    for (int k = 0; k < CB; ++k) {
                        const real cur = m[k + 1];
                        const real a = m[k] < m[k + 2] ? m[k] : m[k + 2], b = u[k] < dn[k] ? u[k] : dn[k];
                        const real v = eikonal_update_sel<real>(a, b, cc[k]);
                        const bool want = (a < cur || b < cur) && cc[k] < INF;
                        bad_cost |= want && !(cc[k] >= cost_range<real>::lo && cc[k] <= cost_range<real>::hi);
                        if (want && v != cur && v <= num<real>::mul(cur, UP)) { rowT[cb + k] = v; dbits |= 1u << (cb + k); }
                    }
                    my_evals += CB;
                    dirty |= dbits;
                    changed |= dbits != 0;
                    __syncwarp();
                }
                steps += 4;
                again = __any_sync(FULL, changed);
            }
            c_check += clock64() - tk0;
            more = again;
            if (steps > P.step_cap) break;
        }
        if (__any_sync(FULL, bad_cost) && lane == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_COSTRANGE);
        n_steps += steps;
        ++n_visits;
        if (steps > P.step_cap) {
            if (lane == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_STEPCAP);
            break;
        }
        const long long tc3 = clock64();

        // ---- write back changed cells (dirty rows only, coalesced) ----
        {
            unsigned rows_dirty = __ballot_sync(FULL, dirty != 0);
            while (rows_dirty) {
                const int j = __ffs(rows_dirty) - 1;
                rows_dirty &= rows_dirty - 1;
                const unsigned dj = __shfl_sync(FULL, dirty, j);
                if ((dj >> lane) & 1u) st_T(&Tq[(long long)(y0 + j) * P.T_pitch + x0 + lane], sT[(j + 1) * PT + lane + 2]);
                n_written += __popc(dj);
            }
        }
        // ---- publish: per edge the lowest changed value that undercuts the halo; retire ----
        {
            const unsigned d_top = __shfl_sync(FULL, dirty, 0), d_bot = __shfl_sync(FULL, dirty, TILE_H - 1);
            real m0 = ((dirty & 1u) && rowT[0] < rowT[-1]) ? rowT[0] : INF;
            real m1 = (((dirty >> (TW - 1)) & 1u) && rowT[TW - 1] < rowT[TW]) ? rowT[TW - 1] : INF;
            real m2 = (((d_top >> lane) & 1u) && sT[PT + lane + 2] < sT[lane + 2]) ? sT[PT + lane + 2] : INF;
            real m3 = (((d_bot >> lane) & 1u) && sT[TILE_H * PT + lane + 2] < sT[(TILE_H + 1) * PT + lane + 2]) ? sT[TILE_H * PT + lane + 2] : INF;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                real v;
                v = __shfl_xor_sync(FULL, m0, o); m0 = v < m0 ? v : m0;
                v = __shfl_xor_sync(FULL, m1, o); m1 = v < m1 ? v : m1;
                v = __shfl_xor_sync(FULL, m2, o); m2 = v < m2 ? v : m2;
                v = __shfl_xor_sync(FULL, m3, o); m3 = v < m3 ? v : m3;
            }
            unsigned act = 0;
            if (m0 < INF && tx > 0) act |= 1u;
            if (m1 < INF && tx < P.ntx - 1) act |= 2u;
            if (m2 < INF && ty > 0) act |= 4u;
            if (m3 < INF && ty < P.nty - 1) act |= 8u;
            const real mine = lane == 0 ? m0 : lane == 1 ? m1 : lane == 2 ? m2 : m3;
            const unsigned long long pbits = (unsigned long long)__double_as_longlong((double)mine);
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
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_xor_sync(FULL, my_evals, o);
    if (lane == 0) {
        atomicAdd(&P.q.ctl->evals, (unsigned long long)my_evals);
        atomicAdd(&P.q.ctl->cells_written, n_written);
        atomicAdd(&P.q.ctl->cyc_wait, (unsigned long long)c_wait);
        atomicAdd(&P.q.ctl->cyc_load, (unsigned long long)c_load);
        atomicAdd(&P.q.ctl->cyc_relax, (unsigned long long)c_relax);
        atomicAdd(&P.q.ctl->cyc_store, (unsigned long long)c_store);
        atomicAdd(&P.q.ctl->visits, n_visits);
        atomicAdd(&P.q.ctl->steps, n_steps);
        atomicAdd(&P.q.ctl->pushes, n_pushes);
        if (n_defer) atomicAdd(&P.q.ctl->pad[0], n_defer);
        atomicAdd(&P.q.ctl->pad[1], (unsigned long long)c_check);
        atomicAdd(&P.q.ctl->noop_visits, n_noop);
        atomicAdd(&P.q.ctl->rounds, n_rounds);
    }
}

}  // namespace fmb
