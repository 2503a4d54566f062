// eikonal2d_cta.cuh -- 2D Eikonal solve, one CTA per tile visit (replaces FastMarching.py:17-29,44-112).
//
// Round-1's engine gave a tile to ONE warp (lane = row, one armed cell per lane per lock-step
// iteration): a 32x32 tile needed ~130 iterations although its front crosses it in 32-64, and the
// chain of ~200 dependent tile visits of a 4096^2 single-source solve ran at ~60 us per link.
// Here a tile visit is run by a whole CTA:
//
//   * every thread owns R horizontally adjacent cells (R = 1, 2, 4; 1024/R threads); a warp covers a
//     patch of 8R x 4 cells, so a front line crossing the tile keeps a handful of warps busy and the
//     others parked at the barrier;
//   * one lock-step iteration relaxes EVERY armed cell of the tile (in-place chaotic relaxation: a
//     neighbour value read in an iteration is the old or the new one, both are upper bounds of the
//     fixed point) and ends with ONE barrier that also carries the convergence vote
//     (__syncthreads_or): the number of iterations of a visit is the length of the longest
//     dependency chain inside the tile, not the number of cells per lane;
//   * a cell is armed by DATA: it is re-evaluated when the pair (min of its horizontal neighbours,
//     min of its vertical neighbours) differs from the pair it was last evaluated with -- no mask
//     exchange between warps.  Per-patch "touched" flags (double-buffered by iteration parity)
//     let the warps whose neighbourhood did not change skip the iteration with one shared load;
//   * values live in registers (own cells, their costs) and in one 34 x 40 shared array (tile + halo;
//     pitch 40 keeps the 8 x 4 lane patches free of bank conflicts).
//
// Scheduling (ring of active tiles, tile state machine, windowed / best-first orders, causal
// activation of neighbour tiles) is the protocol of fm_common.cuh / eikonal2d.cuh, driven by warp 0.
// The update, the acceptance rule (lower, or at most a few ulp higher) and the epsilon = 0 convergence
// test are those of eikonal2d.cuh, so the field is the same exact fixed point.
#pragma once
#include "eikonal2d.cuh"

namespace fmb {

template <typename real>
struct CtaTile2D {
    static constexpr int TW = 32, TH = 32;
    static constexpr int PT = 40;                       // smem row pitch: 40 % 16 == 8 -> the two rows of a half warp hit disjoint banks
    static constexpr int T_ELEMS = (TH + 2) * PT;       // row j (-1..32) at (j+1)*PT; column i (-1..32) at +i+2 (interior 16-byte aligned)
    static constexpr int EDGE_ELEMS = 4 * 32;           // per-edge candidate values [left, right, top, bottom][32]
    static constexpr int MAX_WARPS = 32;
    static constexpr size_t BYTES = sizeof(real) * (T_ELEMS + EDGE_ELEMS) + sizeof(int) * (2 * (MAX_WARPS + 1) + 16);
};

// Take the next tile for this CTA (executed by all 32 lanes of warp 0).  Returns the tile index
// (already RUNNING, fenced) or -1 when the solve is over / aborted.
template <typename real, bool BEST>
__device__ __forceinline__ int cta_acquire(const Problem2D<real> &P, int lane, int &streak, unsigned long long &n_defer,
                                           int *run_level = nullptr) {
    const int tiles_per_q = P.ntx * P.nty;
    const unsigned long long PRIO_INF = 0x7ff0000000000000ULL;
    int item;
    for (;;) {
        int it = -1;
        if (lane == 0) it = q_pop_lane0(P.q);
        item = __shfl_sync(FULL, it, 0);
        if (item < 0 || BEST || !P.windowed) break;
        if (P.windowed == 2) {
            // local causal order: the tile waits while a neighbour tile that is queued or running carries a LOWER
            // priority (= lowest value that activated it): that neighbour is upwind and will still lower this tile's
            // halo, so running now would only be repeated.  The active tile with the lowest priority never waits.
            // (A scheduling heuristic on racy reads: the field never depends on it.)
            const int tt = item % tiles_per_q;
            const int tty = tt / P.ntx, ttx = tt - tty * P.ntx;
            const unsigned long long mine = *reinterpret_cast<const volatile unsigned long long *>(&P.tile_prio[item]);
            const double slack = *reinterpret_cast<const volatile double *>(P.slack);
            bool blocked = false;
            if (lane < 12) {
                // lanes 0-3: the four face neighbours (wait rule above).  lanes 4-11: the second ring (diagonals and the
                // tiles two steps away in a straight line): an upwind tile there has not activated the face neighbour
                // between us yet -- that neighbour is still IDLE and invisible to the rule above -- but will, and the
                // face neighbour will then lower my halo: running now would be repeated.  Such a tile only counts when
                // its priority lies more than `hop` (one tile crossing, a fraction of it) below mine.
                const int ddx[12] = {-1, 1, 0, 0, -1, 1, -1, 1, -2, 2, 0, 0};
                const int ddy[12] = {0, 0, -1, 1, -1, -1, 1, 1, 0, 0, -2, 2};
                const int nx = ttx + ddx[lane], ny = tty + ddy[lane];
                const bool ring2 = lane >= 4;
                if (nx >= 0 && nx < P.ntx && ny >= 0 && ny < P.nty && !(ring2 && P.hop_frac <= 0.0)) {
                    const int n = item + ddy[lane] * P.ntx + ddx[lane];
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
            continue;
        }
        // windowed order (eikonal2d.cuh): run the tile only if it lies within win_window levels of the
        // lowest queued level, otherwise put it back at the tail
        int defer = 0;
        if (P.win_running) {
            // levels count queued AND running tiles: the tile runs only when nothing at a level more than
            // win_window below it is still queued or running (its upwind tiles have delivered their halos)
            if (lane == 0) {
                __threadfence();
                const int L = ld_volatile(&P.tile_level[item]);
                const int h0 = ld_volatile(P.win_hint);
                int h = h0;
                const int stop_at = min(L - P.win_window, h0 + 64);
                while (h < stop_at && ld_volatile(&P.lev_count[h]) <= 0) ++h;
                if (h != h0) atomicCAS(P.win_hint, h0, h);
                if (h < L - P.win_window && h < stop_at && streak < 100000) {
                    q_push(P.q, item);
                    defer = 1;
                } else if (run_level) *run_level = L;
            }
        } else if (lane == 0 && streak < 32) {
            __threadfence();
            const int L = ld_volatile(&P.tile_level[item]);
            atomicSub(&P.lev_count[L], 1);
            const int h0 = ld_volatile(P.win_hint);
            int h = h0;
            const int stop_at = min(L - P.win_window, h0 + 64);
            while (h < stop_at && ld_volatile(&P.lev_count[h]) <= 0) ++h;
            if (h != h0) atomicCAS(P.win_hint, h0, h);
            if (h < L - P.win_window && h < stop_at) {
                atomicAdd(&P.lev_count[L], 1);
                q_push(P.q, item);
                defer = 1;
            }
        } else if (lane == 0) atomicSub(&P.lev_count[ld_volatile(&P.tile_level[item])], 1);
        defer = __shfl_sync(FULL, defer, 0);
        if (!defer) { streak = 0; break; }
        ++streak;
        n_defer += 1;
        __nanosleep(streak < 4 ? 200u : streak < 12 ? 800u : 2000u);
    }
    if (item < 0) return -1;
    if (BEST) {
        // `item` is a query id: claim its lowest-priority queued tile (QUEUED -> RUNNING)
        const int base = item * tiles_per_q;
        const long long t0 = clock64();
        int claimed = -1;
        while (claimed < 0) {
            unsigned long long bp = ~0ULL;
            int bt = -1;
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
        if (claimed < 0) return -1;
        item = claimed;
        if (lane == 0) { atomicExch(&P.tile_prio[item], PRIO_INF); __threadfence(); }
    } else {
        // QUEUED -> RUNNING *before* T is sampled: anything published after this point flips the state to
        // DIRTY and the tile is run again.
        if (lane == 0) {
            if (P.windowed == 2) {
                const unsigned long long key = atomicExch(&P.tile_prio[item], PRIO_INF);
                *reinterpret_cast<volatile unsigned long long *>(&P.run_prio[item]) = key;
            }
            atomicExch(&P.tile_state[item], ST_RUNNING);
            if (P.windowed == 1) atomicExch(&P.tile_prio[item], PRIO_INF);
            __threadfence();
        }
    }
    __syncwarp();
    return item;
}

template <typename real, int R, bool BEST>
__global__ void __launch_bounds__(1024 / R) solve2d_cta_kernel(Problem2D<real> P) {
    using TL = CtaTile2D<real>;
    constexpr int PT = TL::PT;
    constexpr int NW = 32 / R;                  // warps per CTA == patches per tile
    constexpr int NPX = 4 / R;                  // patches across the tile (each 8R columns x 4 rows)
    static_assert(R == 1 || R == 2 || R == 4, "R must be 1, 2 or 4");
    FMB_DYN_SMEM(smem_raw);
    real *sT = reinterpret_cast<real *>(smem_raw);
    real *sEdge = sT + TL::T_ELEMS;
    int *sFlag = reinterpret_cast<int *>(sEdge + TL::EDGE_ELEMS);      // [2][NW + 1], slot NW = dummy target
    int *sCtl = sFlag + 2 * (TL::MAX_WARPS + 1);                       // [0] tile, [1] stop
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = lane & 7, py = lane >> 3;
    const int wx = warp % NPX, wy = warp / NPX;
    const int r = wy * 4 + py;                  // tile row of this thread's cells
    const int c0 = (wx * 8 + px) * R;           // first tile column
    real *cell = sT + (r + 1) * PT + c0 + 2;
    const real INF = num<real>::inf();
    const unsigned long long PRIO_INF = 0x7ff0000000000000ULL;
    const int tiles_per_q = P.ntx * P.nty;
    // patches whose cells read this thread's cells
    const int f_left = (px == 0 && wx > 0) ? warp - 1 : NW;
    const int f_right = (px == 7 && wx < NPX - 1) ? warp + 1 : NW;
    const int f_up = (py == 0 && wy > 0) ? warp - NPX : NW;
    const int f_down = (py == 3 && wy < 7) ? warp + NPX : NW;

    unsigned long long n_visits = 0, n_steps = 0, n_evals = 0, n_pushes = 0, n_written = 0, n_defer = 0;
    int streak = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0;

    for (;;) {
        const long long tc0 = clock64();
        if (warp == 0) {
            const int it = cta_acquire<real, BEST>(P, lane, streak, n_defer);
            if (lane == 0) sCtl[0] = it;
        }
        for (int i = tid; i < 2 * (TL::MAX_WARPS + 1); i += 1024 / R) sFlag[i] = 0;
        __syncthreads();
        const int item = sCtl[0];
        if (item < 0) break;
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        const int t = item - q * tiles_per_q;
        const int ty = t / P.ntx, tx = t - ty * P.ntx;
        const int x0 = tx * TL::TW, y0 = ty * TL::TH;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- stage: own cells straight into registers (+ shared copy), halo ring by the first 128 threads ----
        real cur[R], cst[R], ap[R], bp[R];
        unsigned live = 0;
        {
            const int y = y0 + r;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int x = x0 + c0 + s;
                const bool in = y < P.rows && x < P.cols;
                cst[s] = in ? __ldg(&cq[(long long)y * P.cost_pitch + x]) : INF;
                cur[s] = in ? ld_T(&Tq[(long long)y * P.T_pitch + x]) : INF;
            }
            if (tid < 128) {
                const int e = tid >> 5, k = tid & 31;
                int yy, xx, si;
                if (e == 0) { yy = y0 + k; xx = x0 - 1; si = (k + 1) * PT + 1; }
                else if (e == 1) { yy = y0 + k; xx = x0 + TL::TW; si = (k + 1) * PT + TL::TW + 2; }
                else if (e == 2) { yy = y0 - 1; xx = x0 + k; si = k + 2; }
                else { yy = y0 + TL::TH; xx = x0 + k; si = (TL::TH + 1) * PT + k + 2; }
                real v = INF;
                if (yy >= 0 && yy < P.rows && xx >= 0 && xx < P.cols) v = ld_T(&Tq[(long long)yy * P.T_pitch + xx]);
                sT[si] = v;
            }
#pragma unroll
            for (int s = 0; s < R; ++s) {
                cell[s] = cur[s];
                if (cst[s] < INF) live |= 1u << s;
            }
        }
        __syncthreads();

        // ---- what is armed at entry: cells next to a lower halo value, the neighbours of a source ----
        // (the interior of a tile is a fixed point for the halo of its previous visit)
        unsigned force = 0;
        {
            const real L = cell[-1], Rr = cell[R];
            const int lx = P.seeds[2 * q] - x0, ly = P.seeds[2 * q + 1] - y0;
            const bool arm_all = ((P.arm_rows & 1) && ty == 0) || ((P.arm_rows & 2) && ty == P.nty - 1);
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const real U = cell[s - PT], D = cell[s + PT];
                const real l = s == 0 ? L : cur[s - 1], rr = s == R - 1 ? Rr : cur[s + 1];
                ap[s] = l < rr ? l : rr;
                bp[s] = U < D ? U : D;
                const int c = c0 + s;
                bool f = (r == 0 && U < cur[s]) || (r == TL::TH - 1 && D < cur[s]) || (c == 0 && l < cur[s]) ||
                         (c == TL::TW - 1 && rr < cur[s]);
                const int dx = c - lx, dy = r - ly;
                f = f || (abs(dx) + abs(dy) == 1) || arm_all;
                if (f) force |= 1u << s;
            }
            force &= live;
        }

        // ---- relax to the fixed point: every armed cell, one barrier per iteration ----
        const long long tc2 = clock64();
        unsigned dirty = 0;
        int steps = 0;
        for (;;) {
            int *fl_rd = sFlag + (steps & 1) * (TL::MAX_WARPS + 1);
            int *fl_wr = sFlag + ((steps & 1) ^ 1) * (TL::MAX_WARPS + 1);
            bool changed = false;
            const int touched = fl_rd[warp] | (steps == 0);
            __syncwarp();
            if (touched) {                                  // warp-uniform
                if (lane == 0) fl_rd[warp] = 0;
                const real L = cell[-1], Rr = cell[R];
                real a[R], b[R];
                unsigned armed = 0;
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    const real U = cell[s - PT], D = cell[s + PT];
                    const real l = s == 0 ? L : cur[s - 1], rr = s == R - 1 ? Rr : cur[s + 1];
                    a[s] = l < rr ? l : rr;
                    b[s] = U < D ? U : D;
                    // armed at entry, or an input that changed lies below the cell (only a lower neighbour can
                    // lower a cell; the update from unchanged inputs is the value the cell already weighed)
                    if (((force >> s) & 1u) || (a[s] != ap[s] && a[s] < cur[s]) || (b[s] != bp[s] && b[s] < cur[s])) armed |= 1u << s;
                    ap[s] = a[s];
                    bp[s] = b[s];
                }
                force = 0;
                armed &= live;
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    n_evals += __popc(__ballot_sync(FULL, (armed >> s) & 1u));
                    if ((armed >> s) & 1u) {
                        const real v = eikonal_update<real>(a[s], b[s], cst[s]);
                        // lower wins; a value at most a few ulp HIGHER also replaces the stored one, so that the
                        // end state is a fixed point of the rounded update (eikonal2d.cuh)
                        if (v != cur[s] && v <= num<real>::mul(cur[s], (real)(1.0 + 8.0 / 4503599627370496.0))) {
                            cur[s] = v;
                            cell[s] = v;
                            dirty |= 1u << s;
                            changed = true;
                        }
                    }
                }
                if (changed) {
                    fl_wr[warp] = 1;
                    fl_wr[f_left] = 1;
                    fl_wr[f_right] = 1;
                    fl_wr[f_up] = 1;
                    fl_wr[f_down] = 1;
                }
            }
            ++steps;
            if (!__syncthreads_or(changed)) break;
            if (steps > P.step_cap) break;
        }
        n_steps += steps;
        ++n_visits;
        if (steps > P.step_cap) {
            if (tid == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_STEPCAP);
            break;
        }
        const long long tc3 = clock64();

        // ---- write back the changed cells; per edge, the lowest changed value that undercuts the halo ----
        {
            const int y = y0 + r;
#pragma unroll
            for (int s = 0; s < R; ++s)
                if ((dirty >> s) & 1u) st_T(&Tq[(long long)y * P.T_pitch + x0 + c0 + s], cur[s]);
            n_written += __popc(dirty);
            if (c0 == 0) sEdge[r] = ((dirty & 1u) && cur[0] < cell[-1]) ? cur[0] : INF;
            if (c0 + R == TL::TW) sEdge[32 + r] = (((dirty >> (R - 1)) & 1u) && cur[R - 1] < cell[R]) ? cur[R - 1] : INF;
            if (r == 0) {
#pragma unroll
                for (int s = 0; s < R; ++s) sEdge[64 + c0 + s] = (((dirty >> s) & 1u) && cur[s] < cell[s - PT]) ? cur[s] : INF;
            }
            if (r == TL::TH - 1) {
#pragma unroll
                for (int s = 0; s < R; ++s) sEdge[96 + c0 + s] = (((dirty >> s) & 1u) && cur[s] < cell[s + PT]) ? cur[s] : INF;
            }
        }
        __syncthreads();          // every T store of the CTA precedes warp 0's fence below

        // ---- publish: activate the neighbours that can still improve, retire (or requeue) this tile ----
        if (warp == 0) {
            real m0 = sEdge[lane], m1 = sEdge[32 + lane], m2 = sEdge[64 + lane], m3 = sEdge[96 + lane];
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
            __threadfence();          // the CTA's T stores (+ pending, priorities) are device-visible ...
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
            if (lane == 0) {
                const int drop = (nact - n_new) + (n_req ? 0 : 1);
                if (drop) atomicSub(&P.q.ctl->pending, drop);
                sCtl[1] = ld_volatile(&P.q.ctl->abort);
            }
        }
        __syncthreads();
        const int stop = sCtl[1];
        const long long tc4 = clock64();
        c_wait += tc1 - tc0; c_load += tc2 - tc1; c_relax += tc3 - tc2; c_store += tc4 - tc3;
        if (stop) break;
    }
    // counters: evaluations per warp, written cells per thread (warp-reduced), the rest by thread 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_written += __shfl_xor_sync(FULL, n_written, o);
    if (lane == 0) {
        atomicAdd(&P.q.ctl->evals, n_evals);
        atomicAdd(&P.q.ctl->cells_written, n_written);
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
    }
}


}  // namespace fmb
