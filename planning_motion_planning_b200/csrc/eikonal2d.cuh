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
    int handoff;             // 1: a finishing warp keeps one of the tiles it activated (skips the queue)
};

// FastMarching.py:17-29 getEikonal, written branch-for-branch on the values
// a = min(left,right), b = min(up,down).  Products and sums are individually
// rounded (no FMA contraction) so a cell relaxed from the same inputs gives the
// same bits as the reference.
template <typename real>
__device__ __forceinline__ real eikonal_update(real a, real b, real c) {
    using N = num<real>;
    real m = fmin(a, b);
    real d = N::sub(a, b);
    // one-sided when the other side is too far (or +inf): covers the reference's
    // isinf() branches and `cost < |Thor - Tver|`; (inf - inf) = NaN lands here too
    if (!(fabs(d) <= c)) return N::add(m, c);
    real disc = N::sub(N::mul((real)2, N::mul(c, c)), N::mul(d, d));
    return N::mul((real)0.5, N::add(N::add(a, b), N::sqrt(disc)));
}

template <typename real, int TW>
struct Tile2D {
    static constexpr int PT = TW + 2;                         // smem row pitch (even, PT-1 odd: conflict-free skews)
    static constexpr int T_ELEMS = (TILE_H + 2) * PT;
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
    for (long long i = tid; i < ntiles; i += nth) P.tile_state[i] = ST_IDLE;
    for (long long i = tid; i < ring_slots; i += nth) P.q.ring[i] = -1;
    if (tid == 0) {
        QueueCtl z = {};
        *P.q.ctl = z;
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
    int cand[5][2] = {{tx, ty}, {-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
    if (sx % TW == 0 && tx > 0) { cand[1][0] = tx - 1; cand[1][1] = ty; }
    if (sx % TW == TW - 1 && tx < P.ntx - 1) { cand[2][0] = tx + 1; cand[2][1] = ty; }
    if (sy % TILE_H == 0 && ty > 0) { cand[3][0] = tx; cand[3][1] = ty - 1; }
    if (sy % TILE_H == TILE_H - 1 && ty < P.nty - 1) { cand[4][0] = tx; cand[4][1] = ty + 1; }
    for (int k = 0; k < 5; ++k) {
        if (cand[k][0] < 0) continue;
        int item = base + cand[k][1] * P.ntx + cand[k][0];
        if (tile_activate(P.tile_state, P.q.ctl, item)) { q_push(P.q, item); atomicAdd(&P.q.ctl->pushes, 1ULL); }
    }
}

// ---------------------------------------------------------------------------
// persistent solver
template <typename real, int TW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) solve2d_kernel(Problem2D<real> P) {
    using TL = Tile2D<real, TW>;
    constexpr int PT = TL::PT;
    FMB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    real *sT = reinterpret_cast<real *>(smem_raw) + (size_t)warp * TL::WARP_ELEMS;
    real *sC = sT + TL::T_ELEMS;
    const real INF = num<real>::inf();
    const int tiles_per_q = P.ntx * P.nty;

    unsigned long long n_visits = 0, n_steps = 0, n_evals = 0, n_pushes = 0, n_written = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0;
    int item = -1;

    for (;;) {
        const long long tc0 = clock64();
        if (item < 0) {
            int it = -1;
            if (lane == 0) it = q_pop_lane0(P.q);
            item = __shfl_sync(FULL, it, 0);
            if (item < 0) break;
        }
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        const int t = item - q * tiles_per_q;
        const int ty = t / P.ntx, tx = t - ty * P.ntx;
        const int x0 = tx * TW, y0 = ty * TILE_H;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // QUEUED -> RUNNING *before* sampling T: anything published after this
        // point flips the state to DIRTY and the tile is run again.
        if (lane == 0) { atomicExch(&P.tile_state[item], ST_RUNNING); __threadfence(); }
        __syncwarp();

        // ---- stage tile + halo (rows -1..32 of TW cells, coalesced) ----------
        for (int idx = lane; idx < (TILE_H + 2) * TW; idx += 32) {
            const int j = idx / TW - 1, i = idx - (j + 1) * TW;
            const int y = y0 + j, x = x0 + i;
            real v = INF;
            if (y >= 0 && y < P.rows && x < P.cols) v = ld_T(&Tq[(long long)y * P.T_pitch + x]);
            sT[(j + 1) * PT + i + 1] = v;
        }
        {   // left / right halo columns (lane == row)
            const int y = y0 + lane;
            real vl = INF, vr = INF;
            if (y < P.rows) {
                if (x0 > 0) vl = ld_T(&Tq[(long long)y * P.T_pitch + x0 - 1]);
                if (x0 + TW < P.cols) vr = ld_T(&Tq[(long long)y * P.T_pitch + x0 + TW]);
            }
            sT[(lane + 1) * PT] = vl;
            sT[(lane + 1) * PT + TW + 1] = vr;
        }
        for (int idx = lane; idx < TILE_H * TW; idx += 32) {
            const int j = idx / TW, i = idx - j * TW;
            const int y = y0 + j, x = x0 + i;
            real c = INF;
            if (y < P.rows && x < P.cols) c = __ldg(&cq[(long long)y * P.cost_pitch + x]);
            sC[j * PT + i] = c;
        }
        __syncwarp();

        // ---- per-row masks ---------------------------------------------------
        real *rowT = sT + (lane + 1) * PT + 1;      // rowT[k] = T(row lane, col k); rowT[-1], rowT[TW] halos
        const real *rowC = sC + lane * PT;
        unsigned cmask = 0;                         // cells that can ever be relaxed (finite cost)
#pragma unroll 8
        for (int k = 0; k < TW; ++k) cmask |= (rowC[k] < INF ? 1u : 0u) << k;
        unsigned mask = 0;
        if (rowT[-1] < rowT[0]) mask |= 1u;
        if (rowT[TW] < rowT[TW - 1]) mask |= 1u << (TW - 1);
        {
            const bool in = lane < TW;
            unsigned bt = __ballot_sync(FULL, in && sT[lane + 1] < sT[PT + lane + 1]);
            unsigned bb = __ballot_sync(FULL, in && sT[(TILE_H + 1) * PT + lane + 1] < sT[TILE_H * PT + lane + 1]);
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
        mask &= cmask;

        // ---- relax to the fixed point ---------------------------------------
        const long long tc2 = clock64();
        unsigned dirty = 0;
        int last = 0, dir = 1, steps = 0;
        bool fail = false;
        unsigned active;
        while ((active = __ballot_sync(FULL, mask != 0)) != 0) {
            int k = -1;
            real v = INF, cur = INF, l = INF, r = INF, u = INF, d = INF;
            if (mask) {
                // continue in the current direction along the row, turn round at the end
                const unsigned hi = mask & (~0u << last);
                const unsigned lo = mask & ((2u << last) - 1u);
                if (dir > 0) {
                    if (hi) k = __ffs(hi) - 1; else { k = 31 - __clz(lo); dir = -1; }
                } else {
                    if (lo) k = 31 - __clz(lo); else { k = __ffs(hi) - 1; dir = 1; }
                }
                last = k;
                mask &= ~(1u << k);
                const real *p = rowT + k;
                l = p[-1]; r = p[1]; u = p[-PT]; d = p[PT]; cur = p[0];
                v = eikonal_update<real>(fmin(l, r), fmin(u, d), rowC[k]);
            }
            __syncwarp();                     // every lane has read before anyone writes (Jacobi step)
            int up_msg = -1, dn_msg = -1;
            if (k >= 0 && v < cur) {
                rowT[k] = v;
                dirty |= 1u << k;
                if (k > 0 && l > v) mask |= 1u << (k - 1);
                if (k < TW - 1 && r > v) mask |= 1u << (k + 1);
                if (u > v) up_msg = k;
                if (d > v) dn_msg = k;
            }
            const int from_below = __shfl_down_sync(FULL, up_msg, 1);   // lane+1 asks me to re-check column
            const int from_above = __shfl_up_sync(FULL, dn_msg, 1);
            if (lane < TILE_H - 1 && from_below >= 0) mask |= 1u << from_below;
            if (lane > 0 && from_above >= 0) mask |= 1u << from_above;
            mask &= cmask;
            __syncwarp();                     // writes visible to the next iteration's reads
            n_evals += __popc(active);
            if (++steps > P.step_cap) { fail = true; break; }
        }
        n_steps += steps;
        ++n_visits;
        if (fail) {
            if (lane == 0) atomicCAS(&P.q.ctl->abort, 0, DEV_STEPCAP);
            break;
        }
        const long long tc3 = clock64();

        // ---- write back changed cells (row by row, coalesced) ----------------
        for (int j = 0; j < TILE_H; ++j) {
            const unsigned dj = __shfl_sync(FULL, dirty, j);
            if (dj == 0) continue;
            if (lane < TW && ((dj >> lane) & 1u)) {
                st_T(&Tq[(long long)(y0 + j) * P.T_pitch + x0 + lane], sT[(j + 1) * PT + lane + 1]);
            }
            n_written += __popc(dj);
        }
        // ---- which neighbours can still improve? ----------------------------
        // only an edge cell that changed in this visit AND undercuts the value
        // across the edge can lower anything in the neighbour (causality).
        const bool nl = (dirty & 1u) && rowT[0] < rowT[-1];
        const bool nr = ((dirty >> (TW - 1)) & 1u) && rowT[TW - 1] < rowT[TW];
        const unsigned d_top = __shfl_sync(FULL, dirty, 0), d_bot = __shfl_sync(FULL, dirty, TILE_H - 1);
        const bool nt = lane < TW && ((d_top >> lane) & 1u) && sT[PT + lane + 1] < sT[lane + 1];
        const bool nb = lane < TW && ((d_bot >> lane) & 1u) && sT[TILE_H * PT + lane + 1] < sT[(TILE_H + 1) * PT + lane + 1];
        const bool actL = __any_sync(FULL, nl) && tx > 0;
        const bool actR = __any_sync(FULL, nr) && tx < P.ntx - 1;
        const bool actT = __any_sync(FULL, nt) && ty > 0;
        const bool actB = __any_sync(FULL, nb) && ty < P.nty - 1;
        __threadfence();          // my T stores are device-visible ...
        __syncwarp();             // ... before lane 0 publishes the activations
        int next = -1;
        if (lane == 0) {
            __threadfence();
            const int nbr[4] = {item - 1, item + 1, item - P.ntx, item + P.ntx};
            const bool act[4] = {actL, actR, actT, actB};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (!act[s]) continue;
                if (tile_activate(P.tile_state, P.q.ctl, nbr[s])) {
                    if (next < 0 && P.handoff) next = nbr[s];     // keep one for myself: no queue round trip
                    else { q_push(P.q, nbr[s]); ++n_pushes; }
                }
            }
            if (tile_finish(P.tile_state, P.q.ctl, item)) {
                if (next < 0 && P.handoff) next = item;
                else { q_push(P.q, item); ++n_pushes; }
            }
            if (ld_volatile(&P.q.ctl->abort)) next = -2;     // somebody failed: leave (warp-uniform via shfl)
        }
        item = __shfl_sync(FULL, next, 0);
        const long long tc4 = clock64();
        c_wait += tc1 - tc0; c_load += tc2 - tc1; c_relax += tc3 - tc2; c_store += tc4 - tc3;
        if (item == -2) break;
    }
    // per-warp counters (lane 0 holds the per-warp ones; evals/steps are warp-uniform)
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
