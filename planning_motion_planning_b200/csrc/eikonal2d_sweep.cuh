// eikonal2d_sweep.cuh -- 2D Eikonal solve, sweep engine (replaces FastMarching.py:17-29,44-112).
//
// One tile visit = one CTA of 4 warps; each warp runs ONE of the four diagonal-wavefront Gauss-Seidel sweeps
// (+x+y, -x+y, +x-y, -x-y) over the same 32 x 32 shared tile at the same time.
//
// Why sweeps.  The Jacobi visit of eikonal2d_cta.cuh needs as many iterations as the longest dependency chain
// but re-evaluates a cell every time one of its inputs ripples (measured: 14 evaluations per cell per visit,
// every warp busy in every iteration, issue bound at ~1400 cycles per iteration); round 1's warp-per-tile visit
// relaxed one cell per lane per step (130 steps of ~570 cycles).  A wavefront sweep relaxes the cells of one
// anti-diagonal per step in an order in which both sweep-upwind neighbours of a cell were relaxed one step
// earlier by the same warp: a front that crosses the tile within one quadrant of directions is final after ONE
// sweep of 63 steps with one evaluation per cell, and the step costs one dependent update chain of a single warp:
// no CTA barrier inside a sweep, lane = row, the upwind row neighbour is the lane's own previous result, the
// upwind column neighbour comes from lane - 1 by one shuffle, everything else is loaded a step ahead.  The
// update is evaluated without branches (eikonal_update_sel, exact branch-free sqrt).  The three sweeps that run
// against the front find no cell with a lower sweep-upwind neighbour and skip their steps after one vote.
//
// After the four sweeps, check passes relax every cell Jacobi-style (128 threads x 8 cells); the visit is over when
// a pass changes nothing -- the same epsilon = 0 fixed point as the other engines -- otherwise the sweeps repeat.
// Two warps may relax one cell in the same instant and the later store may carry the higher value: the check pass
// sees any cell that is not at its fixed point, so a lost update costs a round, never correctness.
//
// Hand-off (one map, local causal order).  A single-source solve is a CHAIN of dependent tile visits, so what counts
// is when the next tile can start:
//   * early publish: after the first round of sweeps the changed cells are written back and the neighbours are
//     activated BEFORE the check passes run (which usually change nothing); the tile keeps running, stops blocking
//     its neighbours in the causal order (run_prio = maximum) and publishes again only if a later pass or round
//     changes an edge;
//   * continuation in place: when a visit ends and the tile was re-activated meanwhile (DIRTY) the CTA does not
//     requeue it but reloads the halo ring and goes on with a check pass, which is also the comparison "did my
//     inputs change?".
// (Measured and dropped: streaming the halo through global memory while both tiles sweep, with activation of the
// downwind tiles at step 34 -- twice the visits per tile, slower.)
//
// Staging (Blackwell data movement).  An interior tile is staged by TWO TMA tensor copies issued by one thread:
// cp.async.bulk.tensor.3d of the 38 x 34 T box at (x0 - 2, y0 - 1) -- the halo columns and rows arrive with the tile,
// which is why the shared T layout has a row pitch of 38 -- and of the 34 x 32 cost box at (x0, y0), both completing
// on one mbarrier (complete_tx::bytes) that the CTA waits on by parity.  TMA fills out-of-bounds elements with zeros,
// not +inf, so tiles that touch the map limits (and fp32 / unaligned fields) keep the cp.async / register paths.
#pragma once
#include "eikonal2d_cta.cuh"
#ifndef FMB_HOST_EMU
#include <cuda.h>          // CUtensorMap
#endif

namespace fmb {

// shared-memory layout of one sweep-engine CTA (elements of `real` unless noted)
struct Sweep2DSmem {
    // T row pitch: [+1] left halo, [+2 .. +33] interior, [+34] right halo, [+35 .. +37] padding that the TMA box fills
    // with the next columns.  38: PT - 1 odd keeps the skewed accesses of the sweeps (lane l, column d - l) free of
    // bank conflicts, PT mod 16 = 6 keeps the column accesses of the check passes (lane = row) at two-way (36 is four-way:
    // measured +3.5 % per sweep step), PT * 8 is a multiple of 16 bytes (TMA box rows, cp.async chunks).
    static constexpr int PT = 38;
    static constexpr int PC = 34;                        // cost row pitch
    static constexpr int T_ELEMS = 1296;                 // 34 rows x 38, padded so that the cost box starts 128-byte aligned
    static constexpr int C_ELEMS = TILE_H * PC;
    template <typename real> static constexpr size_t bytes() { return sizeof(real) * (T_ELEMS + C_ELEMS) + 32 * 4 + 4 * 4 + 16; }
};

#ifndef FMB_HOST_EMU
struct TmaMaps2D { CUtensorMap T, C; };
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
struct TmaMaps2D { int unused; };
#endif

#ifndef FMB_HOST_EMU
#define FMB_TMA_PARAM , const __grid_constant__ TmaMaps2D tmaps, int use_tma
#else
#define FMB_TMA_PARAM , TmaMaps2D tmaps = TmaMaps2D(), int use_tma = 0
#endif
// resident CTAs per SM the register allocation is held to (3: 168 registers, 4: 128)
#ifndef FMB_SWEEP2D_MINB
#define FMB_SWEEP2D_MINB 3
#endif
template <typename real, bool BEST>
__global__ void __launch_bounds__(128, FMB_SWEEP2D_MINB) solve2d_sweep_kernel(Problem2D<real> P FMB_TMA_PARAM) {
    constexpr int PT = Sweep2DSmem::PT, PC = Sweep2DSmem::PC, TW = 32, NSTEP = TILE_H + TW - 1;
#ifndef FMB_HOST_EMU
    extern __shared__ __align__(1024) unsigned char smem_raw[];
#else
    FMB_DYN_SMEM(smem_raw);
#endif
    real *sT = reinterpret_cast<real *>(smem_raw);
    real *sC = sT + Sweep2DSmem::T_ELEMS;
    unsigned *sDirty = reinterpret_cast<unsigned *>(sC + Sweep2DSmem::C_ELEMS);      // [32] changed cells per row since the last write-back
    int *sCtl = reinterpret_cast<int *>(sDirty + 32);                        // [0] tile, [1] stop, [2] level, [3] continue in place
    unsigned long long *sBar = reinterpret_cast<unsigned long long *>(sCtl + 4);      // mbarrier of the TMA copies
    unsigned bar_phase = 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const real INF = num<real>::inf();
    const int tiles_per_q = P.ntx * P.nty;
    const real UP = (real)(1.0 + 8.0 / 4503599627370496.0);
    const bool early_publish = !BEST && P.windowed == 2 && P.pipeline;

    unsigned long long n_visits = 0, n_steps = 0, n_pushes = 0, n_defer = 0, n_noop = 0, n_rounds = 0, n_cont = 0;
    unsigned my_evals = 0, my_written = 0;
    int streak = 0;
    long long c_wait = 0, c_load = 0, c_relax = 0, c_store = 0, c_check = 0;
#ifndef FMB_HOST_EMU
    if (use_tma && tid == 0) { mbar_init(sBar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
#endif

    for (;;) {
        const long long tc0 = clock64();
        if (warp == 0) {
            int lv = 0;
            const int it = cta_acquire<real, BEST>(P, lane, streak, n_defer, &lv);
            if (lane == 0) { sCtl[0] = it; sCtl[2] = lv; }
        }
        if (tid < 32) sDirty[tid] = 0;
        __syncthreads();
        const int item = sCtl[0];
        if (item < 0) break;
        const long long tc1 = clock64();
        const int q = item / tiles_per_q;
        const int t = item - q * tiles_per_q;
        const int ty = t / P.ntx, tx = t - ty * P.ntx;
        const int x0 = tx * TW, y0 = ty * TILE_H;
        const real *cq = P.cost + (long long)q * P.cost_qstride;
        real *Tq = P.T + (long long)q * P.T_qstride;

        // ---- the cost map may still be arriving (fmb_solve2d_h2d_f64): wait for the band of rows this tile lies in ----
        if (P.band_ready) {
            if (tid == 0) {
                int gone = 0;
                const long long t0 = clock64();
                while (ld_acquire_sys(&P.band_ready[y0 >> P.band_shift]) == 0) {
                    if (ld_volatile(&P.q.ctl->abort)) { gone = 1; break; }
                    if (clock64() - t0 > P.q.watchdog_cycles) { atomicCAS(&P.q.ctl->abort, 0, DEV_WATCHDOG); gone = 1; break; }
                    __nanosleep(200);
                }
                sCtl[1] = gone;
            }
            __syncthreads();
            if (sCtl[1]) break;                    // aborted while waiting
        }

        // ---- stage tile + halo (cp.async.cg for interior aligned tiles, bounds-checked loads otherwise) ----
        {
            constexpr int EPC = 16 / (int)sizeof(real);
            constexpr int CPR = TW / EPC;
            const bool fast = sizeof(real) == 8 && y0 >= 1 && y0 + TILE_H < P.rows && x0 + TW <= P.cols &&
                              (P.T_pitch % EPC) == 0 && (P.cost_pitch % EPC) == 0 &&
                              ((size_t)Tq % 16) == 0 && ((size_t)cq % 16) == 0;
            bool by_tma = false;
#ifndef FMB_HOST_EMU
            by_tma = use_tma && sizeof(real) == 8 && x0 >= 2 && x0 + PT - 2 <= P.cols && y0 >= 1 && y0 + TILE_H + 1 <= P.rows;
            if (by_tma) {
                // the previous visit's generic-proxy accesses to this shared memory are ordered before the async-proxy writes
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    mbar_expect_tx(sBar, (unsigned)(sizeof(real) * ((TILE_H + 2) * PT + TILE_H * PC)));
                    tma_load_3d(sT, &tmaps.T, x0 - 2, y0 - 1, q, sBar);
                    tma_load_3d(sC, &tmaps.C, x0, y0, P.cost_qstride ? q : 0, sBar);
                }
                mbar_wait(sBar, bar_phase);
                bar_phase ^= 1u;
            } else
#endif
            if (fast) {
                for (int c = tid; c < (TILE_H + 2) * CPR; c += 128) {
                    const int row = c / CPR, col = (c % CPR) * EPC;
                    cp_async16_cg(&sT[row * PT + 2 + col], &Tq[(long long)(y0 - 1 + row) * P.T_pitch + x0 + col]);
                }
                for (int c = tid; c < TILE_H * CPR; c += 128) {
                    const int row = c / CPR, col = (c % CPR) * EPC;
                    cp_async16_cg(&sC[row * PC + col], &cq[(long long)(y0 + row) * P.cost_pitch + x0 + col]);
                }
                if (tid < 64) {                                 // left / right halo columns
                    const int k = tid & 31, y = y0 + k;
                    const int x = tid < 32 ? x0 - 1 : x0 + TW;
                    real v = INF;
                    if (x >= 0 && x < P.cols) v = ld_T(&Tq[(long long)y * P.T_pitch + x]);
                    sT[(k + 1) * PT + (tid < 32 ? 1 : TW + 2)] = v;
                }
                cp_async_wait_all();
            } else {
                for (int idx = tid; idx < (TILE_H + 2) * (TW + 2); idx += 128) {
                    const int j = idx / (TW + 2) - 1, i = idx % (TW + 2) - 1;
                    const int yy = y0 + j, xx = x0 + i;
                    real v = INF;
                    if (yy >= 0 && yy < P.rows && xx >= 0 && xx < P.cols) v = ld_T(&Tq[(long long)yy * P.T_pitch + xx]);
                    sT[(j + 1) * PT + i + 2] = v;
                }
                for (int idx = tid; idx < TILE_H * TW; idx += 128) {
                    const int j = idx / TW, i = idx % TW;
                    const int yy = y0 + j, xx = x0 + i;
                    real c = INF;
                    if (yy < P.rows && xx < P.cols) c = __ldg(&cq[(long long)yy * P.cost_pitch + xx]);
                    sC[j * PC + i] = c;
                }
            }
        }
        __syncthreads();
        const long long tc2 = clock64();

        // ---- per-warp sweep geometry ----
        const int sx = (warp & 1) ? -1 : 1, sy = (warp & 2) ? -1 : 1;
        const int jrow = sy > 0 ? lane : TILE_H - 1 - lane;
        volatile real *rowT = sT + (jrow + 1) * PT + 2;
        const real *rowC = sC + jrow * PC;
        const int dv = sy > 0 ? PT : -PT;              // towards the sweep-downwind row
        const int hcol = sx > 0 ? -1 : TW;             // my upwind halo column (tile coordinates)
        int steps = 0, round = 0;
        bool spec = false;                 // this pass through the publish block is the early one
        long long t_rounds = 0;
        for (;;) {      // continuation loop: one pass per (re-)activation served in place
            const long long tr0 = clock64();
            for (;; ++round) {
                // ---- check passes: one Jacobi relaxation of every cell (thread = row `lane`, columns 8*warp .. +7, all
                // inputs loaded before the first store); converged when a pass changes nothing.  What a round leaves
                // behind is usually a few cells that settle within a few Jacobi steps, so up to check_passes passes run
                // before another round of sweeps is paid for.  (Round 0 of a fresh visit goes straight to the sweeps
                // unless precheck is set.)
                const long long tk0 = clock64();
                int again = 1;
                const int npass = round == 0 ? P.precheck : P.check_passes;
                for (int pass = 0; pass < npass; ++pass) {
                    bool changed = false;
                    const real *rT = sT + (lane + 1) * PT + 2 + 8 * warp;
                    const real *rC = sC + lane * PC + 8 * warp;
                    real m[10], u[8], dn[8], cc[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { m[k + 1] = rT[k]; u[k] = rT[k - PT]; dn[k] = rT[k + PT]; cc[k] = rC[k]; }
                    m[0] = rT[-1]; m[9] = rT[8];
                    unsigned dbits = 0;
                    bool bad_cost = false;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const real cur = m[k + 1];
                        const real a = m[k] < m[k + 2] ? m[k] : m[k + 2], b = u[k] < dn[k] ? u[k] : dn[k];
                        const real v = eikonal_update_sel<real>(a, b, cc[k]);
                        const bool ok = (a < cur || b < cur) && cc[k] < INF && v != cur && v <= num<real>::mul(cur, UP);
                        bad_cost |= cc[k] < INF && !(cc[k] >= cost_range<real>::lo && cc[k] <= cost_range<real>::hi);
                        if (ok) { sT[(lane + 1) * PT + 2 + 8 * warp + k] = v; dbits |= 1u << (8 * warp + k); }
                    }
                    my_evals += 8;
                    if (dbits) { atomicOr(&sDirty[lane], dbits); changed = true; }
                    if (bad_cost) atomicCAS(&P.q.ctl->abort, 0, DEV_COSTRANGE);      // outside the update's supported range
                    steps += 4;
                    again = __syncthreads_or(changed);
                    if (!again) break;
                }
                c_check += clock64() - tk0;
                if (!again) { n_noop += round == 0; break; }
                if (steps > P.step_cap) break;
                ++n_rounds;

                // ---- one sweep per warp.  Lane l relaxes cell (i, jrow) at step d, i = d - l counted along the sweep.
                int i = sx > 0 ? -lane : TW - 1 + lane;
                real res = rowT[hcol];                     // before my first cell: the halo column
                unsigned dirty = 0;
                bool hot = false;
                int ic = min(max(i, 0), TW - 1);
                real n_cur = rowT[ic], n_c = rowC[ic], n_dwh = rowT[ic + sx], n_dwv = rowT[ic + dv], n_up0 = rowT[ic - dv];
                if (P.variant & 1) {
                    // straight-line step: the update is evaluated in every step and its result applied by predication --
                    // no vote, no branch, no reconvergence point between one step's result and the next step's shuffle;
                    // the cell's value "as it is now" is loaded at the top of the step with the other operands
                    for (int d = 0; d < NSTEP; ++d, i += sx) {
                        const bool valid = (unsigned)i < (unsigned)TW;
                        const real cur = n_cur, c = n_c, dwh = n_dwh, dwv = n_dwv, up0 = n_up0;
                        const int iw = ic;
                        real now = rowT[iw];
                        ic = min(max(i + sx, 0), TW - 1);
                        n_cur = rowT[ic]; n_c = rowC[ic]; n_dwh = rowT[ic + sx]; n_dwv = rowT[ic + dv];
                        if (lane == 0) n_up0 = rowT[ic - dv];
                        real up = __shfl_up_sync(FULL, res, 1);
                        if (lane == 0) up = up0;
                        const bool go = valid && (res < cur || up < cur) && c < INF;
                        const real v = eikonal_update_sel<real>(res < dwh ? res : dwh, up < dwv ? up : dwv, c);
                        const bool acc = go && v != cur && v <= num<real>::mul(cur, UP);
                        if (P.variant & 2) now = rowT[iw];          // as late as possible: the window for a lost update shrinks
                        const bool st = acc && v != now && v <= num<real>::mul(now, UP);
                        if (st) rowT[iw] = v;
                        dirty |= st ? 1u << iw : 0u;
                        my_evals += go;
                        res = valid ? (acc ? v : cur) : res;
                    }
                } else
                for (int d = 0; d < NSTEP; ++d, i += sx) {
                    const bool valid = (unsigned)i < (unsigned)TW;
                    const real cur = n_cur, c = n_c, dwh = n_dwh, dwv = n_dwv, up0 = n_up0;
                    ic = min(max(i + sx, 0), TW - 1);
                    n_cur = rowT[ic]; n_c = rowC[ic]; n_dwh = rowT[ic + sx]; n_dwv = rowT[ic + dv];
                    if (lane == 0) n_up0 = rowT[ic - dv];
                    real up = __shfl_up_sync(FULL, res, 1);
                    if (lane == 0) up = up0;
                    // this sweep only offers something to a cell whose sweep-upwind neighbours lie below it.  The sweep
                    // that runs with the front has work in every step: while the previous step had any (`hot`), the
                    // update starts straight after the shuffle and the vote only prepares the next step; a sweep that
                    // runs against the front asks first and skips the step.
                    const bool go = valid && (res < cur || up < cur) && c < INF;
                    real out = cur;
                    if (hot || __any_sync(FULL, go)) {
                        const real v = eikonal_update_sel<real>(res < dwh ? res : dwh, up < dwv ? up : dwv, c);
                        my_evals += go;
                        if (go && v != cur && v <= num<real>::mul(cur, UP)) {
                            out = v;
                            // `cur` was loaded a step ago: weigh the value against the cell as it is NOW, so that a lower
                            // value another sweep stored meanwhile is not overwritten
                            const real now = rowT[i];
                            if (v != now && v <= num<real>::mul(now, UP)) {
                                rowT[i] = v;
                                dirty |= 1u << i;
                            }
                        }
                    }
                    hot = __any_sync(FULL, go);
                    if (valid) res = out;
                }
                if (dirty) atomicOr(&sDirty[jrow], dirty);
                __syncthreads();
                steps += NSTEP;
                if (early_publish && round == 0) { spec = true; break; }     // publish now, check afterwards
            }
            t_rounds += clock64() - tr0;
            if (steps > P.step_cap) break;

            // ---- write back changed cells (lane = column, coalesced rows) ----
#pragma unroll
            for (int rr = 0; rr < TILE_H / 4; ++rr) {
                const int j = warp + 4 * rr;
                const unsigned m = sDirty[j];
                if ((m >> lane) & 1u) {
                    st_T(&Tq[(long long)(y0 + j) * P.T_pitch + x0 + lane], sT[(j + 1) * PT + lane + 2]);
                    ++my_written;
                }
            }
            __syncthreads();          // every T store of the CTA precedes warp 0's fence below

            // ---- publish: per edge the lowest changed value that undercuts the halo; retire, or continue in place ----
            if (warp == 0) {
                const unsigned dl = sDirty[lane];
                const unsigned d_top = sDirty[0], d_bot = sDirty[TILE_H - 1];
                const real *rT = sT + (lane + 1) * PT + 2;
                real m0 = ((dl & 1u) && rT[0] < rT[-1]) ? rT[0] : INF;
                real m1 = (((dl >> (TW - 1)) & 1u) && rT[TW - 1] < rT[TW]) ? rT[TW - 1] : INF;
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
                __threadfence();          // the CTA's T stores (+ pending, priorities) are device-visible ...
                __syncwarp();             // ... before any state transition is published
                bool pushed = false, newly = false, cont = false;
                {
                    const bool is_nbr = lane < 4 && ((act >> lane) & 1u);
                    const bool is_self = lane == 4;
                    const int tgt = is_self ? item : item + (lane == 0 ? -1 : lane == 1 ? 1 : lane == 2 ? -P.ntx : P.ntx);
                    if (is_nbr || is_self) {
                        int *st = &P.tile_state[tgt];
                        int old = ST_RUNNING;
                        if (!(is_self && spec)) old = atomicCAS(st, is_self ? ST_RUNNING : ST_IDLE, is_self ? ST_IDLE : ST_QUEUED);
                        if (is_self) {
                            if (spec) {
                                // early publish: keep running, but no longer hold back the neighbours
                                *reinterpret_cast<volatile unsigned long long *>(&P.run_prio[item]) = ~0ULL;
                                cont = true;
                            } else if (old != ST_RUNNING) { atomicExch(st, ST_RUNNING); cont = true; }   // re-activated while it ran: serve that here
                        } else {
                            for (;;) {
                                if (old == ST_IDLE) { newly = true; break; }
                                if (old == ST_QUEUED || old == ST_DIRTY) break;
                                if (atomicCAS(st, ST_RUNNING, ST_DIRTY) == ST_RUNNING) break;          // ask the runner to look again
                                old = atomicCAS(st, ST_IDLE, ST_QUEUED);
                            }
                        }
                        if (newly) {
                            if (!BEST && P.windowed == 1) win_count_push<real>(P, tgt);
                            q_push(P.q, BEST ? q : tgt);
                            pushed = true;
                        }
                    }
                }
                const int n_new = __popc(__ballot_sync(FULL, newly));
                const int n_cnt = __popc(__ballot_sync(FULL, cont));
                n_pushes += __popc(__ballot_sync(FULL, pushed));
                if (lane == 0) {
                    const int drop = (nact - n_new) + (n_cnt ? 0 : 1);
                    if (drop) atomicSub(&P.q.ctl->pending, drop);
                    if (!n_cnt && !BEST && P.windowed == 1 && P.win_running) atomicSub(&P.lev_count[sCtl[2]], 1);
                    sCtl[1] = ld_volatile(&P.q.ctl->abort);
                    sCtl[3] = n_cnt;
                    if (n_cnt) __threadfence();        // RUNNING again before the halo is sampled
                }
            }
            __syncthreads();
            if (!sCtl[3] || sCtl[1]) break;
            if (tid < 32) sDirty[tid] = 0;                         // what was written back is published
            // ---- continue in place: fresh halo ring, then a check pass decides whether anything is left to do ----
            if (round == 0) round = 1;          // a continuation always opens with check passes
            if (spec) { spec = false; __syncthreads(); continue; }       // (early publish: the halo is as fresh as it was)
            ++n_cont;
            {
                const int e = tid >> 5, k = lane;
                int yy, xx, si;
                if (e == 0) { yy = y0 + k; xx = x0 - 1; si = (k + 1) * PT + 1; }
                else if (e == 1) { yy = y0 + k; xx = x0 + TW; si = (k + 1) * PT + TW + 2; }
                else if (e == 2) { yy = y0 - 1; xx = x0 + k; si = k + 2; }
                else { yy = y0 + TILE_H; xx = x0 + k; si = (TILE_H + 1) * PT + k + 2; }
                real v = INF;
                if (yy >= 0 && yy < P.rows && xx >= 0 && xx < P.cols) v = ld_T(&Tq[(long long)yy * P.T_pitch + xx]);
                sT[si] = v;
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
        atomicAdd(&P.q.ctl->noop_visits, n_noop);
        atomicAdd(&P.q.ctl->rounds, n_rounds);
        atomicAdd(&P.q.ctl->continuations, n_cont);
    }
}

}  // namespace fmb
