// truncate.cuh -- rebuilds the PARTIAL field the reference returns when its loop exits
// early (FastMarching.py:108-109 start accepted, :150-155 fronts met;
// FastMarching3D.py:141-142) from the FULL field the tile solver produces.
//
// After k pops the reference's Tmap holds
//   * accepted cells (pop rank <= k): their final value,
//   * narrow-band cells (free, not accepted, next to an accepted cell): the value of
//     their LAST relaxation, i.e. the local update evaluated when their most recently
//     accepted neighbour was popped, from the neighbour values AS OF THAT MOMENT
//     (accepted ones final, the others whatever tentative value they had then),
//   * everything else: +inf.
// "As of that moment" recurses: tent(c, t) = update(val(n, t') for n in N(c)) with
// t' = max{rank(n) <= t}, val(n, t') = F(n) if rank(n) <= t' else tent(n, t').  Times
// strictly decrease along the recursion (the grid has no triangles, so two neighbours
// of one cell are never adjacent), hence it terminates; it is evaluated per narrow-band
// cell with a small explicit stack.  The recursion is a walk over a DAG of (cell, time)
// states.  A cell is re-relaxed only when one of its neighbours pops, so it has at most
// one state per neighbour: a dense memo of NN values per cell (global memory, NaN = not
// yet known) is shared by all threads -- a state is evaluated once instead of being
// re-walked as a tree by every narrow-band cell that reaches it (exponential in 3D, and
// chains along a uniform-cost front are dozens of states deep).  Concurrent evaluations
// of one state write identical bits.  Ranks come from a stable sort of F (host side).
#pragma once
#include "eikonal2d.cuh"
#include "eikonal3d.cuh"

namespace fmb {

constexpr int TRUNC_MAX_DEPTH = 256;      // frames per thread (88-104 B each, local memory)
constexpr int TRUNC_MAX_EVALS = 1 << 16;   // work cap per narrow-band cell (safety net; never reached with the memo)

#ifdef FMB_HOST_EMU
// design-time counters, emulator build only: evaluations, deepest stack, fullest memo, recursive descents
inline long long g_trunc_stats[4] = {0, 0, 0, 0};
#define TRUNC_STAT(i, expr) (g_trunc_stats[i] = (expr))
#else
#define TRUNC_STAT(i, expr) ((void)0)
#endif

template <int D> struct Grid;
template <> struct Grid<2> {
    int rows, cols;
    static constexpr int NN = 4;
    __device__ __forceinline__ long long size() const { return (long long)rows * cols; }
    // neighbour order: x-1, x+1, y-1, y+1
    __device__ __forceinline__ long long nbr(long long c, int i) const {
        const int y = (int)(c / cols), x = (int)(c - (long long)y * cols);
        switch (i) {
            case 0: return x > 0 ? c - 1 : -1;
            case 1: return x < cols - 1 ? c + 1 : -1;
            case 2: return y > 0 ? c - cols : -1;
            default: return y < rows - 1 ? c + cols : -1;
        }
    }
    template <typename real> static __device__ __forceinline__ real update(const real *v, real cost) {
        return eikonal_update<real>(fmin(v[0], v[1]), fmin(v[2], v[3]), cost);
    }
};
template <> struct Grid<3> {
    int ny, nx, nz;
    static constexpr int NN = 6;
    __device__ __forceinline__ long long size() const { return (long long)ny * nx * nz; }
    // neighbour order: x-1, x+1, y-1, y+1, z-1, z+1   (array is [y][x][z])
    __device__ __forceinline__ long long nbr(long long c, int i) const {
        const int z = (int)(c % nz);
        const long long r = c / nz;
        const int x = (int)(r % nx), y = (int)(r / nx);
        switch (i) {
            case 0: return x > 0 ? c - nz : -1;
            case 1: return x < nx - 1 ? c + nz : -1;
            case 2: return y > 0 ? c - (long long)nx * nz : -1;
            case 3: return y < ny - 1 ? c + (long long)nx * nz : -1;
            case 4: return z > 0 ? c - 1 : -1;
            default: return z < nz - 1 ? c + 1 : -1;
        }
    }
    template <typename real> static __device__ __forceinline__ real update(const real *v, real cost) {
        return solve3d_update<real>(v[0] < v[1] ? v[0] : v[1], v[2] < v[3] ? v[2] : v[3], v[4] < v[5] ? v[4] : v[5], cost);
    }
};

// smallest final value among the neighbours of c accepted by time tp (the update's anchor)
template <typename real, int D>
__device__ __forceinline__ real accepted_min(const Grid<D> &g, const int *rank, const real *F, long long c, int tp) {
    real m = num<real>::inf();
#pragma unroll
    for (int i = 0; i < Grid<D>::NN; ++i) {
        const long long n = g.nbr(c, i);
        if (n >= 0 && rank[n] <= tp) m = fmin(m, F[n]);
    }
    return m;
}

// largest rank <= t among the neighbours of c, or -1; `which` = that neighbour's index
template <int D>
__device__ __forceinline__ int last_update_time(const Grid<D> &g, const int *rank, long long c, int t, int *which = nullptr) {
    int best = -1, bi = 0;
#pragma unroll
    for (int i = 0; i < Grid<D>::NN; ++i) {
        const long long n = g.nbr(c, i);
        if (n < 0) continue;
        const int r = rank[n];
        if (r <= t && r > best) { best = r; bi = i; }
    }
    if (which) *which = bi;
    return best;
}

// pass 1: accepted cells and far cells are final here; narrow-band cells (free, not accepted, next
// to an accepted cell) are collected into a dense list so that pass 2 runs with full warps.
template <typename real, int D>
__global__ void truncate_mark_kernel(Grid<D> g, const real *F, const real *cost, const int *rank, int k, real *out,
                                     int *list, int *count) {
    const real INF = num<real>::inf();
    const long long total = g.size();
    for (long long c0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += (long long)gridDim.x * blockDim.x) {
        if (rank[c0] <= k) { out[c0] = F[c0]; continue; }
        out[c0] = INF;
        if (!(cost[c0] < INF)) continue;
        if (last_update_time<D>(g, rank, c0, k) < 0) continue;
        list[atomicAdd(count, 1)] = (int)c0;
    }
}

// pass 2: replay the last relaxation of every narrow-band cell (one thread per listed cell).
// memo: NN doubles per cell, all NaN on entry; slot c*NN + i = tent(c, rank of c's neighbour i).
// A walk that runs out of stack (chains along a uniform-cost front are ~100 states deep at planner
// scale) uses the final value as a stand-in for the state it cannot descend into; the states above
// that point are "tainted": they are not memoised, and the cell is counted in *overflow.
template <typename real, int D>
__global__ void truncate_replay_kernel(Grid<D> g, const real *F, const real *cost, const int *rank, int k, real *out,
                                       const int *list, const int *count, int *overflow, real *memo) {
    constexpr int NN = Grid<D>::NN;
    const real INF = num<real>::inf();
    const int n_list = *count;
    struct Frame { long long c; int tp; int stage; int slot; real amin; real u; real v[NN]; };
    for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < n_list; li += gridDim.x * blockDim.x) {
        const long long c0 = list[li];
        int w0;
        const int t0 = last_update_time<D>(g, rank, c0, k, &w0);
        Frame st[TRUNC_MAX_DEPTH];
        real result = INF;
        bool tainted_top = false;
        {
            int sp = 0, taint = -1;                 // frames 0..taint depend on a stand-in
            bool top_dirty = false;
            st[0].c = c0; st[0].tp = t0; st[0].stage = 0; st[0].slot = w0; st[0].amin = accepted_min<real, D>(g, rank, F, c0, t0);
            bool have_result = false;
            int budget = TRUNC_MAX_EVALS;
            while (sp >= 0) {
                Frame &f = st[sp];
                if (have_result) {
                    have_result = false;
                    if (f.stage == NN + 1) {             // came back with this cell's previous tentative value
                        result = result < f.u ? result : f.u;
                        if (sp > taint) { if (sp > 0) __stcg(&memo[f.c * NN + f.slot], result); }
                        else { taint = sp - 1; top_dirty |= sp == 0; }
                        have_result = true;
                        --sp;
                        continue;
                    }
                    f.v[f.stage++] = result;
                }
                bool descended = false;
                while (f.stage < NN) {
                    const long long n = g.nbr(f.c, f.stage);
                    real val;
                    if (n < 0) val = INF;
                    else if (rank[n] <= f.tp) val = F[n];
                    else if (!(cost[n] < INF)) val = INF;
                    else {
                        // Exact pruning.  A tentative value is >= the cell's final value F[n], so it cannot
                        // matter when (a) it is not the minimum of its axis, or (b) it is at least one
                        // cost above the smallest accepted neighbour (the upwind solvers then drop it):
                        // in both cases any stand-in >= F[n] gives the same update.
                        const int sib_i = f.stage ^ 1;
                        const long long sib = g.nbr(f.c, sib_i);
                        const bool sib_wins = (sib >= 0 && rank[sib] <= f.tp && F[sib] <= F[n]) ||
                                              (sib_i < f.stage && f.v[sib_i] <= F[n]);
                        int wn = 0;
                        const int tn = (sib_wins || !(F[n] - f.amin < cost[f.c])) ? -2 : last_update_time<D>(g, rank, n, f.tp, &wn);
                        if (tn == -2) val = F[n];
                        else if (tn < 0) val = INF;
                        else {
                            const real known = __ldcg(&memo[n * NN + wn]);       // tent(n, tn), if already evaluated
                            if (known == known) val = known;
                            else if (sp + 1 >= TRUNC_MAX_DEPTH || budget <= 0) { val = F[n]; taint = sp; }
                            else {
                                ++sp;
                                TRUNC_STAT(1, sp > g_trunc_stats[1] ? sp : g_trunc_stats[1]);
                                TRUNC_STAT(3, g_trunc_stats[3] + 1);
                                st[sp].c = n; st[sp].tp = tn; st[sp].stage = 0; st[sp].slot = wn;
                                st[sp].amin = accepted_min<real, D>(g, rank, F, n, tn);
                                descended = true;
                                break;
                            }
                        }
                    }
                    f.v[f.stage++] = val;
                }
                if (descended) continue;
                result = Grid<D>::template update<real>(f.v, cost[f.c]);
                --budget;
                TRUNC_STAT(0, g_trunc_stats[0] + 1);
                // The reference keeps a new value only when it is lower (`if T < Tmap[child]`,
                // FastMarching.py:70, FastMarching3D.py:86), and in floating point a later update can come
                // out an ulp above an earlier one: the tentative value is the minimum over this cell's
                // updates so far.  Nothing earlier can undercut the final value itself.
                if (result > F[f.c]) {
                    int wp = 0;
                    const int tprev = last_update_time<D>(g, rank, f.c, f.tp - 1, &wp);
                    if (tprev >= 0) {
                        const real known = __ldcg(&memo[f.c * NN + wp]);
                        if (known == known) result = known < result ? known : result;
                        else if (sp + 1 >= TRUNC_MAX_DEPTH || budget <= 0) taint = sp;
                        else {
                            f.u = result; f.stage = NN + 1;
                            ++sp;
                            TRUNC_STAT(1, sp > g_trunc_stats[1] ? sp : g_trunc_stats[1]);
                            TRUNC_STAT(3, g_trunc_stats[3] + 1);
                            st[sp].c = f.c; st[sp].tp = tprev; st[sp].stage = 0; st[sp].slot = wp;
                            st[sp].amin = accepted_min<real, D>(g, rank, F, f.c, tprev);
                            continue;
                        }
                    }
                }
                if (sp > taint) { if (sp > 0) __stcg(&memo[f.c * NN + f.slot], result); }
                else { taint = sp - 1; top_dirty |= sp == 0; }
                have_result = true;
                --sp;
            }
            tainted_top = top_dirty;
        }
        if (tainted_top) atomicAdd(overflow, 1);
        out[c0] = result;
    }
}

}  // namespace fmb
