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
// cell with a small explicit stack.  The recursion is a walk over a small DAG of
// (cell, time) states; a per-thread memo table keeps it from being re-walked as a tree
// (which is exponential in 3D).  Ranks come from a stable sort of F (host side).
#pragma once
#include "eikonal2d.cuh"
#include "eikonal3d.cuh"

namespace fmb {

constexpr int TRUNC_MAX_DEPTH = 64;
constexpr int TRUNC_MAX_EVALS = 1 << 16;   // work cap per narrow-band cell (safety net; never reached with the memo)
constexpr int TRUNC_MEMO = 1024;            // memo entries per thread (power of two)

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

// largest rank <= t among the neighbours of c, or -1
template <int D>
__device__ __forceinline__ int last_update_time(const Grid<D> &g, const int *rank, long long c, int t) {
    int best = -1;
#pragma unroll
    for (int i = 0; i < Grid<D>::NN; ++i) {
        const long long n = g.nbr(c, i);
        if (n < 0) continue;
        const int r = rank[n];
        if (r <= t && r > best) best = r;
    }
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

// pass 2: replay the last relaxation of every narrow-band cell (one thread per listed cell)
template <typename real, int D>
__global__ void truncate_replay_kernel(Grid<D> g, const real *F, const real *cost, const int *rank, int k, real *out,
                                       const int *list, const int *count, int *overflow) {
    constexpr int NN = Grid<D>::NN;
    const real INF = num<real>::inf();
    const int n_list = *count;
    struct Frame { long long c; int tp; int stage; real amin; real v[NN]; };
    for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < n_list; li += gridDim.x * blockDim.x) {
        const long long c0 = list[li];
        const int t0 = last_update_time<D>(g, rank, c0, k);
        Frame st[TRUNC_MAX_DEPTH];
        int sp = 0;
        st[0].c = c0; st[0].tp = t0; st[0].stage = 0; st[0].amin = accepted_min<real, D>(g, rank, F, c0, t0);
        real result = INF;
        bool have_result = false;
        int budget = TRUNC_MAX_EVALS;
        int memo_cell[TRUNC_MEMO], memo_tp[TRUNC_MEMO], memo_used = 0;
        real memo_val[TRUNC_MEMO];
        for (int i = 0; i < TRUNC_MEMO; ++i) memo_cell[i] = -1;
        while (sp >= 0) {
            Frame &f = st[sp];
            if (have_result) { f.v[f.stage++] = result; have_result = false; }
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
                    const int tn = (sib_wins || !(F[n] - f.amin < cost[f.c])) ? -2 : last_update_time<D>(g, rank, n, f.tp);
                    if (tn == -2) val = F[n];
                    else if (tn < 0) val = INF;
                    else if (sp + 1 >= TRUNC_MAX_DEPTH || budget <= 0) { val = F[n]; atomicAdd(overflow, 1); }
                    else {
                        // memo lookup: tent(n, .) depends only on (n, tn)
                        unsigned h = ((unsigned)n * 2654435761u + (unsigned)tn * 40503u) & (TRUNC_MEMO - 1);
                        bool hit = false;
                        for (int probe = 0; probe < 16; ++probe) {
                            const unsigned e = (h + probe) & (TRUNC_MEMO - 1);
                            if (memo_cell[e] < 0) break;
                            if (memo_cell[e] == (int)n && memo_tp[e] == tn) { val = memo_val[e]; hit = true; break; }
                        }
                        if (hit) { f.v[f.stage++] = val; continue; }
                        ++sp;
                        st[sp].c = n; st[sp].tp = tn; st[sp].stage = 0; st[sp].amin = accepted_min<real, D>(g, rank, F, n, tn);
                        descended = true;
                        break;
                    }
                }
                f.v[f.stage++] = val;
            }
            if (descended) continue;
            result = Grid<D>::template update<real>(f.v, cost[f.c]);
            --budget;
            if (sp > 0 && memo_used < (TRUNC_MEMO * 3) / 4) {      // remember tent(f.c, f.tp)
                unsigned h = ((unsigned)f.c * 2654435761u + (unsigned)f.tp * 40503u) & (TRUNC_MEMO - 1);
                for (int probe = 0; probe < 16; ++probe) {
                    const unsigned e = (h + probe) & (TRUNC_MEMO - 1);
                    if (memo_cell[e] < 0) { memo_cell[e] = (int)f.c; memo_tp[e] = f.tp; memo_val[e] = result; ++memo_used; break; }
                }
            }
            have_result = true;
            --sp;
        }
        out[c0] = result;
    }
}

}  // namespace fmb
