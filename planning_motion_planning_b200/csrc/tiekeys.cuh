// tiekeys.cuh -- one refinement step of the reference's pop order among exactly equal T values
// (FastMarching.py:65-67,76-78: bisect_left + insert => the node (re)inserted last pops first).
//
// For every cell: the pop ("time") at which its final value is inserted = the first pop of one of
// its neighbours at which every input of its final update already carries its final value
// (tentative neighbour values count, :57-62), and its child index in that updateNode call (:46-54).
// The sort key (tie group of T ascending, insertion time descending, child index descending) is
// packed into 64 bits so that one stable sort per iteration yields the next ranks; the host loop
// (FastMarching/_compat.py) iterates to the fixed point.  Same arithmetic as the torch reference
// implementation there, which the tests compare against the reference's true pop order.
#pragma once
#include "fm_common.cuh"

namespace fmb {

__global__ void tie_keys2d_kernel(const double *T, const double *cost, const int *rank, const int *tau, const int *group,
                                  int rows, int cols, int seed_idx, int *tau_new, long long *key) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const long long BIG = 0x7fffffffLL;
    const long long total = (long long)rows * cols;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const double t = T[c];
        if (!(t < INF)) { tau_new[c] = 0x7fffffff; key[c] = 0x7fffffffffffffffLL; continue; }
        const int y = (int)(c / cols), x = (int)(c - (long long)y * cols);
        double tv[4]; long long rv[4], av[4];
        const long long nb[4] = {x > 0 ? c - 1 : -1, x < cols - 1 ? c + 1 : -1, y > 0 ? c - cols : -1, y < rows - 1 ? c + cols : -1};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            tv[i] = INF; rv[i] = BIG; av[i] = BIG;
            if (nb[i] >= 0) {
                tv[i] = T[nb[i]];
                if (tv[i] < INF) { rv[i] = rank[nb[i]]; av[i] = tau[nb[i]]; }
                if (nb[i] == seed_idx) { rv[i] = 0; av[i] = -1; }
            }
        }
        // insertion time = the earliest neighbour pop at which the update, fed only with neighbour
        // values that are already final by then (the others count as +inf), reproduces the final value
        const double c_cost = cost[c];
        const double limit = t * (1.0 + 1e-14);               // the solver's field is a fixed point to a few ulp
        long long best = BIG; int cidx = 0;
        const int ci[4] = {4, 3, 2, 1};          // popped neighbour left/right/up/down => my child index in its updateNode
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long ti = rv[i];
            if (ti >= BIG || ti >= best) continue;
            const double l = av[0] <= ti ? tv[0] : INF, r = av[1] <= ti ? tv[1] : INF;
            const double u = av[2] <= ti ? tv[2] : INF, d = av[3] <= ti ? tv[3] : INF;
            const double a = l < r ? l : r, b = u < d ? u : d;
            const double m = a < b ? a : b, dd = a - b;
            double v;
            if (!(fabs(dd) <= c_cost)) v = m + c_cost;
            else v = 0.5 * (a + b + sqrt(2.0 * (c_cost * c_cost) - dd * dd));
            if (v <= limit) { best = ti; cidx = ci[i]; }
        }
        if (c == seed_idx) best = -1;
        tau_new[c] = (int)best;
        key[c] = ((long long)group[c] << 35) | ((long long)(0xffffffffLL - (unsigned long long)(best + 1)) << 3) | (long long)(7 - cidx);
    }
}

// Re-rank inside tie groups after a key step: the global stable sort of `key` (its top bits are the
// tie group) only permutes cells WITHIN a group, so a cell's new rank is its group's first rank plus
// the number of members that sort before it (smaller key, or equal key and smaller cell index --
// what a stable sort of the cell-indexed key array does).  members[gstart .. gstart+gsize) lists the
// cells of the group; singleton groups keep their rank.  *changed is raised when any rank or
// insertion time moved, so the host polls one flag instead of comparing arrays.
__global__ void tie_rerank_kernel(const long long *key, const int *members, const int *gstart, const int *gsize,
                                  const int *rank, const int *tau, const int *tau_new, long long total, int *rank_new,
                                  int *changed) {
    bool moved = false;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int n = gsize[c];
        int r = rank[c];
        if (n > 1) {
            const int s = gstart[c];
            const long long kc = key[c];
            int before = 0;
            for (int j = 0; j < n; ++j) {
                const int m = members[s + j];
                const long long km = key[m];
                before += (km < kc || (km == kc && m < (int)c)) ? 1 : 0;
            }
            r = s + before;
        }
        moved |= r != rank[c] || tau_new[c] != tau[c];
        rank_new[c] = r;
    }
    if (moved) *changed = 1;
}

}  // namespace fmb
