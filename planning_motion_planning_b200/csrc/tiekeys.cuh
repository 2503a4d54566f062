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
#include "truncate.cuh"      // Grid<D>: neighbours and the local update in 2D / 3D

namespace fmb {

__global__ void tie_keys2d_kernel(const double *T, const double *cost, const int *rank, const int *tau, const int *group,
                                  int rows, int cols, int seed_idx, int transposed, int *tau_new, long long *key) {
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
        // popped neighbour left/right/up/down => my child index in its updateNode; when the map is the
        // transpose of the caller's (F-ordered input), left/right are the caller's up/down
        const int ci[4] = {transposed ? 2 : 4, transposed ? 1 : 3, transposed ? 4 : 2, transposed ? 3 : 1};
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

// ---- exact tie order in one ordered sweep ----------------------------------------------------------
// Only STRICTLY upwind neighbours (smaller T) can take part in a cell's final update, and they all
// pop before the cell's tie group starts, so insertion times and in-group ranks can be settled group
// by group in ascending T with no iteration: ticket p = p-th cell of the T-sorted list.  A cell waits
// for its upwind neighbours (earlier groups => smaller tickets, held by running threads or done),
// computes its insertion time and child index exactly as tie_keys2d_kernel does, then waits for the
// rest of its tie group and takes its place among them (key ascending, cell index on equal keys).
// Groups must fit among the resident threads (the host falls back to the iterated sort otherwise).
constexpr int TIE_MAX_SPINS = 1 << 22;
__device__ __forceinline__ bool tie_wait_nonzero(const int *flag, int want_at_least) {
    for (int s = 0; s < TIE_MAX_SPINS; ++s) {
        if (ld_volatile(flag) >= want_at_least) { __threadfence(); return true; }
        __nanosleep(40);                              // (busy polling was measured slower: it starves the producers)
    }
    return false;
}
// child index (1-based position in the reference's updateNode loop) of a cell as seen from its popped
// neighbour i (Grid<D>::nbr order).  2D (FastMarching.py:46-54): children (0,-1), (0,+1), (-1,0), (+1,0);
// 3D (FastMarching3D.py:22-33): z-1, z+1, x-1, x+1, y+1, y-1.
// `transposed` (2D only): the field is the transpose of the caller's map (an F-ordered input solved as its
// C-ordered transpose), so this grid's x-neighbours are the caller's y-neighbours.
template <int D> __device__ __forceinline__ int tie_child_index(int i, int transposed);
template <> __device__ __forceinline__ int tie_child_index<2>(int i, int transposed) {
    const int ci[4] = {4, 3, 2, 1}, ct[4] = {2, 1, 4, 3};
    return transposed ? ct[i] : ci[i];
}
template <> __device__ __forceinline__ int tie_child_index<3>(int i, int) { const int ci[6] = {4, 3, 5, 6, 2, 1}; return ci[i]; }

template <int D>
__global__ void tie_sweep_kernel(Grid<D> g, const double *T, const double *cost, const int *members, const int *gstart,
                                 const int *gsize, int seed_idx, int transposed, int *rank, int *tau, long long *key, int *done,
                                 int *gcount, int *ticket, int *failed, const int *enable = nullptr) {
    constexpr int NN = Grid<D>::NN;
    if (enable && !*enable) return;                // decided on the device (fm_capi_ranks.inc): no ties, or nothing to settle
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const long long BIG = 0x7fffffffLL;
    const long long total = g.size();
    const int lane = threadIdx.x & 31;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ticket, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= total) break;
        const long long p = (long long)base + lane;
        if (p < total) {
            const int c = members[p];
            const double t = T[c];
            if (!(t < INF)) rank[c] = 0x7fffffff;
            else {
                long long best = BIG;
                int cidx = 0;
                if (c == seed_idx) best = -1;
                else {
                    double tv[NN]; long long rv[NN], av[NN];
#pragma unroll
                    for (int i = 0; i < NN; ++i) {
                        tv[i] = INF; rv[i] = BIG; av[i] = BIG;
                        const long long nb = g.nbr(c, i);
                        if (nb >= 0) {
                            const double v = T[nb];
                            // strictly upwind: its rank and insertion time are final before mine.  A neighbour of my own
                            // tie group is never upwind (groups may be formed with a tolerance of a few ulp, and waiting
                            // for a member of my own group would deadlock on the group barrier below)
                            if (v < t && gstart[nb] != gstart[c]) {
                                if (!tie_wait_nonzero(&done[nb], 1)) atomicAdd(failed, 1);
                                tv[i] = v; rv[i] = ld_volatile(&rank[nb]); av[i] = ld_volatile(&tau[nb]);
                            }
                        }
                    }
                    const double c_cost = cost[c];
                    const double limit = t * (1.0 + 1e-14);
#pragma unroll
                    for (int i = 0; i < NN; ++i) {
                        const long long ti = rv[i];
                        if (ti >= BIG || ti >= best) continue;
                        double v[NN];
#pragma unroll
                        for (int j = 0; j < NN; ++j) v[j] = av[j] <= ti ? tv[j] : INF;
                        if (Grid<D>::template update<double>(v, c_cost) <= limit) { best = ti; cidx = tie_child_index<D>(i, transposed); }
                    }
                }
                tau[c] = (int)best;
                int r = gstart[c];
                const int n = gsize[c];
                if (n > 1) {
                    const long long kc = ((long long)(0xffffffffLL - (unsigned long long)(best + 1)) << 3) | (long long)(7 - cidx);
                    *reinterpret_cast<volatile long long *>(&key[c]) = kc;
                    __threadfence();
                    atomicAdd(&gcount[r], 1);
                    if (!tie_wait_nonzero(&gcount[r], n)) atomicAdd(failed, 1);
                    int before = 0;
                    for (int j = 0; j < n; ++j) {
                        const int m = members[r + j];
                        const long long km = *reinterpret_cast<const volatile long long *>(&key[m]);
                        before += (km < kc || (km == kc && m < c)) ? 1 : 0;
                    }
                    r += before;
                }
                rank[c] = r;
                __threadfence();
                *reinterpret_cast<volatile int *>(&done[c]) = 1;
            }
        }
        __syncwarp();
    }
}

// ---- where the two fronts of biComputeTmap meet (FastMarching.py:141-155) ------------------------
// Both fronts pop one node per round; the loop stops in the first round k in which G's node is already
// closed in S (tested first, :150-152) or S's node is closed in G (:153-155).  With pop ranks that is
// k = min over cells of max(rankG, rankS); the join node is the cell popped by G in round k if it
// attains the minimum, else the one popped by S.  out[0] = k, out[1] = join cell (both INT_MAX when the
// fronts never meet); pass 1 reduces k, pass 2 picks the cell (G's pop wins, then the smaller index).
__global__ void bi_join_k_kernel(const int *rankG, const int *rankS, long long total, int *out) {
    int best = 0x7fffffff;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int a = rankG[c], b = rankS[c];
        const int m = a > b ? a : b;
        best = m < best ? m : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(FULL, best, o); best = v < best ? v : best; }
    if ((threadIdx.x & 31) == 0 && best != 0x7fffffff) atomicMin(&out[0], best);
}
__global__ void bi_join_cell_kernel(const int *rankG, const int *rankS, long long total, int *out, unsigned long long *pick) {
    const int k = out[0];
    if (k == 0x7fffffff) return;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int a = rankG[c], b = rankS[c];
        if ((a > b ? a : b) == k) atomicMin(pick, ((unsigned long long)(a == k ? 0 : 1) << 40) | (unsigned long long)c);
    }
}
__global__ void bi_join_finish_kernel(int *out, const unsigned long long *pick) {
    if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = out[0] == 0x7fffffff ? 0x7fffffff : (int)(*pick & 0xffffffffffULL);
}

}  // namespace fmb
