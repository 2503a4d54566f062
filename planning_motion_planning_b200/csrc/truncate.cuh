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
// "As of that moment" chains backwards in time: tent(c, t) = update(val(n, t) for n in N(c)),
// val(n, t) = F(n) if rank(n) <= t, else the tentative value n itself held at time t.  A cell is
// re-relaxed only when one of its neighbours pops, so it has at most one tentative state per
// neighbour; the reference keeps a new value only if it is lower (`if T < Tmap[child]`), so a state
// is the minimum of the update and the cell's previous state.
//
// Instead of recursing from the narrow band (deep, serial chains along uniform-cost fronts), the
// kernels replay the relaxations the reference performed during its first k pops, in the
// reference's own order, in parallel: ticket r*NN + j = "pop r relaxes its j-th neighbour".  Tickets
// are handed out in order; a state waits (spin) for the few earlier states it reads -- they belong to
// strictly earlier pops, i.e. to smaller tickets, which are held by running threads or finished, so
// the wait cannot deadlock and the critical path is the longest dependency chain, not a walk.  The
// states live in a dense memo of NN doubles per cell (slot c*NN + i = value of c after its neighbour i
// popped; all-ones bits = not yet written).  The state of a narrow-band cell after its last update
// <= k is its entry in the partial field.  Ranks come from a sort of F (fm_capi_ranks.inc).
//
// Two forms.  The DENSE form replays every relaxation: NN (k + 1) tickets (33.5 M for a 4096^2 bi-solve: 17 ms per
// front).  The SPARSE form (round 2, the default behind fmb_bisolve2d_f64 / fmb_solve*_until_f64) first computes which
// states the narrow band of time k depends on at all -- a backward closure over "reads the tentative state of" that
// runs along the front for hundreds of rounds but touches only ~1.5 % of the relaxations (475 k tickets there) -- and
// replays those, sorted, with the same code: 2.5 ms per front.  It reports failure on the device (a list overflowed,
// the closure did not settle) and the dense form then runs instead, without a host round trip.
#pragma once
#include "eikonal2d.cuh"
#include "eikonal3d.cuh"

namespace fmb {

constexpr int TRUNC_MAX_SPINS = 1 << 22;    // safety net of the dependency wait (never reached; counted in *overflow)

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
    // coordinate forms (32-bit, one division per cell): p = {x, y}; step = neighbour i of the cell at p: flat index or -1, q = its coordinates
    __device__ __forceinline__ void split(int c, int *p) const { p[1] = c / cols; p[0] = c - p[1] * cols; }
    __device__ __forceinline__ int step(const int *p, int i, int *q) const {
        q[0] = p[0] + (i == 0 ? -1 : i == 1 ? 1 : 0);
        q[1] = p[1] + (i == 2 ? -1 : i == 3 ? 1 : 0);
        const bool in = (unsigned)q[0] < (unsigned)cols && (unsigned)q[1] < (unsigned)rows;
        return in ? q[1] * cols + q[0] : -1;
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
    // coordinate forms: p = {x, y, z}
    __device__ __forceinline__ void split(int c, int *p) const {
        const int r = c / nz;
        p[2] = c - r * nz; p[1] = r / nx; p[0] = r - p[1] * nx;
    }
    __device__ __forceinline__ int step(const int *p, int i, int *q) const {
        q[0] = p[0] + (i == 0 ? -1 : i == 1 ? 1 : 0);
        q[1] = p[1] + (i == 2 ? -1 : i == 3 ? 1 : 0);
        q[2] = p[2] + (i == 4 ? -1 : i == 5 ? 1 : 0);
        const bool in = (unsigned)q[0] < (unsigned)nx && (unsigned)q[1] < (unsigned)ny && (unsigned)q[2] < (unsigned)nz;
        return in ? (q[1] * nx + q[0]) * nz + q[2] : -1;
    }
};

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

// pass 1: accepted cells keep their final value, everything else starts at +inf (narrow-band cells
// are overwritten by pass 2); order[r] = the cell popped r-th, for r <= k.
template <typename real, int D>
__global__ void truncate_mark_kernel(Grid<D> g, const real *F, const int *rank, int k, real *out, int *order, const int *k_dev = nullptr) {
    const real INF = num<real>::inf();
    const long long total = g.size();
    if (k_dev) k = *k_dev;                         // k decided on the device (fm_capi_ranks.inc)
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int r = rank[c];
        if (r <= k) { out[c] = F[c]; if (r >= 0 && r < total) order[r] = (int)c; }
        else out[c] = INF;
    }
}

constexpr unsigned long long MEMO_UNKNOWN = 0xffffffffffffffffULL;

template <typename real>
__device__ __forceinline__ real memo_wait(const real *memo, long long slot, real fallback, int *overflow) {
    const volatile unsigned long long *p = reinterpret_cast<const volatile unsigned long long *>(memo) + slot;
    for (int spins = 0; spins < TRUNC_MAX_SPINS; ++spins) {
        const unsigned long long b = *p;
        if (b != MEMO_UNKNOWN) { __threadfence(); return (real)__longlong_as_double((long long)b); }
        __nanosleep(40);                              // (busy polling was measured slower: it starves the producers)
    }
    atomicAdd(overflow, 1);
    return fallback;
}

// one relaxation of the replay: ticket tk = r * NN + j = "pop r relaxes its j-th neighbour" (see the header of this file)
template <typename real, int D>
__device__ __forceinline__ void replay_ticket(long long tk, const Grid<D> &g, const real *F, const real *cost, const int *rank,
                                              const int *order, int k, real *out, real *memo, int *overflow) {
    constexpr int NN = Grid<D>::NN;
    const real INF = num<real>::inf();
    const int r = (int)(tk / NN), j = (int)(tk - (long long)r * NN);
    const long long a = order[r];
    const long long c = g.nbr(a, j);
    if (c >= 0 && rank[c] > r && cost[c] < INF) {           // a free neighbour that is not accepted yet (:56, FastMarching3D.py:35)
        real v[NN];
        int newer = -1;                                      // a neighbour of c that pops in (r, k]: not c's last update
#pragma unroll
        for (int i = 0; i < NN; ++i) {
            const long long m = g.nbr(c, i);
            real val = INF;
            if (m >= 0) {
                const int rm = rank[m];
                if (rm <= r) val = F[m];
                else {
                    if (rm <= k) newer = i;
                    if (cost[m] < INF) {
                        int wm = 0;
                        const int tm = last_update_time<D>(g, rank, m, r, &wm);
                        if (tm >= 0) val = memo_wait<real>(memo, m * NN + wm, F[m], overflow);
                    }
                }
            }
            v[i] = val;
        }
        real u = Grid<D>::template update<real>(v, cost[c]);
        int wp = 0;
        const int tprev = last_update_time<D>(g, rank, c, r - 1, &wp);
        if (tprev >= 0) {
            const real prev = memo_wait<real>(memo, c * NN + wp, u, overflow);
            u = prev < u ? prev : u;
        }
        if (newer < 0 && rank[c] > k) out[c] = u;            // c's state when the reference stops
        __threadfence();
        *(reinterpret_cast<volatile unsigned long long *>(memo) + (c * NN + (j ^ 1))) =
            (unsigned long long)__double_as_longlong((double)u);
    }
}

// pass 2, dense form: every relaxation of the first k pops, in pop order.  run_if (optional): the kernel only runs when
// *run_if != 0 -- the sparse form below failed (its frontier or ticket list overflowed, or its marking did not settle).
template <typename real, int D>
__global__ void truncate_sweep_kernel(Grid<D> g, const real *F, const real *cost, const int *rank, const int *order, int k,
                                      real *out, real *memo, int *ticket, int *overflow, const int *k_dev = nullptr,
                                      const int *run_if = nullptr) {
    constexpr int NN = Grid<D>::NN;
    const int lane = threadIdx.x & 31;
    if (run_if && *run_if == 0) return;
    if (k_dev) k = *k_dev;
    if (k >= g.size()) return;                     // every reached cell is accepted (no early exit): nothing to replay
    const long long n_tickets = ((long long)k + 1) * NN;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ticket, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n_tickets) break;
        const long long tk = (long long)base + lane;
        if (tk < n_tickets) replay_ticket<real, D>(tk, g, F, cost, rank, order, k, out, memo, overflow);
        __syncwarp();
    }
}
template <typename real>
__global__ void memo_fill_kernel(real *memo, long long n, const int *run_if) {
    if (run_if && *run_if == 0) return;
    unsigned long long *p = reinterpret_cast<unsigned long long *>(memo);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = MEMO_UNKNOWN;
}

// ---------------------------------------------------------------------------
// pass 2, sparse form: only the relaxations the narrow band of time k depends on.
//
// The partial field only shows the states of the narrow-band cells at time k.  A state reads (a) the earlier states of its
// own cell and (b), for every neighbour that was not yet accepted at that moment, that neighbour's state after ITS last
// relaxation before the moment -- which reads its own earlier states and its tentative neighbours, and so on backwards in
// time along the front.  need_t[c] = the latest moment up to which the states of cell c are needed (-1: none); the
// closure is computed by a frontier expansion (a cell whose bound rises is expanded again), a few rounds in practice
// because every link steps back in time towards the moment a cell was first touched.  The needed relaxations are then
// emitted as tickets from the cell side, sorted (the in-order hand-out is what makes the dependency waits deadlock
// free) and replayed by the same code as the dense form: a few 10^5 tickets instead of 4 (k + 1).
// counters (int[64], zeroed by the host): [0] dense ticket cursor, [1] waits at the limit, [4] tickets emitted,
// [5] sparse form failed -> the dense form runs, [6] list cursor, [7] rounds of the expansion, [8 + i] size of the
// frontier of grid-wide round i.
constexpr int CONE_ROUNDS = 8;          // grid-wide rounds before the one-block tail
constexpr int CONE_TAIL_ROUNDS = 1 << 16;
constexpr int CONE_TICKET_PAD = 0x7f7f7f7f;       // memset(0x7f) of the ticket list: above every valid ticket

template <typename real, int D>
__global__ void cone_seed_kernel(Grid<D> g, const real *cost, const int *rank, const int *k_dev, int *need_t, int *front,
                                 int *counters, int cap) {
    constexpr int NN = Grid<D>::NN;
    const real INF = num<real>::inf();
    const int k = *k_dev;
    const long long total = g.size();
    if (k >= total) return;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int t = -1;
        if (rank[c] > k && cost[c] < INF) {
#pragma unroll
            for (int i = 0; i < NN; ++i) {
                const long long n = g.nbr(c, i);
                if (n >= 0 && rank[n] <= k) t = k;
            }
        }
        need_t[c] = t;
        if (t >= 0) {
            const int pos = atomicAdd(&counters[8], 1);
            if (pos < cap) front[pos] = (int)c; else counters[5] = 1;
        }
    }
}

// expands one frontier cell: every needed relaxation of c (those up to moment t) reads the tentative neighbours of that
// moment.  All ranks the decisions need (the neighbours' and the neighbours' neighbours') are loaded up front -- one
// memory latency per cell instead of one per decision: the tail of the expansion is a chain of hundreds of rounds of a
// few cells each.  push(cell, bound) receives every cell whose bound this call raised.
template <typename real, int D, typename Push>
__device__ __forceinline__ void cone_expand_cell(const Grid<D> &g, const real *cost, const int *rank, int *need_t, long long c, int t,
                                                 Push push) {
    constexpr int NN = Grid<D>::NN;
    const real INF = num<real>::inf();
    int nb[NN], rn[NN], r2[NN][NN];
    bool open[NN];
    int pc[3], pn[3], pq[3];
    g.split((int)c, pc);                           // (cells < 2^31: truncate_dk's guard) one division per cell, the rest by coordinates
#pragma unroll
    for (int i = 0; i < NN; ++i) {
        nb[i] = g.step(pc, i, pn);
        rn[i] = nb[i] >= 0 ? rank[nb[i]] : -1;
        open[i] = nb[i] >= 0 && cost[nb[i]] < INF;
#pragma unroll
        for (int q = 0; q < NN; ++q) {
            const int m2 = nb[i] >= 0 ? g.step(pn, q, pq) : -1;
            r2[i][q] = m2 >= 0 ? rank[m2] : 0x7fffffff;
        }
    }
    // per neighbour the latest bound any needed relaxation of c asks of it (one atomic per neighbour, all of them in
    // flight before the first result is looked at)
    int want[NN], old[NN];
#pragma unroll
    for (int i = 0; i < NN; ++i) {
        want[i] = -1;
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            const int r = rn[j];
            const bool event = nb[j] >= 0 && r <= t;           // a relaxation of c that is needed
            const bool tentative = nb[i] >= 0 && rn[i] > r && open[i];         // not accepted by then, no obstacle
            int tm = -1;                                       // the neighbour's last relaxation before that moment
#pragma unroll
            for (int q = 0; q < NN; ++q) tm = (r2[i][q] <= r && r2[i][q] > tm) ? r2[i][q] : tm;
            want[i] = (event && tentative && tm > want[i]) ? tm : want[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NN; ++i) old[i] = want[i] >= 0 ? atomicMax(&need_t[nb[i]], want[i]) : 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NN; ++i)
        if (old[i] < want[i]) push(nb[i], want[i]);
}

// the first rounds of the expansion, grid-wide (thousands of cells per round) ...
template <typename real, int D>
__global__ void cone_expand_kernel(Grid<D> g, const real *cost, const int *rank, const int *k_dev, int *need_t,
                                   const int *fin, int *fout, int *counters, int round, int cap) {
    if (*k_dev >= g.size()) return;
    const int n_in = min(counters[8 + round], cap);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_in; idx += gridDim.x * blockDim.x) {
        const long long c = fin[idx];
        cone_expand_cell<real, D>(g, cost, rank, need_t, c, *reinterpret_cast<volatile int *>(&need_t[c]), [&](int m, int) {
            const int pos = atomicAdd(&counters[8 + round + 1], 1);
            if (pos < cap) fout[pos] = m; else counters[5] = 1;
        });
    }
}
// ... and its long thin tail (tens of cells per round for hundreds of rounds: the chains that run along the front) in
// ONE block that loops over the rounds with a block barrier instead of a launch per round.  The first CONE_SMEM_CAP
// entries of a frontier (cell, bound) live in shared memory, the rest (cells only) spill to the global lists: a round
// of the tail then costs two global round trips (the ranks, the atomic), not four.
constexpr int CONE_SMEM_CAP = 2048;
constexpr int CONE_TAIL_SMEM = 16 + 2 * CONE_SMEM_CAP * 8;
template <typename real, int D>
__global__ void __launch_bounds__(1024) cone_tail_kernel(Grid<D> g, const real *cost, const int *rank, const int *k_dev, int *need_t,
                                                         int *fa, int *fb, int *counters, int first_round, int cap, int max_rounds) {
    FMB_DYN_SMEM(smem_raw);
    int &s_in = reinterpret_cast<int *>(smem_raw)[0], &s_out = reinterpret_cast<int *>(smem_raw)[1];
    int2 *sin_ = reinterpret_cast<int2 *>(smem_raw + 16), *sout = sin_ + CONE_SMEM_CAP;
    if (*k_dev >= g.size()) return;
    int *gin = (first_round & 1) ? fb : fa, *gout = (first_round & 1) ? fa : fb;
    if (threadIdx.x == 0) { s_in = min(counters[8 + first_round], cap); s_out = 0; }
    __syncthreads();
    int rounds = 0, in_smem = 0;                   // entries of the input frontier that sit in shared memory
    for (; rounds < max_rounds; ++rounds) {
        const int n_in = s_in;
        if (n_in == 0) break;
        for (int idx = threadIdx.x; idx < n_in; idx += blockDim.x) {
            int c, t;
            if (idx < in_smem) { c = sin_[idx].x; t = sin_[idx].y; }
            else { c = gin[idx - in_smem]; t = *reinterpret_cast<volatile int *>(&need_t[c]); }
            cone_expand_cell<real, D>(g, cost, rank, need_t, c, t, [&](int m, int tm) {
                const int pos = atomicAdd(&s_out, 1);
                if (pos < CONE_SMEM_CAP) sout[pos] = make_int2(m, tm);
                else if (pos - CONE_SMEM_CAP < cap) gout[pos - CONE_SMEM_CAP] = m;
                else counters[5] = 1;
            });
        }
        __syncthreads();
        const int n_out = s_out;
        __syncthreads();
        if (threadIdx.x == 0) { s_in = min(n_out, cap); s_out = 0; }
        in_smem = min(n_out, CONE_SMEM_CAP);
        int2 *ts = sin_; sin_ = sout; sout = ts;
        int *tg = gin; gin = gout; gout = tg;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_in != 0) counters[5] = 1;              // did not settle within max_rounds
        counters[7] = first_round + rounds;
    }
}

template <typename real, int D>
__global__ void cone_emit_kernel(Grid<D> g, const int *rank, const int *k_dev, const int *need_t, int *tickets, real *memo,
                                 int *counters, int cap_tickets) {
    constexpr int NN = Grid<D>::NN;
    const long long total = g.size();
    if (*k_dev >= total) return;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int t = need_t[c];
        if (t < 0) continue;
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            const long long a = g.nbr(c, j);
            if (a < 0) continue;
            const int r = rank[a];
            if (r > t) continue;
            const int pos = atomicAdd(&counters[4], 1);
            if (pos < cap_tickets) {
                tickets[pos] = r * NN + (j ^ 1);                   // seen from a: c is its neighbour j ^ 1
                reinterpret_cast<unsigned long long *>(memo)[c * NN + j] = MEMO_UNKNOWN;
            } else counters[5] = 1;
        }
    }
}

template <typename real, int D>
__global__ void truncate_sweep_list_kernel(Grid<D> g, const real *F, const real *cost, const int *rank, const int *order,
                                           real *out, real *memo, const int *tickets, int *counters, const int *k_dev, int cap_tickets) {
    const int lane = threadIdx.x & 31;
    if (counters[5]) return;                       // the dense form takes over
    const int k = *k_dev;
    if (k >= g.size()) return;
    const int n_list = min(counters[4], cap_tickets);
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[6], 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n_list) break;
        const int idx = base + lane;
        if (idx < n_list) replay_ticket<real, D>((long long)tickets[idx], g, F, cost, rank, order, k, out, memo, &counters[1]);
        __syncwarp();
    }
}

}  // namespace fmb
