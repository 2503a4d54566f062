// trace2d.cuh -- gradient-descent path extraction over a 2D field
// (replaces FastMarching.py:164-236 getPathGDM, :242-300 computeGradient,
//  :305-338 interpolatePoint).
//
// The reference allocates four full-size arrays per step and fills a 6x6 window
// with a Python loop (FastMarching.py:255-297); only the normalised gradients of
// the 2x2 nodes around the current point -- 12 field values -- ever reach the
// result.  Here one warp walks one path.  The sequential chain per step is
// bilinear -> hypot -> divide -> hypot -> divide, and the kernel is built around keeping
// everything else off it: the node gradients of a 16 x 16 block around the path sit in
// shared memory (recomputed by all lanes when the path leaves the block), the bilinear
// coefficients of the current cell in registers, the common step runs as one
// straight-line chain with the branch-free exact square root / division of
// fm_common.cuh, and the stop test of a waypoint is evaluated beside the next step.
// Whatever that fast step does not cover (NaN / inf, vanishing gradient, field edge,
// out-of-range operands) is redone by the reference-order step, branch for branch.
// All arithmetic is individually rounded fp64 in the reference's operation order,
// including its mis-normalised dy (:226-227), exact-zero special cases
// (:327-336) and the numpy>=2 behaviour of the NaN fallback (:178-218).
#pragma once
#include "fm_common.cuh"

namespace fmb {

enum : int { TR_OK = 0, TR_EARLY = 1, TR_VALUEERROR = 2, TR_INDEXERROR = 3, TR_OVERFLOW = 4 };

template <typename real>
struct TraceArgs2D {
    const real *T;
    long long T_pitch, T_qstride;
    int rows, cols, npaths;
    const int *field_of_path;
    const double *init, *end;
    double tau;
    int max_steps;
    double *out;
    long long cap;
    int *count, *status;
    // optional log of the field cells a path read (fmb_trace2d_logged_f64): origin node (bx, by) of every gradient block
    // it staged -- all reads of a path lie in the (TB + 2)^2 windows of its blocks.  blocks [npaths][cap_blocks][2],
    // nblocks [npaths] = number of blocks, cap_blocks + 1 when the log overflowed, -1 when the field is too small for blocks
    int *blocks = nullptr, *nblocks = nullptr;
    int cap_blocks = 0;
};

#ifndef FMB_HOST_EMU
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
inline void prefetch_l1(const void *) {}
#endif
__device__ __forceinline__ double dsq(double a) { return __dmul_rn(a, a); }
__device__ __forceinline__ double dhyp2(double a, double b) { return __dsqrt_rn(__dadd_rn(dsq(a), dsq(b))); }
__device__ __forceinline__ bool d_isinf(double v) { return fabs(v) == __longlong_as_double(0x7ff0000000000000LL); }

// FastMarching.py:262-297 for node (i = x, j = y) of an m x n field, from the node's value and its four neighbours
// (at the field edge the missing neighbour is not read by the edge form, so any value may stand in for it)
__device__ __forceinline__ void grad_from5(double c, double cu, double cd, double cl, double cr, int m, int n, int i, int j,
                                           double &gnx, double &gny) {
    double Gx, Gy;
    if (j == 0) Gy = __dsub_rn(cd, c);
    else if (j == m - 1) Gy = __dsub_rn(c, cu);
    else if (d_isinf(cd)) Gy = d_isinf(cu) ? 0.0 : __dsub_rn(c, cu);
    else Gy = d_isinf(cu) ? __dsub_rn(cd, c) : __dmul_rn(__dsub_rn(cd, cu), 0.5);
    if (i == 0) Gx = __dsub_rn(cr, c);
    else if (i == n - 1) Gx = __dsub_rn(c, cl);
    else if (d_isinf(cr)) Gx = d_isinf(cl) ? 0.0 : __dsub_rn(c, cl);
    else Gx = d_isinf(cl) ? __dsub_rn(cr, c) : __dmul_rn(__dsub_rn(cr, cl), 0.5);
    const double nrm = dhyp2(Gx, Gy);
    gnx = __ddiv_rn(Gx, nrm);
    gny = __ddiv_rn(Gy, nrm);
}
template <typename real>
__device__ __forceinline__ void grad_node2d(const real *T, long long pitch, int m, int n, int i, int j,
                                            double &gnx, double &gny) {
    const int jm = max(j - 1, 0), jp = min(j + 1, m - 1), im = max(i - 1, 0), ip = min(i + 1, n - 1);
    const double c = (double)T[(long long)j * pitch + i];
    const double cu = (double)T[(long long)jm * pitch + i], cd = (double)T[(long long)jp * pitch + i];
    const double cl = (double)T[(long long)j * pitch + im], cr = (double)T[(long long)j * pitch + ip];
    grad_from5(c, cu, cd, cl, cr, m, n, i, j, gnx, gny);
}

// FastMarching.py:323-336
__device__ __forceinline__ double bilinear_ref(double m00, double m01, double m10, double m11, double a, double b) {
    const double a00 = m00;
    const double a10 = __dsub_rn(m01, m00);
    const double a01 = __dsub_rn(m10, m00);
    const double a11 = __dsub_rn(__dsub_rn(__dadd_rn(m11, m00), m01), m10);
    if (a == 0.0) {
        if (b == 0.0) return a00;
        return __dadd_rn(a00, __dmul_rn(a01, b));
    }
    if (b == 0.0) return __dadd_rn(a00, __dmul_rn(a10, a));
    return __dadd_rn(__dadd_rn(__dadd_rn(a00, __dmul_rn(a10, a)), __dmul_rn(a01, b)), __dmul_rn(__dmul_rn(a11, a), b));
}

// Gradient block of the fast step: the normalised gradients of TB x TB nodes around the path live in shared memory
// (one block per warp), computed -- all lanes at once, from a (TB + 2)^2 window of T staged in shared memory with one
// round of loads -- only when the path leaves the block (every ~30 steps); entering another cell inside the block costs
// four 16-byte shared-memory loads.  Same values as the reference's per-step recomputation (FastMarching.py:255-297).
constexpr int TB = 16, TWIN = TB + 2;
constexpr int TRACE2D_SMEM_PER_WARP = TB * TB * 16 + TWIN * TWIN * 8;

// (Re)computes the gradient block whose origin node is (bx, by): the (TB + 2)^2 window of T into sW with every load in
// flight before the first store, then eight nodes per lane, four at a time (the branch-free exact square root /
// division let their chains interleave; a node outside their range -- zero or infinite gradient -- is redone by
// grad_from5).  Not inlined: it runs once per ~30 steps and its registers must not weigh on the step loop.
template <typename real>
__device__ __noinline__ void load_gradient_block(const real *T, long long pitch, int m, int n, int bx, int by, double *sW,
                                                 double2 *sG, int lane) {
    __syncwarp();
    {   // the window: every load in flight before the first store (one memory latency for all of it)
        constexpr int NW = (TWIN * TWIN + 31) / 32;
        double v[NW];
#pragma unroll
        for (int u = 0; u < NW; ++u) {
            const int t = min(lane + 32 * u, TWIN * TWIN - 1);
            const int wy = t / TWIN, wx = t - wy * TWIN;
            const int gy = min(max(by - 1 + wy, 0), m - 1), gx = min(max(bx - 1 + wx, 0), n - 1);
            v[u] = (double)T[(long long)gy * pitch + gx];
        }
#pragma unroll
        for (int u = 0; u < NW; ++u) sW[min(lane + 32 * u, TWIN * TWIN - 1)] = v[u];
    }
    __syncwarp();
    // eight nodes per lane, four at a time: the branch-free exact square root / division let their chains
    // interleave; a node outside their range (zero or infinite gradient) is redone by grad_from5
#pragma unroll
    for (int t0 = 0; t0 < TB * TB; t0 += 128) {
        double gxs[4], gys[4];
        bool oks[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + 32 * u + lane;
            const int ny_ = t / TB, nx_ = t - ny_ * TB;
            const double *w = sW + (ny_ + 1) * TWIN + nx_ + 1;
            const int i = bx + nx_, j = by + ny_;
            const double c = w[0], cu = w[-TWIN], cd = w[TWIN], cl = w[-1], cr = w[1];
            const bool iu = d_isinf(cu), id = d_isinf(cd), il = d_isinf(cl), ir = d_isinf(cr);
            double Gy = id ? (iu ? 0.0 : __dsub_rn(c, cu)) : (iu ? __dsub_rn(cd, c) : __dmul_rn(__dsub_rn(cd, cu), 0.5));
            Gy = j == 0 ? __dsub_rn(cd, c) : (j == m - 1 ? __dsub_rn(c, cu) : Gy);
            double Gx = ir ? (il ? 0.0 : __dsub_rn(c, cl)) : (il ? __dsub_rn(cr, c) : __dmul_rn(__dsub_rn(cr, cl), 0.5));
            Gx = i == 0 ? __dsub_rn(cr, c) : (i == n - 1 ? __dsub_rn(c, cl) : Gx);
            const double s2 = __dadd_rn(dsq(Gx), dsq(Gy));
            bool ok = sqrt_fast_ok(s2);
            const double nrm = sqrt_rn_fast(ok ? s2 : 1.0);
            gxs[u] = ddiv_rn_fast(Gx, nrm, ok);
            gys[u] = ddiv_rn_fast(Gy, nrm, ok);
            oks[u] = ok;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + 32 * u + lane;
            if (!oks[u]) {
                const int ny_ = t / TB, nx_ = t - ny_ * TB;
                const double *w = sW + (ny_ + 1) * TWIN + nx_ + 1;
                grad_from5(w[0], w[-TWIN], w[TWIN], w[-1], w[1], m, n, bx + nx_, by + ny_, gxs[u], gys[u]);
            }
            sG[t] = make_double2(gxs[u], gys[u]);
        }
    }
    __syncwarp();
}

template <typename real, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) trace2d_kernel(TraceArgs2D<real> A) {
    FMB_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31;
    double2 *sG = reinterpret_cast<double2 *>(smem_raw + (size_t)(threadIdx.x >> 5) * TRACE2D_SMEM_PER_WARP);
    double *sW = reinterpret_cast<double *>(sG + TB * TB);
    const int p = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (p >= A.npaths) return;
    const int f = A.field_of_path ? A.field_of_path[p] : p;
    const real *T = A.T + (long long)f * A.T_qstride;
    const int m = A.rows, n = A.cols;
    double *out = A.out + (long long)p * A.cap * 2;
    const double ex = A.end[2 * p], ey = A.end[2 * p + 1];
    double px = A.init[2 * p], py = A.init[2 * p + 1];
    long long K = 0;
    int status = TR_OK;
    bool append_end = true;
    if (lane == 0) { out[0] = px; out[1] = py; }
    K = 1;
    int ci = -1, cj = -1;                                   // reference-order step: cell whose node gradients are cached
    double x00 = 0, x01 = 0, x10 = 0, x11 = 0, y00 = 0, y01 = 0, y10 = 0, y11 = 0;
    const bool blocked = n >= TB && m >= TB;
    int bx = 0, by = 0;                                     // origin node of the block
    double bxlo = 1.0, bxhi = 0.0, bylo = 1.0, byhi = 0.0;  // cells of the block as doubles (empty: no block yet)
    double lnx = 0.0, lny = 0.0;                            // last step direction (block placement)
    double cfi = -1.0, cfj = -1.0;                          // fast step: cell whose bilinear coefficients are cached
    double fx0 = 0, fx1 = 0, fx2 = 0, fx3 = 0, fy0 = 0, fy1 = 0, fy2 = 0, fy3 = 0;
    const double nd = (double)n, md = (double)m;
    int n_logged = blocked ? 0 : -1;

    for (int step = 0; step < A.max_steps; ++step) {
        double nx = 0.0, ny = 0.0;
        // the stop test of the waypoint appended by the previous step (:231), evaluated beside this step's arithmetic
        // instead of between two steps.  sqrt(s) < 1.5 <=> s < 2.25 exactly for a correctly rounded sqrt.
        const bool arrived = step > 0 && __dadd_rn(dsq(__dsub_rn(px, ex)), dsq(__dsub_rn(py, ey))) < 2.25;
        // ---- fast step: the common case as ONE straight-line dependent chain.  Every exceptional condition (waypoint
        // outside the field, NaN / inf anywhere, a vanishing gradient, operands outside the range of the branch-free
        // exact square root / division) only clears `fast`, and the step is then redone by the reference-order code
        // below -- which is also what runs under every branch of FastMarching.py:178-227.  In range the branch-free
        // forms return the bits of sqrt.rn / div.rn, and the general bilinear form equals the reference's exact-zero
        // forms whenever no NaN arises (x + 0*y == x for finite y; a NaN falls back).
        bool fast = false;
        if (blocked) {
            const double fi = trunc(px), fj = trunc(py);
            bool have = true;
            if (fi != cfi || fj != cfj) {                   // another cell (or the first step, or NaN)
                cfi = cfj = -1.0;
                bool inb = fi >= bxlo && fi <= bxhi && fj >= bylo && fj <= byhi;          // NaN -> false
                if (!inb && fi >= 0.0 && fj >= 0.0 && fi + 1.0 < nd && fj + 1.0 < md) {
                    // the path walks DOWN the gradient (p -= tau * n): it heads towards -lnx, -lny; the block is laid
                    // out ahead of it, the current cell one cell inside the trailing edge
                    const int i = (int)fi, j = (int)fj;
                    bx = min(max(i - (lnx > 0.0 ? TB - 3 : (lnx < 0.0 ? 1 : TB / 2 - 1)), 0), n - TB);
                    by = min(max(j - (lny > 0.0 ? TB - 3 : (lny < 0.0 ? 1 : TB / 2 - 1)), 0), m - TB);
                    load_gradient_block<real>(T, A.T_pitch, m, n, bx, by, sW, sG, lane);
                    if (A.blocks) {
                        if (n_logged < A.cap_blocks) {
                            if (lane == 0) { int *e = A.blocks + ((long long)p * A.cap_blocks + n_logged) * 2; e[0] = bx; e[1] = by; }
                            ++n_logged;
                        } else n_logged = A.cap_blocks + 1;
                    }
                    bxlo = (double)bx; bxhi = (double)(bx + TB - 2); bylo = (double)by; byhi = (double)(by + TB - 2);
                    inb = true;
                }
                if (inb) {
                    const int l00 = ((int)fj - by) * TB + ((int)fi - bx);
                    const double2 g00 = sG[l00], g01 = sG[l00 + 1], g10 = sG[l00 + TB], g11 = sG[l00 + TB + 1];
                    fx0 = g00.x; fx1 = __dsub_rn(g01.x, g00.x); fx2 = __dsub_rn(g10.x, g00.x);
                    fx3 = __dsub_rn(__dsub_rn(__dadd_rn(g11.x, g00.x), g01.x), g10.x);
                    fy0 = g00.y; fy1 = __dsub_rn(g01.y, g00.y); fy2 = __dsub_rn(g10.y, g00.y);
                    fy3 = __dsub_rn(__dsub_rn(__dadd_rn(g11.y, g00.y), g01.y), g10.y);
                    cfi = fi; cfj = fj;
                }
                have = inb;
            }
            if (have) {
                const double a = __dsub_rn(px, fi), b = __dsub_rn(py, fj);
                const double dx = __dadd_rn(__dadd_rn(__dadd_rn(fx0, __dmul_rn(fx1, a)), __dmul_rn(fx2, b)), __dmul_rn(__dmul_rn(fx3, a), b));
                const double dy = __dadd_rn(__dadd_rn(__dadd_rn(fy0, __dmul_rn(fy1, a)), __dmul_rn(fy2, b)), __dmul_rn(__dmul_rn(fy3, a), b));
                const double dy2 = dsq(dy);
                const double s1 = __dadd_rn(dsq(dx), dy2);
                bool ok = !isnan(dx) && !isnan(dy) && sqrt_fast_ok(s1);
                const double h1 = sqrt_rn_fast(ok ? s1 : 1.0);
                ok = ok && h1 >= 0.01;
                const double fnx = ddiv_rn_fast(dx, h1, ok);
                const double s2 = __dadd_rn(dsq(fnx), dy2);
                ok = ok && sqrt_fast_ok(s2);
                const double h2 = sqrt_rn_fast(ok ? s2 : 1.0);
                const double fny = ddiv_rn_fast(dy, h2, ok);           // sic (:226-227): dy over hypot(nx, dy)
                fast = ok;
                nx = fnx; ny = fny;
            }
        }
        if (arrived) break;
        if (!fast) {
            if (isnan(px) || isnan(py)) { status = TR_VALUEERROR; append_end = false; break; }   // int(nan) at :250
            if (d_isinf(px) || d_isinf(py)) { status = TR_OVERFLOW; append_end = false; break; }
            const double fi = trunc(px), fj = trunc(py);
            if (!(fi >= 0.0) || !(fj >= 0.0) || fi + 1.0 >= (double)n || fj + 1.0 >= (double)m) {
                status = TR_INDEXERROR; append_end = false; break;
            }
            const int i = (int)fi, j = (int)fj;
            const double a = __dsub_rn(px, (double)i), b = __dsub_rn(py, (double)j);
            if (i != ci || j != cj) {
                ci = i; cj = j;
                double gx, gy;
                grad_node2d<real>(T, A.T_pitch, m, n, i + (lane & 1), j + ((lane >> 1) & 1), gx, gy);
                x00 = __shfl_sync(FULL, gx, 0); x01 = __shfl_sync(FULL, gx, 1);
                x10 = __shfl_sync(FULL, gx, 2); x11 = __shfl_sync(FULL, gx, 3);
                y00 = __shfl_sync(FULL, gy, 0); y01 = __shfl_sync(FULL, gy, 1);
                y10 = __shfl_sync(FULL, gy, 2); y11 = __shfl_sync(FULL, gy, 3);
            }
            const double dx = bilinear_ref(x00, x01, x10, x11, a, b);
            const double dy = bilinear_ref(y00, y01, y10, y11, a, b);

            if (isnan(dx) || isnan(dy)) {
                // :178-218 under numpy >= 2: prune, append the nearest node, then the child scan
                // raises inside the bare `except` and the path so far is returned without `end`.
                status = TR_EARLY; append_end = false;
                if (lane == 0) {
                    double n0 = rint(px), n1 = rint(py);
                    bool ok = true;
                    for (;;) {
                        if (n0 < 0 || n1 < 0 || n0 >= n || n1 >= m) { ok = false; break; }
                        if (!d_isinf((double)T[(long long)n1 * A.T_pitch + (long long)n0])) break;
                        --K;
                        if (K == 0) { ok = false; break; }
                        n0 = rint(out[2 * (K - 1)]); n1 = rint(out[2 * (K - 1) + 1]);
                    }
                    if (ok) {
                        while (K > 0) {
                            const double qx = __dsub_rn(out[2 * (K - 1)], n0), qy = __dsub_rn(out[2 * (K - 1) + 1], n1);
                            if (!(dhyp2(qx, qy) < 1.0)) break;
                            --K;
                        }
                        out[2 * K] = n0; out[2 * K + 1] = n1; ++K;
                    }
                }
                K = __shfl_sync(FULL, K, 0);
                break;
            }

            if (dhyp2(dx, dy) < 0.01) {
                const double s = dhyp2(dx, dy);
                nx = __ddiv_rn(dx, s); ny = __ddiv_rn(dy, s);
            } else {
                nx = __ddiv_rn(dx, dhyp2(dx, dy));
                ny = __ddiv_rn(dy, dhyp2(nx, dy));          // sic (:226-227)
            }
        }
        lnx = nx; lny = ny;
        px = __dsub_rn(px, __dmul_rn(A.tau, nx));
        py = __dsub_rn(py, __dmul_rn(A.tau, ny));
        if (lane == 0) { out[2 * K] = px; out[2 * K + 1] = py; }
        ++K;
    }
    if (append_end) {
        if (lane == 0) { out[2 * K] = ex; out[2 * K + 1] = ey; }
        ++K;
    }
    if (lane == 0) { A.count[p] = (int)K; A.status[p] = status; if (A.nblocks) A.nblocks[p] = n_logged; }
}

// Bitwise comparison of a traced device field with a host array over the windows a path read (see TraceArgs2D::blocks):
// one CTA per logged block.  h is DEVICE-ACCESSIBLE host memory (page-locked, unified addressing): the kernel reads the
// ~(TB + 2)^2 values of a window over the bus, nothing else of the array moves.  Element (y, x) of the field sits at
// h[y * hs_y + x * hs_x].  *flag |= 1 on any difference, on a log that overflowed or is missing.
__global__ void windows_differ_kernel(const double *field, long long pitch, int rows, int cols, const double *h, long long hs_y,
                                      long long hs_x, const int *blocks, const int *nblocks, int cap_blocks, int *flag) {
    const int nb = nblocks[0];
    if (nb < 0 || nb > cap_blocks) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(flag, 1); return; }
    int diff = 0;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int bx = blocks[2 * b], by = blocks[2 * b + 1];
        for (int t = threadIdx.x; t < TWIN * TWIN; t += blockDim.x) {
            const int wy = t / TWIN, wx = t - wy * TWIN;
            const int y = min(max(by - 1 + wy, 0), rows - 1), x = min(max(bx - 1 + wx, 0), cols - 1);
            const unsigned long long a = reinterpret_cast<const unsigned long long *>(field)[(long long)y * pitch + x];
            const unsigned long long c = reinterpret_cast<const unsigned long long *>(h)[(long long)y * hs_y + (long long)x * hs_x];
            diff |= a != c;
        }
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicExch(flag, 1);
}

}  // namespace fmb
