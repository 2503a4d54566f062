// trace2d.cuh -- gradient-descent path extraction over a 2D field
// (replaces FastMarching.py:164-236 getPathGDM, :242-300 computeGradient,
//  :305-338 interpolatePoint).
//
// The reference allocates four full-size arrays per step and fills a 6x6 window
// with a Python loop (FastMarching.py:255-297); only the normalised gradients of
// the 2x2 nodes around the current point -- 12 field values -- ever reach the
// result.  Here one warp walks one path: lanes 0..3 each evaluate one node's
// inf-aware normalised gradient from five cached loads, the four results are
// exchanged by shuffle, and every lane advances the (warp-uniform) position, so
// the sequential chain per step is load -> gradient -> bilinear -> normalise.
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
};

#ifndef FMB_HOST_EMU
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
inline void prefetch_l1(const void *) {}
#endif
__device__ __forceinline__ double dsq(double a) { return __dmul_rn(a, a); }
__device__ __forceinline__ double dhyp2(double a, double b) { return __dsqrt_rn(__dadd_rn(dsq(a), dsq(b))); }
__device__ __forceinline__ bool d_isinf(double v) { return fabs(v) == __longlong_as_double(0x7ff0000000000000LL); }

// FastMarching.py:262-297 for node (i = x, j = y) of an m x n field
template <typename real>
__device__ __forceinline__ void grad_node2d(const real *T, long long pitch, int m, int n, int i, int j,
                                            double &gnx, double &gny) {
    const int jm = max(j - 1, 0), jp = min(j + 1, m - 1), im = max(i - 1, 0), ip = min(i + 1, n - 1);
    const double c = (double)T[(long long)j * pitch + i];
    const double cu = (double)T[(long long)jm * pitch + i], cd = (double)T[(long long)jp * pitch + i];
    const double cl = (double)T[(long long)j * pitch + im], cr = (double)T[(long long)j * pitch + ip];
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

template <typename real, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) trace2d_kernel(TraceArgs2D<real> A) {
    const int lane = threadIdx.x & 31;
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
    int ci = -1, cj = -1;                                   // cell whose node gradients are cached
    double x00 = 0, x01 = 0, x10 = 0, x11 = 0, y00 = 0, y01 = 0, y10 = 0, y11 = 0;
    // Node gradients are a pure function of T and of the node, and the path moves at most tau per step: the warp keeps
    // the normalised gradients of a BLOCK of BW x BH nodes in registers (lane = node), recomputed -- all lanes at once,
    // one round of loads -- only when the path leaves the block; entering another cell inside the block costs eight
    // shuffles and no load.  Same values as the reference's per-step recomputation (FastMarching.py:255-297).
    constexpr int BW = 6, BH = 5;                           // 30 nodes = 5 x 4 cells
    const bool blocked = n >= BW && m >= BH;
    int bx = 0, by = -1000000;                              // origin node of the block (none yet)
    double gbx = 0.0, gby = 0.0;                            // my node's gradient
    double lnx = 0.0, lny = 0.0;                            // last step direction (block placement)

    for (int step = 0; step < A.max_steps; ++step) {
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
            if (blocked) {
                if (i < bx || i > bx + BW - 2 || j < by || j > by + BH - 2) {
                    // the path walks DOWN the gradient: p -= tau * n, so it heads towards -lnx, -lny
                    bx = min(max(i - (lnx > 0.0 ? 3 : (lnx < 0.0 ? 1 : 2)), 0), n - BW);
                    by = min(max(j - (lny > 0.0 ? 2 : 1), 0), m - BH);
                    if (lane < BW * BH) grad_node2d<real>(T, A.T_pitch, m, n, bx + lane % BW, by + lane / BW, gbx, gby);
                    else {                                  // the two idle lanes pull the rows ahead into L1
                        const int pj = min(max(lny > 0.0 ? by - 3 - (lane & 1) : by + BH + 2 + (lane & 1), 0), m - 1);
                        prefetch_l1(&T[(long long)pj * A.T_pitch + i]);
                    }
                }
                const int l00 = (j - by) * BW + (i - bx);
                x00 = __shfl_sync(FULL, gbx, l00); x01 = __shfl_sync(FULL, gbx, l00 + 1);
                x10 = __shfl_sync(FULL, gbx, l00 + BW); x11 = __shfl_sync(FULL, gbx, l00 + BW + 1);
                y00 = __shfl_sync(FULL, gby, l00); y01 = __shfl_sync(FULL, gby, l00 + 1);
                y10 = __shfl_sync(FULL, gby, l00 + BW); y11 = __shfl_sync(FULL, gby, l00 + BW + 1);
            } else {
                double gx, gy;
                grad_node2d<real>(T, A.T_pitch, m, n, i + (lane & 1), j + ((lane >> 1) & 1), gx, gy);
                x00 = __shfl_sync(FULL, gx, 0); x01 = __shfl_sync(FULL, gx, 1);
                x10 = __shfl_sync(FULL, gx, 2); x11 = __shfl_sync(FULL, gx, 3);
                y00 = __shfl_sync(FULL, gy, 0); y01 = __shfl_sync(FULL, gy, 1);
                y10 = __shfl_sync(FULL, gy, 2); y11 = __shfl_sync(FULL, gy, 3);
            }
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

        double nx, ny;
        if (dhyp2(dx, dy) < 0.01) {
            const double s = dhyp2(dx, dy);
            nx = __ddiv_rn(dx, s); ny = __ddiv_rn(dy, s);
        } else {
            nx = __ddiv_rn(dx, dhyp2(dx, dy));
            ny = __ddiv_rn(dy, dhyp2(nx, dy));          // sic (:226-227)
        }
        lnx = nx; lny = ny;
        px = __dsub_rn(px, __dmul_rn(A.tau, nx));
        py = __dsub_rn(py, __dmul_rn(A.tau, ny));
        if (lane == 0) { out[2 * K] = px; out[2 * K + 1] = py; }
        ++K;
        // stop within 1.5 cells of `end` (:231).  sqrt(s) < 1.5 <=> s < 2.25 exactly for a correctly
        // rounded sqrt (1.5^2 is representable), which keeps the square root off the step chain.
        if (__dadd_rn(dsq(__dsub_rn(px, ex)), dsq(__dsub_rn(py, ey))) < 2.25) break;
    }
    if (append_end) {
        if (lane == 0) { out[2 * K] = ex; out[2 * K + 1] = ey; }
        ++K;
    }
    if (lane == 0) { A.count[p] = (int)K; A.status[p] = status; }
}

}  // namespace fmb
