// trace3d.cuh -- gradient-descent path extraction over a 3D field
// (replaces FastMarching3D.py:198-271 getPathGDM and :275-314 interpolatePoint).
//
// The reference runs np.gradient over the WHOLE volume (three full-size arrays,
// FastMarching3D.py:200) and then reads 8 nodes x 3 components per step.  Here the
// gradient is evaluated only where it is read: one warp walks one path, lanes 0..7
// each produce np.gradient's three components at one corner node of the current
// cell (unit spacing, one-sided at the array faces, NOT inf-aware -- exactly
// np.gradient's edge_order=1 arithmetic), the 24 values are exchanged by shuffle
// and every lane advances the warp-uniform position with the reference's
// tri-linear-like interpolant (including its non-standard a7, :290), raw
// un-normalised steps (:262-264) and the snap-to-node fallback (:212-253).
#pragma once
#include "fm_common.cuh"
#include "trace2d.cuh"

namespace fmb {

template <typename real>
struct TraceArgs3D {
    const real *T;
    long long T_qstride;
    int ny, nx, nz, npaths;
    const int *field_of_path;
    const double *init, *end;
    double tau;
    int max_steps;
    double *out;
    long long cap;
    int *count, *status;
};

// np.gradient(T) along one axis at node (y,x,z): central difference inside,
// first-order one-sided on the two faces.
template <typename real>
__device__ __forceinline__ double npgrad_axis(const real *T, long long stride, int p, int len, long long center) {
    if (p == 0) return __dsub_rn((double)T[center + stride], (double)T[center]);
    if (p == len - 1) return __dsub_rn((double)T[center], (double)T[center - stride]);
    return __dmul_rn(__dsub_rn((double)T[center + stride], (double)T[center - stride]), 0.5);
}

// FastMarching3D.py:283-312 (generic branch)
__device__ __forceinline__ double trilinear_ref(double m000, double m010, double m100, double m001, double m110,
                                                double m011, double m101, double m111, double a, double b, double c) {
    // naming: m<j><i><k> = mapI[j + dj, i + di, k + dk]
    const double a0 = m000;
    const double a1 = __dsub_rn(m010, m000);
    const double a2 = __dsub_rn(m100, m000);
    const double a3 = __dsub_rn(m001, m000);
    const double a4 = __dsub_rn(__dsub_rn(__dadd_rn(m110, m000), m010), m100);
    const double a5 = __dsub_rn(__dsub_rn(__dadd_rn(m011, m000), m010), m001);
    const double a6 = __dsub_rn(__dsub_rn(__dadd_rn(m101, m000), m100), m001);
    const double a7 = __dsub_rn(__dsub_rn(__dsub_rn(__dadd_rn(m111, m000), m100), m001), m010);   // sic (:290)
    double r = __dadd_rn(a0, __dmul_rn(a1, a));
    r = __dadd_rn(r, __dmul_rn(a2, b));
    r = __dadd_rn(r, __dmul_rn(a3, c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(a4, a), b));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(a5, a), c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(a6, b), c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(a7, a), b), c));
    return r;
}

__device__ __forceinline__ double dhyp3(double a, double b, double c) {
    return __dsqrt_rn(__dadd_rn(__dadd_rn(dsq(a), dsq(b)), dsq(c)));
}

template <typename real, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) trace3d_kernel(TraceArgs3D<real> A) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (p >= A.npaths) return;
    const int f = A.field_of_path ? A.field_of_path[p] : p;
    const real *T = A.T + (long long)f * A.T_qstride;
    const int ny = A.ny, nx = A.nx, nz = A.nz;
    const long long sy = (long long)nx * nz, sx = nz;
    double *out = A.out + (long long)p * A.cap * 3;
    const double ex = A.end[3 * p], ey = A.end[3 * p + 1], ez = A.end[3 * p + 2];
    double px = A.init[3 * p], py = A.init[3 * p + 1], pz = A.init[3 * p + 2];
    long long K = 1;
    int status = TR_OK;
    bool append_end = true;
    if (lane == 0) { out[0] = px; out[1] = py; out[2] = pz; }

    for (int step = 0; step < A.max_steps; ++step) {
        // a non-finite waypoint ends in int(round(.)) of the fallback (:213-215): x, y, z in turn
        {
            const double qv[3] = {px, py, pz};
            int bad = 0;
#pragma unroll
            for (int c = 0; c < 3 && !bad; ++c) {
                if (d_isinf(qv[c])) bad = TR_OVERFLOW;
                else if (isnan(qv[c])) bad = TR_VALUEERROR;
            }
            if (bad) { status = bad; append_end = false; break; }
        }
        const double fi = trunc(px), fj = trunc(py), fk = trunc(pz);
        if (!(fi >= 0.0) || !(fj >= 0.0) || !(fk >= 0.0) || fi + 1.0 >= (double)nx || fj + 1.0 >= (double)ny || fk + 1.0 >= (double)nz) {
            status = TR_INDEXERROR; append_end = false; break;
        }
        const int i = (int)fi, j = (int)fj, k = (int)fk;
        const double a = __dsub_rn(px, (double)i), b = __dsub_rn(py, (double)j), c = __dsub_rn(pz, (double)k);
        // lane bits: bit0 = di (x), bit1 = dj (y), bit2 = dk (z)
        const int xi = i + (lane & 1), yj = j + ((lane >> 1) & 1), zk = k + ((lane >> 2) & 1);
        const long long center = (long long)yj * sy + (long long)xi * sx + zk;
        const double g1 = npgrad_axis<real>(T, sx, xi, nx, center);   // d/dx  (G1, axis 1)
        const double g2 = npgrad_axis<real>(T, sy, yj, ny, center);   // d/dy  (G2, axis 0)
        const double g3 = npgrad_axis<real>(T, 1, zk, nz, center);    // d/dz  (G3, axis 2)
        // gather the 8 corners: index = di + 2*dj + 4*dk
#define CORNERS(g) __shfl_sync(FULL, g, 0), __shfl_sync(FULL, g, 1), __shfl_sync(FULL, g, 2), __shfl_sync(FULL, g, 4), \
                   __shfl_sync(FULL, g, 3), __shfl_sync(FULL, g, 5), __shfl_sync(FULL, g, 6), __shfl_sync(FULL, g, 7)
        // trilinear_ref order: m000, m010(di), m100(dj), m001(dk), m110(di+dj), m011(di+dk), m101(dj+dk), m111
        double dx = trilinear_ref(CORNERS(g1), a, b, c);
        double dy = trilinear_ref(CORNERS(g2), a, b, c);
        double dz = trilinear_ref(CORNERS(g3), a, b, c);
#undef CORNERS

        if (isnan(dx) || isnan(dy) || isnan(dz)) {
            // :212-253 snap to the nearest node, prune, step towards the lowest face neighbour
            int st = TR_OK;
            double ndx = dx, ndy = dy, ndz = dz, npx = px, npy = py, npz = pz;
            if (lane == 0) {
                long long n0 = (long long)rint(px), n1 = (long long)rint(py), n2 = (long long)rint(pz);
                for (;;) {
                    if (n0 < 0 || n1 < 0 || n2 < 0 || n0 >= nx || n1 >= ny || n2 >= nz) { st = TR_INDEXERROR; break; }
                    if (!d_isinf((double)T[n1 * sy + n0 * sx + n2])) break;
                    --K;
                    if (K == 0) { st = TR_INDEXERROR; break; }
                    n0 = (long long)rint(out[3 * (K - 1)]); n1 = (long long)rint(out[3 * (K - 1) + 1]); n2 = (long long)rint(out[3 * (K - 1) + 2]);
                }
                if (st == TR_OK) {
                    while (K > 0) {
                        const double qx = __dsub_rn(out[3 * (K - 1)], (double)n0), qy = __dsub_rn(out[3 * (K - 1) + 1], (double)n1),
                                     qz = __dsub_rn(out[3 * (K - 1) + 2], (double)n2);
                        if (!(dhyp3(qx, qy, qz) < 1.0)) break;
                        --K;
                    }
                    out[3 * K] = (double)n0; out[3 * K + 1] = (double)n1; out[3 * K + 2] = (double)n2; ++K;
                    npx = (double)n0; npy = (double)n1; npz = (double)n2;
                    double currentT = (double)T[n1 * sy + n0 * sx + n2];
                    const int ch[6][3] = {{0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 0, -1}, {0, 0, 1}};
                    for (int s = 0; s < 6; ++s) {
                        const long long c0 = n0 + ch[s][0], c1 = n1 + ch[s][1], c2 = n2 + ch[s][2];
                        if (c0 >= nx || c1 >= ny || c2 >= nz) { st = TR_INDEXERROR; break; }
                        const long long w0 = c0 < 0 ? c0 + nx : c0, w1 = c1 < 0 ? c1 + ny : c1, w2 = c2 < 0 ? c2 + nz : c2;   // python wrap
                        const double tv = (double)T[w1 * sy + w0 * sx + w2];
                        if (tv < currentT) {
                            currentT = tv;
                            ndx = __ddiv_rn((double)(n0 - c0), A.tau); ndy = __ddiv_rn((double)(n1 - c1), A.tau); ndz = __ddiv_rn((double)(n2 - c2), A.tau);
                        }
                    }
                }
            }
            st = __shfl_sync(FULL, st, 0);
            K = __shfl_sync(FULL, K, 0);
            if (st != TR_OK) { status = st; append_end = false; break; }
            dx = __shfl_sync(FULL, ndx, 0); dy = __shfl_sync(FULL, ndy, 0); dz = __shfl_sync(FULL, ndz, 0);
            px = __shfl_sync(FULL, npx, 0); py = __shfl_sync(FULL, npy, 0); pz = __shfl_sync(FULL, npz, 0);
        }

        const double norm = dhyp3(dx, dy, dz);
        if (norm < 0.01) {
            px = __dsub_rn(px, __dmul_rn(A.tau, __ddiv_rn(dx, norm)));
            py = __dsub_rn(py, __dmul_rn(A.tau, __ddiv_rn(dy, norm)));
            pz = __dsub_rn(pz, __dmul_rn(A.tau, __ddiv_rn(dz, norm)));
        } else {
            px = __dsub_rn(px, __dmul_rn(A.tau, dx));
            py = __dsub_rn(py, __dmul_rn(A.tau, dy));
            pz = __dsub_rn(pz, __dmul_rn(A.tau, dz));
        }
        if (lane == 0) { out[3 * K] = px; out[3 * K + 1] = py; out[3 * K + 2] = pz; }
        ++K;
        if (dhyp3(__dsub_rn(px, ex), __dsub_rn(py, ey), __dsub_rn(pz, ez)) < 1.5) break;
    }
    if (append_end) {
        if (lane == 0) { out[3 * K] = ex; out[3 * K + 1] = ey; out[3 * K + 2] = ez; }
        ++K;
    }
    if (lane == 0) { A.count[p] = (int)K; A.status[p] = status; }
}

}  // namespace fmb
