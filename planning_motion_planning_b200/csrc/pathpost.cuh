// pathpost.cuh -- what the planner does to the traced paths, on the device.
//
//   stitch2d  Coupled_motion_planner.py:1232-1234  roverPath = resolution * (vstack(flipud(pathS), pathG[1:]) + 1)
//   post3d    Coupled_motion_planner.py:1641-1671  per-axis scaling, savgol_filter(., 11, 3) (scipy, mode='interp'),
//             shift back to the global frame, last row := the sample pose, linear resampling to m rows
//             (interp1d(range(n), g)(linspace(0, n - 1, m)) -- np.interp semantics: exact at the nodes)
//
// Batched: one block per path (pair), row counts read from the tracer's device-side counters, so a batch of
// queries hands K x D waypoints to the host instead of 30 002-row slabs.
#pragma once
#include "fm_common.cuh"

namespace fmb {

constexpr int SG_WIN = 11, SG_HALF = 5;

struct PathPost3DArgs {
    const double *paths;        // [np][cap][3] rows [x, y, z] in cell units (tracer output)
    const int *count;           // [np] rows written
    long long cap;
    int np, m;                  // paths, rows of every resampled output
    double scale[3], offset[3];
    const double *last;         // [np][3] the pose the last row is set to before resampling, or nullptr
    const double *W;            // [11][11] least-squares cubic projection of an 11-point window (row 5 = the filter)
    double *out;                // [np][m][3]
    int *status;                // [np] 0 ok, 1 = fewer than 11 rows (scipy raises ValueError)
};

// smoothed, shifted row i of path p (component d); g(j) = paths[j][d] * scale[d]
__device__ __forceinline__ double pp_smoothed(const PathPost3DArgs &A, const double *P, int n, int i, int d, const double *lastp) {
    if (lastp && i == n - 1) return lastp[d];
    const double s = A.scale[d];
    int w0, r;                                        // window start, row of W
    if (i < SG_HALF) { w0 = 0; r = i; }
    else if (i > n - 1 - SG_HALF) { w0 = n - SG_WIN; r = i - w0; }
    else { w0 = i - SG_HALF; r = SG_HALF; }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < SG_WIN; ++j) acc = __fma_rn(A.W[r * SG_WIN + j], __dmul_rn(P[(long long)(w0 + j) * 3 + d], s), acc);
    return __dadd_rn(acc, A.offset[d]);
}

__global__ void path_post3d_kernel(PathPost3DArgs A) {
    const int p = blockIdx.x;
    if (p >= A.np) return;
    const int n = A.count[p];
    const double *P = A.paths + (long long)p * A.cap * 3;
    double *O = A.out + (long long)p * A.m * 3;
    if (n < SG_WIN) {
        if (threadIdx.x == 0) A.status[p] = 1;
        for (int k = threadIdx.x; k < A.m * 3; k += blockDim.x) O[k] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (threadIdx.x == 0) A.status[p] = 0;
    const double *lastp = A.last ? A.last + (long long)p * 3 : nullptr;
    // np.linspace(0, n - 1, m): k * step, the last sample exactly n - 1
    const double step = A.m > 1 ? __ddiv_rn((double)(n - 1), (double)(A.m - 1)) : 0.0;
    for (int k = threadIdx.x; k < A.m; k += blockDim.x) {
        const double x = (k == A.m - 1 && A.m > 1) ? (double)(n - 1) : __dmul_rn((double)k, step);
        int j = (int)x;                               // xp = 0, 1, ..., n - 1
        if (j > n - 1) j = n - 1;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double yj = pp_smoothed(A, P, n, j, d, lastp);
            double v = yj;
            if (j < n - 1 && x != (double)j) {        // np.interp: slope * (x - xp[j]) + fp[j]
                const double slope = __dsub_rn(pp_smoothed(A, P, n, j + 1, d, lastp), yj);
                v = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, (double)j)), yj);
            }
            O[(long long)k * 3 + d] = v;
        }
    }
}

struct Stitch2DArgs {
    const double *pathS, *pathG;   // [np][cap][2]
    const int *countS, *countG;    // [np]
    long long cap;
    int np;
    double resolution;
    double *out;                   // [np][2 * cap][2]
    int *count_out;                // [np] = countS + countG - 1
};

__global__ void path_stitch2d_kernel(Stitch2DArgs A) {
    const int p = blockIdx.x;
    if (p >= A.np) return;
    const int nS = A.countS[p], nG = A.countG[p];
    const double *S = A.pathS + (long long)p * A.cap * 2, *G = A.pathG + (long long)p * A.cap * 2;
    double *O = A.out + (long long)p * 2 * A.cap * 2;
    const int total = nS + (nG > 0 ? nG - 1 : 0);
    if (threadIdx.x == 0) A.count_out[p] = total;
    for (int k = threadIdx.x; k < total * 2; k += blockDim.x) {
        const int row = k >> 1, d = k & 1;
        const double v = row < nS ? S[(long long)(nS - 1 - row) * 2 + d] : G[(long long)(row - nS + 1) * 2 + d];
        O[k] = __dmul_rn(A.resolution, __dadd_rn(v, 1.0));
    }
}

}  // namespace fmb
