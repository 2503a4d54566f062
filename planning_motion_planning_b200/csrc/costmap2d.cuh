// costmap2d.cuh -- the planner's 2D cost-map construction on the device:
// DEM -> slope obstacles -> hole filling / opening / closing -> distance band -> 50x50 box blur
// (replaces Coupled_motion_planner.py:37-80 surface_normal, :83-95 image_filling,
//  :97-109 structural_disk and the inline pipeline :1144-1216 of main()).
//
// The reference strings together cv2.erode / cv2.dilate / cv2.floodFill,
// scipy.ndimage.distance_transform_edt and a 2500-tap scipy.signal.convolve2d.  Here:
//   * every erosion / dilation by the reference's digital disk {di^2 + dj^2 <= r^2} is a threshold
//     on an exact squared Euclidean distance (dilate: a set pixel within r; erode: no clear pixel
//     within r), computed by a vertical scan (64-row segment summaries) + a bounded horizontal scan -- bit-identical
//     to cv2 including its border rule (outside counts as set for erode, clear for dilate);
//   * image_filling (flood fill from pixel (0,0), then OR of the unreached zeros) is a
//     union-find labelling of the zero pixels (row runs pre-linked by ballot, vertical links by
//     atomicMin), holes = zero pixels whose root differs from pixel (0,0)'s;
//   * the exact EDT is the same two-scan scheme without a bound (the horizontal scan stops as
//     soon as dx^2 reaches the best value), integer d^2 then one correctly rounded sqrt;
//   * the blur is two 50-tap passes (row sums, then column sums) instead of 2500 taps per cell.
// Everything up to the blur reproduces the reference's bits (individually rounded fp64 in its
// operation order); the blur differs only by summation order (<= 1e-14 relative, tests: 1e-12).
// All kernels are grid-stride and collective-free except cm_runs_kernel (warp ballots), so the
// host emulator (tools/host_emu) runs them unchanged.
#pragma once
#include "fm_common.cuh"

namespace fmb {

struct CostmapCtl {               // first 256 bytes of the workspace
    int d2max;                    // max squared distance to an obstacle
    int n_positive;               // cells with a positive distance-band value (0 => the reference raises)
    unsigned long long min_pos;   // bits of the smallest positive distance-band value
    int pad[60];
};

constexpr int CM_FAR = 30000;     // "no feature in this column" (FAR^2 fits an int, exceeds any real d^2 for n <= 8192)

#define CM_GRID_STRIDE(i, total) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)

// ---- DEM -> raw obstacle map (:37-80, :1144-1160) ---------------------------------------------
// The reference pads grid and DEM by one extrapolated ring before its 'valid' convolutions; the ring
// only ever reaches the outermost pixels, which are cleared right after (:1156-1160), so interior
// pixels read plain neighbours.
// 0.5*a + (-0.5)*b: the only two non-zero taps of the reference's 3x3 stencils
__device__ __forceinline__ double cm_half_diff(double a, double b) { return __dadd_rn(__dmul_rn(0.5, a), __dmul_rn(-0.5, b)); }

__global__ void cm_slope_kernel(const double *dem, const double *grid, int n, double slope_max, unsigned char *out) {
    CM_GRID_STRIDE(i, (long long)n * n) {
        const int y = (int)(i / n), x = (int)(i - (long long)y * n);
        unsigned char o = 0;
        if (y > 0 && x > 0 && y < n - 1 && x < n - 1) {          // map limits are cleared (:1156-1160)
            const double ax = -cm_half_diff(grid[x + 1], grid[x - 1]);
            const double ay = -cm_half_diff(grid[y], grid[y]);
            const double az = -cm_half_diff(dem[i + 1], dem[i - 1]);
            const double bx = cm_half_diff(grid[x], grid[x]);
            const double by = cm_half_diff(grid[y + 1], grid[y - 1]);
            const double bz = cm_half_diff(dem[i + n], dem[i - n]);
            const double nx = -__dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by));
            const double ny = -__dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz));
            const double nz = -__dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
            double mag = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
            if (mag == 0.0) mag = 2.220446049250313e-16;
            o = acos(__ddiv_rn(nz, mag)) > slope_max ? 1 : 0;
        }
        out[i] = o;
    }
}

// ---- hole filling (:83-95) -----------------------------------------------------------------------
__device__ __forceinline__ int cm_find(const int *L, int x) {
    int y;
    while ((y = ld_volatile(&L[x])) != x) x = y;
    return x;
}
__device__ __forceinline__ void cm_union(int *L, int a, int b) {
    for (;;) {
        a = cm_find(L, a);
        b = cm_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // a > b: hang the larger root under the smaller
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}
// label of a zero pixel = first pixel of its run inside its 32-pixel chunk; set pixels get -1
__global__ void cm_runs_kernel(const unsigned char *im, int n, int *L) {
    const int lane = threadIdx.x & 31;
    const long long chunks_per_row = (n + 31) / 32;
    const long long nchunks = chunks_per_row * n;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp0; c < nchunks; c += nwarps) {
        const int y = (int)(c / chunks_per_row), x = (int)(c - (long long)y * chunks_per_row) * 32 + lane;
        const bool zero = x < n && im[(long long)y * n + x] == 0;
        const unsigned m = __ballot_sync(FULL, zero);
        if (x < n) {
            const unsigned below_set = ~m & ((1u << lane) - 1u);       // non-zero pixels before me in the chunk
            const int start = below_set ? 32 - __clz(below_set) : 0;
            L[(long long)y * n + x] = zero ? (int)((long long)y * n + x - lane + start) : -1;
        }
    }
}
// link runs across chunk boundaries and between rows (4-connectivity)
__global__ void cm_link_kernel(const unsigned char *im, int n, int *L) {
    CM_GRID_STRIDE(i, (long long)n * n) {
        if (im[i] != 0) continue;
        const int y = (int)(i / n), x = (int)(i - (long long)y * n);
        const bool left = x > 0 && im[i - 1] == 0;
        if (left && (x & 31) == 0) cm_union(L, (int)i, (int)i - 1);
        if (y > 0 && im[i - n] == 0) {
            // when the left pixel and the pixel above it are zero too, this pixel already reaches the
            // one above through them (runs are linked horizontally, the left pixel links upwards itself)
            const bool via_left = left && im[i - n - 1] == 0;
            if (!via_left) cm_union(L, (int)i, (int)i - n);
        }
    }
}
__global__ void cm_fill_apply_kernel(const unsigned char *im, int n, const int *L, unsigned char *out) {
    const int root0 = im[0] == 0 ? cm_find(L, 0) : -1;
    CM_GRID_STRIDE(i, (long long)n * n) {
        unsigned char v = im[i];
        if (v == 0 && cm_find(L, (int)i) != root0) v = 1;
        out[i] = v;
    }
}

// ---- erosion / dilation by a disk, exact EDT -----------------------------------------------------
// out = 1 where a pixel equal to `fv` lies within Euclidean distance r (dilate: fv = 1, on_hit = 1),
// out = 0 where one does (erode: fv = 0, on_hit = 0)
__global__ void cm_hscan_threshold_kernel(const int *g, int n, int r, int on_hit, unsigned char *out) {
    const int r2 = r * r;
    CM_GRID_STRIDE(i, (long long)n * n) {
        const int x = (int)(i % n);
        bool hit = false;
        const int lo = max(-r, -x), hi = min(r, n - 1 - x);
        for (int dx = lo; dx <= hi && !hit; ++dx) {
            const int gv = g[i + dx];
            hit = gv <= r && dx * dx + gv * gv <= r2;
        }
        out[i] = (unsigned char)(hit ? on_hit : 1 - on_hit);
    }
}
// unbounded vertical scan in two kernels, one thread per (64-row segment, column):
//   cm_vseg_kernel   first / last feature row of every segment (-1 = none);
//   cm_vscan_full_kernel  nearest feature above / below the segment from those summaries (<= n/64
//                    reads each way instead of a walk over rows), then one sweep down and one up.
constexpr int CM_SEG = 64;
__global__ void cm_vseg_kernel(const unsigned char *im, int fv, int n, int *seg_first, int *seg_last) {
    const int nseg = (n + CM_SEG - 1) / CM_SEG;
    CM_GRID_STRIDE(t, (long long)nseg * n) {
        const int sgm = (int)(t / n), x = (int)(t - (long long)sgm * n);
        const int y0 = sgm * CM_SEG, y1 = min(y0 + CM_SEG, n);
        int first = -1, last = -1;
        for (int y = y0; y < y1; ++y)
            if (im[(long long)y * n + x] == fv) { if (first < 0) first = y; last = y; }
        seg_first[t] = first;
        seg_last[t] = last;
    }
}
__global__ void cm_vscan_full_kernel(const unsigned char *im, int fv, int n, const int *seg_first, const int *seg_last, int *g) {
    const int nseg = (n + CM_SEG - 1) / CM_SEG;
    CM_GRID_STRIDE(t, (long long)nseg * n) {
        const int sgm = (int)(t / n), x = (int)(t - (long long)sgm * n);
        const int y0 = sgm * CM_SEG, y1 = min(y0 + CM_SEG, n);
        int d = CM_FAR;                                   // distance from row y0 - 1 to the nearest feature above
        for (int q = sgm - 1; q >= 0; --q) {
            const int l = seg_last[(long long)q * n + x];
            if (l >= 0) { d = y0 - 1 - l; break; }
        }
        for (int y = y0; y < y1; ++y) {
            d = im[(long long)y * n + x] == fv ? 0 : min(d + 1, CM_FAR);
            g[(long long)y * n + x] = d;
        }
        d = CM_FAR;                                       // distance from row y1 to the nearest feature below
        for (int q = sgm + 1; q < nseg; ++q) {
            const int f = seg_first[(long long)q * n + x];
            if (f >= 0) { d = f - y1; break; }
        }
        for (int y = y1 - 1; y >= y0; --y) {
            d = im[(long long)y * n + x] == fv ? 0 : min(d + 1, CM_FAR);
            const long long i = (long long)y * n + x;
            if (d < g[i]) g[i] = d;
        }
    }
}
// exact squared distance to the nearest feature pixel; the scan stops once dx^2 cannot improve
__global__ void cm_hscan_exact_kernel(const int *g, int n, int *d2, CostmapCtl *ctl) {
    int local_max = 0;
    CM_GRID_STRIDE(i, (long long)n * n) {
        const int x = (int)(i % n);
        const int g0 = g[i];
        int best = g0 * g0;
        // four offsets per round, all eight loads issued before `best` is touched; offsets past the
        // stopping bound are still genuine candidates, so looking at them cannot change the minimum
        for (int dx = 1; dx * dx < best && (x - dx >= 0 || x + dx < n); dx += 4) {
            int cand[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = dx + u;
                const int gl = x - e >= 0 ? g[i - e] : CM_FAR, gr = x + e < n ? g[i + e] : CM_FAR;
                cand[2 * u] = e * e + gl * gl;
                cand[2 * u + 1] = e * e + gr * gr;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) best = cand[u] < best ? cand[u] : best;
        }
        d2[i] = best;
        local_max = best > local_max ? best : local_max;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(FULL, local_max, o); local_max = v > local_max ? v : local_max; }
    if ((threadIdx.x & 31) == 0 && local_max > 0) atomicMax(&ctl->d2max, local_max);
}

// ---- distance band, composition, blur (:1186-1216) -----------------------------------------------
__global__ void cm_init_ctl_kernel(CostmapCtl *ctl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->d2max = 0; ctl->n_positive = 0; ctl->min_pos = ~0ULL; }
}
__global__ void cm_set_border_kernel(unsigned char *im, int n, unsigned char v) {
    CM_GRID_STRIDE(k, (long long)n) {
        im[k] = v; im[(long long)(n - 1) * n + k] = v; im[k * n] = v; im[k * n + n - 1] = v;
    }
}
__device__ __forceinline__ double cm_band(int dil, int d2, double res, double maxd) {
    // dilatedObstMap*(1 - obstDistance/max(obstDistance)), obstDistance = resolution*edt
    const double dist = __dmul_rn(res, __dsqrt_rn((double)d2));
    return __dmul_rn((double)dil, __dsub_rn(1.0, __ddiv_rn(dist, maxd)));
}
__global__ void cm_band_min_kernel(const unsigned char *dil, const int *d2, int n, double res, CostmapCtl *ctl) {
    const double maxd = __dmul_rn(res, __dsqrt_rn((double)ld_volatile(&ctl->d2max)));
    unsigned long long local = ~0ULL;
    int cnt = 0;
    CM_GRID_STRIDE(i, (long long)n * n) {
        const double v = cm_band(dil[i], d2[i], res, maxd);
        if (v > 0.0) { const unsigned long long b = (unsigned long long)__double_as_longlong(v); local = b < local ? b : local; ++cnt; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(FULL, local, o);
        local = v < local ? v : local;
        cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicMin(&ctl->min_pos, local); atomicAdd(&ctl->n_positive, cnt); }
}
// pre-blur cost in [y][x] order: 1 + (obst*300 + band*10)   (the reference holds its transpose)
__global__ void cm_compose_kernel(const unsigned char *obst, const unsigned char *dil, const int *d2, int n, double res,
                                  const CostmapCtl *ctl, double *pre) {
    const double maxd = __dmul_rn(res, __dsqrt_rn((double)ctl->d2max));
    const double mind = __longlong_as_double((long long)ctl->min_pos);
    CM_GRID_STRIDE(i, (long long)n * n) {
        double v = cm_band(dil[i], d2[i], res, maxd);
        if (v > 0.0) v = __dsub_rn(v, mind);
        pre[i] = __dadd_rn(1.0, __dadd_rn(__dmul_rn((double)obst[i], 300.0), __dmul_rn(v, 10.0)));
    }
}
// 50-tap row sums over x-25 .. x+24, 300 outside the map (convolve2d mode='same', fillvalue=300)
__global__ void cm_blur_rows_kernel(const double *pre, int n, double *tmp) {
    CM_GRID_STRIDE(i, (long long)n * n) {
        const int x = (int)(i % n);
        double s = 0.0;
#pragma unroll 10
        for (int u = -25; u < 25; ++u) {
            const int xx = x + u;
            s = __dadd_rn(s, (xx >= 0 && xx < n) ? pre[i + u] : 300.0);
        }
        tmp[i] = s;
    }
}
// 50-tap column sums over y-25 .. y+24, scaled by 1/2500; map limits become +inf (:1210-1214).
// One thread produces CM_VB vertically adjacent outputs from one pass over the 49 + CM_VB rows they
// share (each output still adds its taps in ascending order).
constexpr int CM_VB = 8;
__global__ void cm_blur_cols_kernel(const double *tmp, int n, double *out) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int nyb = (n + CM_VB - 1) / CM_VB;
    CM_GRID_STRIDE(t, (long long)nyb * n) {
        const int yb = (int)(t / n), x = (int)(t - (long long)yb * n);
        const int y0 = yb * CM_VB;
        double s[CM_VB];
#pragma unroll
        for (int k = 0; k < CM_VB; ++k) s[k] = 0.0;
#pragma unroll 2
        for (int j = 0; j < 49 + CM_VB; ++j) {             // row y0 - 25 + j feeds output k as its tap j - k
            const int yy = y0 - 25 + j;
            const double v = (yy >= 0 && yy < n) ? tmp[(long long)yy * n + x] : 15000.0;
#pragma unroll
            for (int k = 0; k < CM_VB; ++k)
                if (j - k >= 0 && j - k < 50) s[k] = __dadd_rn(s[k], v);
        }
#pragma unroll
        for (int k = 0; k < CM_VB; ++k) {
            const int y = y0 + k;
            if (y < n) {
                const bool edge = y == 0 || x == 0 || y == n - 1 || x == n - 1;
                out[(long long)y * n + x] = edge ? inf : __dmul_rn(s[k], 1.0 / 2500.0);
            }
        }
    }
}

}  // namespace fmb
