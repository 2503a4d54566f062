// costvolume.cuh -- the planner's 3D arm-workspace cost volume on the device:
// GetObstMap (Coupled_motion_planner.py:319-358), TunnelCost (:505-725) and Cmap1*Cmap2 (:1627).
//
// The reference fills the tunnel by a sequential scatter whose result depends on write order:
// graded costs are written only while a voxel still holds its initial 10 ("first writer wins"),
// +inf writes are unconditional.  Per voxel that is: +inf if any inf event hits it, else the value
// of the FIRST value event in loop order (events whose value is exactly 10 leave the voxel
// untouched), else 10.  Here every event of the three loop nests (tube around the base path,
// closing wall behind the first pose, half sphere around the last pose) gets its position in that
// order as a sequence number; one kernel scatters `atomicMin(sequence)` / inf flags, a second one
// decodes the winning sequence number back into its table value and multiplies by the terrain
// volume, which needs no scatter at all (one voxel per DEM column is +inf, map limits are +inf).
// Frames, linspace axes and per-(i,k) value tables are prepared by the host wrapper with the
// reference's own scalar expressions (a few thousand numbers); voxel coordinates use the
// product + three fused multiply-adds that numpy's dot (OpenBLAS dgemm) performs.
#pragma once
#include "fm_common.cuh"

namespace fmb {

struct CostVolumeArgs {
    // terrain (GetObstMap)
    const double *Zs; int zm, zn;                 // DEM crop, rows x cols
    double resX, resY, resZ, xm, ym;
    int sX, sY, sZ;
    // tunnel (TunnelCost)
    const double *frames;                         // (npose + 1) x 12: base frames of :530-533 per pose, then the half-sphere frame (:661-664)
    int npose;
    const double *li, *lk; int nX, nZ;            // linspace axes of the tube cross-section (:546-547)
    const double *norm, *val;                     // nX x nZ: sqrt(i^2 + k^2) and the graded cost of :573
    double rlim;
    const double *lr, *hval; int nK;              // radii of the half sphere and their cost (:676, :697)
    const double *ct, *st, *cs, *ss;              // cos / sin of the 100 theta and 90 sigma angles (:668-681)
    double shell;                                 // rlim + 2 resZ (:704)
    long long fin[3], ini[3];                     // sample node and initial end-effector node (never set to inf)
    int *first; unsigned char *blocked;             // workspace: winning sequence number / inf flag per voxel
    double *cmap, *tunnel, *terrain;              // outputs (tunnel / terrain may be null)
};

constexpr int CV_NONE = 0x7fffffff;
#define CV_GRID_STRIDE(i, total) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ double cv_dot_last(const double *row, double x, double y, double z) {
    double s = __dmul_rn(row[0], x);
    s = __fma_rn(row[1], y, s);
    s = __fma_rn(row[2], z, s);
    return __fma_rn(row[3], 1.0, s);
}
// voxel of the point (x, y, z) given in frame F; false when it falls outside the volume
__device__ __forceinline__ bool cv_voxel(const CostVolumeArgs &A, const double *F, double x, double y, double z,
                                         long long &ix, long long &iy, long long &iz) {
    const double fx = rint(__ddiv_rn(cv_dot_last(F, x, y, z), A.resX));
    const double fy = rint(__ddiv_rn(cv_dot_last(F + 4, x, y, z), A.resY));
    const double fz = rint(__ddiv_rn(cv_dot_last(F + 8, x, y, z), A.resZ));
    if (!(fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx < (double)A.sX && fy < (double)A.sY && fz < (double)A.sZ)) return false;
    ix = (long long)fx; iy = (long long)fy; iz = (long long)fz;
    return true;
}
__device__ __forceinline__ bool cv_may_block(const CostVolumeArgs &A, long long ix, long long iy, long long iz) {
    return (ix != A.fin[0] || iy != A.fin[1] || iz != A.fin[2]) && (ix != A.ini[0] || iy != A.ini[1] || iz != A.ini[2]);
}
__device__ __forceinline__ long long cv_cell(const CostVolumeArgs &A, long long ix, long long iy, long long iz) {
    return (iy * A.sX + ix) * A.sZ + iz;          // Cmap[iy, ix, iz] of shape (sY, sX, sZ)
}

__global__ void cv_init_kernel(CostVolumeArgs A) {
    CV_GRID_STRIDE(c, (long long)A.sX * A.sY * A.sZ) { A.first[c] = CV_NONE; A.blocked[c] = 0; }
}

// all events of TunnelCost, numbered in the reference's loop order
__global__ void cv_scatter_kernel(CostVolumeArgs A) {
    const long long plane = (long long)A.nX * A.nZ;
    const long long n1 = (long long)A.npose * plane;              // tube: two writes per (pose, i, k)
    const long long n2 = plane;                                    // closing wall
    const long long n3 = 100LL * 90LL * (A.nK + 1);                // half sphere: nK radii + the shell point per direction
    CV_GRID_STRIDE(e, n1 + n2 + n3) {
        long long ix, iy, iz;
        if (e < n1) {
            const long long j = e / plane, r = e - j * plane;
            const int a = (int)(r / A.nZ), b = (int)(r - (long long)a * A.nZ);
            const double *F = A.frames + 12 * j;
            const double i = A.li[a], k = A.lk[b];
            const bool inside_reach = A.norm[r] < A.rlim;
            const bool counts = A.val[r] != 10.0;                  // writing 10 over 10 changes nothing
            if (cv_voxel(A, F, i, 0.0, k, ix, iy, iz)) {           // :556-579
                const long long c = cv_cell(A, ix, iy, iz);
                if (inside_reach) { if (counts) atomicMin(&A.first[c], (int)(2 * e)); }
                else if (cv_may_block(A, ix, iy, iz)) A.blocked[c] = 1;
            }
            if (inside_reach && counts && cv_voxel(A, F, i, A.resY, k, ix, iy, iz))      // :583-598 one step ahead
                atomicMin(&A.first[cv_cell(A, ix, iy, iz)], (int)(2 * e + 1));
        } else if (e < n1 + n2) {
            const long long r = e - n1;
            const int a = (int)(r / A.nZ), b = (int)(r - (long long)a * A.nZ);
            if (A.norm[r] < A.rlim && cv_voxel(A, A.frames, A.li[a], -A.resY, A.lk[b], ix, iy, iz) &&
                cv_may_block(A, ix, iy, iz))                        // :620-642
                A.blocked[cv_cell(A, ix, iy, iz)] = 1;
        } else {
            const long long r = e - n1 - n2;
            const long long dir = r / (A.nK + 1);
            const int c = (int)(r - dir * (A.nK + 1));
            const int ti = (int)(dir / 90), si = (int)(dir - (long long)ti * 90);
            const double *F = A.frames + 12 * (long long)A.npose;
            const double ct = A.ct[ti], st = A.st[ti], cs = A.cs[si], ss = A.ss[si];
            const double rad = c < A.nK ? A.lr[c] : A.shell;
            const double px = __dmul_rn(__dmul_rn(rad, ct), cs), py = __dmul_rn(__dmul_rn(rad, ct), ss), pz = __dmul_rn(rad, st);
            if (!cv_voxel(A, F, px, py, pz, ix, iy, iz)) continue;
            const long long cell = cv_cell(A, ix, iy, iz);
            if (c < A.nK) {                                        // :684-697
                if (A.hval[c] != 10.0) atomicMin(&A.first[cell], (int)(2 * n1 + dir * A.nK + c));
            } else if (cv_may_block(A, ix, iy, iz)) A.blocked[cell] = 1;   // :701-721
        }
    }
}

// tunnel volume from the winners, terrain volume per voxel, and their product
__global__ void cv_compose_kernel(CostVolumeArgs A) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const long long plane = (long long)A.nX * A.nZ;
    const long long n1x2 = 2 * (long long)A.npose * plane;
    CV_GRID_STRIDE(c, (long long)A.sX * A.sY * A.sZ) {
        double t = 10.0;
        if (A.blocked[c]) t = inf;
        else {
            const int s = A.first[c];
            if (s != CV_NONE) t = s < n1x2 ? A.val[(s >> 1) % plane] : A.hval[(s - n1x2) % A.nK];
        }
        // GetObstMap: arrays of shape (sX, sY, sZ) indexed [j, i, iz]
        const int iz = (int)(c % A.sZ);
        const long long ji = c / A.sZ;
        const int i = (int)(ji % A.sY), j = (int)(ji / A.sY);
        double g = 2.0;                                             // obstMap + groundMap = 1 + 1
        if (j == 0 || j == A.sX - 1 || i == 0 || i == A.sY - 1 || iz == 0 || iz == A.sZ - 1) g = inf;
        else if (j < A.zm && i < A.zn && i < A.sX && j < A.sY && __dmul_rn(A.resX, (double)i) != A.xm &&
                 __dmul_rn(A.resY, (double)j) != A.ym && rint(__ddiv_rn(A.Zs[(long long)j * A.zn + i], A.resZ)) == (double)iz)
            g = inf;                                                // the terrain voxel of this column (obstacle or ground)
        if (A.tunnel) A.tunnel[c] = t;
        if (A.terrain) A.terrain[c] = g;
        if (A.cmap) A.cmap[c] = __dmul_rn(g, t);
    }
}

}  // namespace fmb
