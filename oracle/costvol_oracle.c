/*
 * costvol_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the planner's 3D cost-volume construction:
 *   GetObstMap   src/Coupled_motion_planner.py:319-358
 *   TunnelCost   src/Coupled_motion_planner.py:505-725
 *   Cmap = Cmap1 * Cmap2   :1627
 * written as the same sequential loops (order matters: "first writer wins" guards `Cmap == 10`,
 * unconditional inf writes), in the reference's floating-point operation order.  Third-party
 * arithmetic restated here:
 *   np.linspace(a, b, N)      v[t] = t*step + a, step = (b - a)/(N - 1), v[N-1] = b   (numpy 2.3)
 *   np.dot(4x4, 4x4)[r][3]    OpenBLAS 0.3.30 dgemm as measured in this container: one product,
 *                             then three fused multiply-adds in column order
 *   x**2 on numpy scalars     libm pow(x, 2.0)
 *   round()                   half to even (rint)
 * Pinned by tests/golden/costvolume.npz (inputs / outputs captured from the unmodified planner
 * run and from direct calls of the unmodified functions, oracle/gen_golden.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
static volatile double cv_two = 2.0;
static inline double pw2(double x) { return pow(x, cv_two); }

static void linspace(double a, double b, int n, double *v) {
    if (n == 1) { v[0] = a; return; }
    const double step = (b - a) / (double)(n - 1);
    for (int t = 0; t < n; ++t) v[t] = (double)t * step + a;
    v[n - 1] = b;
}

/* row r of Toa (4 entries) times the last column of the translation (x, y, z, 1) */
static inline double dot_last(const double *row, double x, double y, double z) {
    double s = row[0] * x;
    s = fma(row[1], y, s);
    s = fma(row[2], z, s);
    s = fma(row[3], 1.0, s);
    return s;
}

static void base_frame(double alpha, double beta, double gamma, double px, double py, double pz, double *T /* 3x4 */) {
    const double ca = cos(alpha), cb = cos(beta), cg = cos(gamma), sa = sin(alpha), sb = sin(beta), sg = sin(gamma);
    T[0] = ca * cb; T[1] = ca * sb * sg - sa * cg; T[2] = ca * sb * cg + sa * sg; T[3] = px;
    T[4] = sa * cb; T[5] = sa * sb * sg + ca * cg; T[6] = sa * sb * cg - ca * sg; T[7] = py;
    T[8] = -sb;     T[9] = cb * sg;                T[10] = cb * cg;               T[11] = pz;
}

/* GetObstMap: finalMap of shape (sX, sY, sZ) indexed [j][i][iz]; returns -1 on an index the
 * reference would fault on */
int cv_obst_map(const double *Zs, int m, int n, double resX, double resY, double resZ, int sX, int sY, int sZ,
                const double *newObst, double xm, double ym, double *finalMap) {
    const long long total = (long long)sX * sY * sZ;
    double *obst = malloc(sizeof(double) * total), *ground = malloc(sizeof(double) * total);
    for (long long t = 0; t < total; ++t) { obst[t] = 1.0; ground[t] = 1.0; }
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i)
        for (int j = 0; j < m; ++j) {
            if (resX * (double)i != xm && resY * (double)j != ym) {
                const double q = rint(Zs[(long long)j * n + i] / resZ);
                if (i < sX && j < sY && q < (double)sZ) {
                    const long long iz = (long long)q;
                    if (j >= sX || i >= sY || iz < 0) { rc = -1; break; }
                    const long long o = ((long long)j * sY + i) * sZ + iz;
                    if (newObst[(long long)j * n + i] == 1.0) obst[o] = INFINITY; else ground[o] = INFINITY;
                }
            }
        }
    for (long long t = 0; t < total; ++t) finalMap[t] = obst[t] + ground[t];
    for (int a = 0; a < sX; ++a)
        for (int b = 0; b < sY; ++b)
            for (int c = 0; c < sZ; ++c)
                if (a == 0 || a == sX - 1 || b == 0 || b == sY - 1 || c == 0 || c == sZ - 1)
                    finalMap[((long long)a * sY + b) * sZ + c] = INFINITY;
    free(obst); free(ground);
    return rc;
}

static inline int special(long long ix, long long iy, long long iz, const int64_t *fin, const int64_t *ini) {
    /* True when the cell is neither the sample node nor the initial end-effector node */
    return (ix != fin[0] || iy != fin[1] || iz != fin[2]) && (ix != ini[0] || iy != ini[1] || iz != ini[2]);
}

/* TunnelCost: Cmap of shape (sY, sX, sZ) indexed [iy][ix][iz]; gamma2D (m x 3), heading (m x 3) */
int cv_tunnel_cost(double rlim, double rO, double rm, const double *gamma2D, int m, int sX, int sY, int sZ,
                   double resX, double resY, double resZ, const double *heading, const int64_t *fin,
                   const int64_t *ini, double *Cmap) {
    const long long total = (long long)sY * sX * sZ;
    for (long long t = 0; t < total; ++t) Cmap[t] = 10.0;
    const double gradient = 15.0;
    const double tunnelRad = rlim + 2 * resX;
    const int nX = (int)(rint(2 * tunnelRad / resX) + 1), nZ = (int)(rint(2 * tunnelRad / resZ) + 1);
    if (nX < 1 || nZ < 1) return -1;
    double *li = malloc(sizeof(double) * nX), *lk = malloc(sizeof(double) * nZ);
    linspace(-tunnelRad, tunnelRad, nX, li);
    linspace(-tunnelRad, tunnelRad, nZ, lk);
    const double mid = (rO + rm) / 2;
    double T[12];
#define CELL(ix, iy, iz) Cmap[((long long)(iy) * sX + (ix)) * sZ + (iz)]
#define INSIDE(ix, iy, iz) ((ix) >= 0 && (iy) >= 0 && (iz) >= 0 && (ix) < sX && (iy) < sY && (iz) < sZ)
    for (int j = 0; j < m; ++j) {
        base_frame(heading[3 * j + 2] - M_PI / 2, heading[3 * j + 1], heading[3 * j], gamma2D[3 * j], gamma2D[3 * j + 1],
                   gamma2D[3 * j + 2], T);
        for (int a = 0; a < nX; ++a)
            for (int b = 0; b < nZ; ++b) {
                const double i = li[a], k = lk[b];
                long long ix = (long long)rint(dot_last(T, i, 0.0, k) / resX);
                long long iy = (long long)rint(dot_last(T + 4, i, 0.0, k) / resY);
                long long iz = (long long)rint(dot_last(T + 8, i, 0.0, k) / resZ);
                const double norm = sqrt(pw2(i) + pw2(k));
                if (INSIDE(ix, iy, iz)) {
                    if (norm < rlim) {
                        if (CELL(ix, iy, iz) == 10.0) CELL(ix, iy, iz) = gradient * pw2(norm - mid) + 2 + 4 * (i + rlim + 2 * resZ);
                    } else if (special(ix, iy, iz, fin, ini)) CELL(ix, iy, iz) = INFINITY;
                }
                ix = (long long)rint(dot_last(T, i, resY, k) / resX);
                iy = (long long)rint(dot_last(T + 4, i, resY, k) / resY);
                iz = (long long)rint(dot_last(T + 8, i, resY, k) / resZ);
                if (INSIDE(ix, iy, iz) && CELL(ix, iy, iz) == 10.0 && norm < rlim)
                    CELL(ix, iy, iz) = gradient * pw2(norm - mid) + 2 + 4 * (i + rlim + 2 * resZ);
            }
    }
    /* closing wall one step behind the first base position (:601-640) */
    base_frame(heading[2] - M_PI / 2, heading[1], heading[0], gamma2D[0], gamma2D[1], gamma2D[2], T);
    for (int a = 0; a < nX; ++a)
        for (int b = 0; b < nZ; ++b) {
            const double i = li[a], k = lk[b];
            const long long ix = (long long)rint(dot_last(T, i, -resY, k) / resX);
            const long long iy = (long long)rint(dot_last(T + 4, i, -resY, k) / resY);
            const long long iz = (long long)rint(dot_last(T + 8, i, -resY, k) / resZ);
            const double norm = sqrt(pw2(i) + pw2(k));
            if (INSIDE(ix, iy, iz) && norm < rlim && special(ix, iy, iz, fin, ini)) CELL(ix, iy, iz) = INFINITY;
        }
    /* half sphere around the last base position (:642-723); note: no -pi/2 on alpha here (:649) */
    base_frame(heading[3 * (m - 1) + 2], heading[3 * (m - 1) + 1], heading[3 * (m - 1)], gamma2D[3 * (m - 1)],
               gamma2D[3 * (m - 1) + 1], gamma2D[3 * (m - 1) + 2], T);
    const int nK = (int)rint((double)nZ / 2) + 1;
    double *lr = malloc(sizeof(double) * nK);
    linspace(0.0, tunnelRad, nK, lr);
    const double shell = rlim + 2 * resZ;
    for (int a = -100; a < 100; a += 2) {
        const double theta = M_PI * a / 180, ct = cos(theta), st = sin(theta);
        for (int b = -90; b < 90; b += 2) {
            const double sigma = M_PI * b / 180, cs = cos(sigma), ss = sin(sigma);
            for (int c = 0; c < nK; ++c) {
                const double k = lr[c];
                const long long ix = (long long)rint(dot_last(T, k * ct * cs, k * ct * ss, k * st) / resX);
                const long long iy = (long long)rint(dot_last(T + 4, k * ct * cs, k * ct * ss, k * st) / resY);
                const long long iz = (long long)rint(dot_last(T + 8, k * ct * cs, k * ct * ss, k * st) / resZ);
                if (INSIDE(ix, iy, iz) && CELL(ix, iy, iz) == 10.0) CELL(ix, iy, iz) = gradient * pw2(k - mid) + 2;
            }
            const long long ix = (long long)rint(dot_last(T, shell * ct * cs, shell * ct * ss, shell * st) / resX);
            const long long iy = (long long)rint(dot_last(T + 4, shell * ct * cs, shell * ct * ss, shell * st) / resY);
            const long long iz = (long long)rint(dot_last(T + 8, shell * ct * cs, shell * ct * ss, shell * st) / resZ);
            if (INSIDE(ix, iy, iz) && special(ix, iy, iz, fin, ini)) CELL(ix, iy, iz) = INFINITY;
        }
    }
    free(li); free(lk); free(lr);
    return 0;
}
