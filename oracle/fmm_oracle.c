/*
 * oracle/fmm_oracle.c -- CPU restatement of the reference's Fast Marching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (FastMarching/,
 * planning_motion_planning_b200/) may import, link or call this file.  It is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, and nothing else.
 *
 * It follows the reference (esa-prl/planning-motion_planning, src/FastMarching)
 * expression-for-expression so that fields agree with the Python reference
 * bitwise and paths agree to the last ulp:
 *
 *   orc_eikonal2d      <- FastMarching.py:17-29     getEikonal
 *   f2_update          <- FastMarching.py:44-80     updateNode (2D)
 *   f2_pop             <- FastMarching.py:82-89     getMinNB
 *   orc_fmm2d          <- FastMarching.py:92-112    computeTmap (2D; *intended*
 *                         semantics -- the shipped function raises ValueError at
 *                         :107; this mirrors the working 3D driver :126-145)
 *   orc_bifmm2d        <- FastMarching.py:114-162   biComputeTmap
 *   orc_gradient2d     <- FastMarching.py:242-300   computeGradient
 *   orc_interp2d       <- FastMarching.py:305-338   interpolatePoint (2D)
 *   orc_trace2d        <- FastMarching.py:164-236   getPathGDM (2D)
 *   f3_update          <- FastMarching3D.py:19-101  updateNode (3D)
 *   orc_fmm3d          <- FastMarching3D.py:126-145 computeTmap (3D)
 *   orc_interp3d       <- FastMarching3D.py:275-314 interpolatePoint (3D)
 *   orc_trace3d        <- FastMarching3D.py:198-271 getPathGDM (3D)
 *
 * The reference keeps the narrow band as two parallel Python lists kept sorted
 * with bisect_left + list.insert (FastMarching.py:65-67).  bisect_left puts a new
 * key BEFORE equal keys, so among equal T the most recently (re)inserted node
 * pops first.  A binary heap ordered by (T ascending, insertion sequence
 * descending) with lazy deletion pops in exactly that order.
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path
 * ("parity unpinned" by the reference itself).  This file is
 * pinned instead against the reference *run in the build container*
 * (oracle/gen_golden.py -> tests/golden/), numpy 2.3.5 semantics.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no FMA contraction so
 * every product and sum rounds exactly as CPython/numpy do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/*
 * numpy *scalar* `x**2` (np.float64.__pow__) calls libm pow(x, 2.0), which glibc does
 * not round correctly in ~0.07 % of cases, whereas np.power(x, 2) and ndarray**2 take
 * numpy's exact `square` fast path (measured in the build container, numpy 2.3.5).
 * To stay bitwise on the reference, every place where the reference squares a numpy
 * scalar with `**2` uses pw2() (libm pow, kept opaque to the optimiser); places that
 * use np.power / ndarray**2 use a plain product.
 */
static volatile double orc_two = 2.0;
static inline double pw2(double x) { return pow(x, orc_two); }

#define ORC_OK 0
#define ORC_ERR_NOJOIN 1      /* biComputeTmap: fronts never met -> NameError (FastMarching.py:161) */
#define ORC_ERR_ALLOC 2

/* tracer status codes (observable behaviour of the reference) */
#define ORC_TRACE_OK 0          /* normal return, `end` appended */
#define ORC_TRACE_EARLY 1       /* 2D NaN fallback: bare except -> path so far, no `end` */
#define ORC_TRACE_VALUEERROR 2  /* NaN waypoint -> int(nan) raises ValueError */
#define ORC_TRACE_INDEXERROR 3  /* stencil leaves the array */
#define ORC_TRACE_OVERFLOW 4    /* 3D: inf waypoint -> int(round(inf)) raises OverflowError */

/* ------------------------------------------------------------------ heap -- */
typedef struct { double t; int64_t seq; int64_t idx; } hent;
typedef struct { hent *a; size_t n, cap; } heap_t;

static int hless(const hent *p, const hent *q) {
    if (p->t != q->t) return p->t < q->t;
    return p->seq > q->seq;               /* LIFO among equal keys (bisect_left) */
}
static int heap_push(heap_t *h, hent e) {
    if (h->n == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 1024;
        hent *na = (hent *)realloc(h->a, nc * sizeof(hent));
        if (!na) return -1;
        h->a = na; h->cap = nc;
    }
    size_t i = h->n++;
    while (i > 0) {
        size_t p = (i - 1) >> 1;
        if (!hless(&e, &h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = e;
    return 0;
}
static hent heap_pop(heap_t *h) {
    hent top = h->a[0];
    hent e = h->a[--h->n];
    size_t i = 0, n = h->n;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        const hent *best = &e;
        if (l < n && hless(&h->a[l], best)) { m = l; best = &h->a[l]; }
        if (r < n && hless(&h->a[r], best)) { m = r; best = &h->a[r]; }
        if (m == i) break;
        h->a[i] = h->a[m]; i = m;
    }
    if (n) h->a[i] = e;
    return top;
}

/* ------------------------------------------------------------- 2D solver -- */
/* FastMarching.py:17-29 */
double orc_eikonal2d(double Thor, double Tver, double cost) {
    if (isinf(Thor)) {
        if (isinf(Tver)) return INFINITY;
        return Tver + cost;
    }
    if (isinf(Tver)) return Thor + cost;
    if (cost < fabs(Thor - Tver)) return fmin(Thor, Tver) + cost;
    {
        double d = Thor - Tver;
        double disc = 2 * (cost * cost) - d * d;
        return .5 * (Thor + Tver + sqrt(disc));
    }
}

typedef struct {
    int rows, cols;
    const double *cost;
    double *T;
    uint8_t *closed;
    int64_t *cur_seq;
    int64_t seq, live;      /* live = cells currently in the narrow band */
    int64_t evals;          /* local-solver evaluations (for the bench's evals/s) */
    heap_t h;
} front2d;

static inline double f2_T(const front2d *f, int x, int y) {
    /* out-of-domain reads behave as +inf (the caller guarantees an inf border,
     * Coupled_motion_planner.py:1213-1216; the reference would wrap or raise) */
    if (x < 0 || y < 0 || x >= f->cols || y >= f->rows) return INFINITY;
    return f->T[(int64_t)y * f->cols + x];
}

/* FastMarching.py:44-80 */
static int f2_update(front2d *f, int nx, int ny) {
    static const int off[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
    for (int i = 0; i < 4; ++i) {
        int x = nx + off[i][0], y = ny + off[i][1];
        if (x < 0 || y < 0 || x >= f->cols || y >= f->rows) continue;
        int64_t id = (int64_t)y * f->cols + x;
        if (f->closed[id]) continue;
        double Thor = fmin(f2_T(f, x + 1, y), f2_T(f, x - 1, y));
        double Tver = fmin(f2_T(f, x, y + 1), f2_T(f, x, y - 1));
        double T = orc_eikonal2d(Thor, Tver, f->cost[id]);
        f->evals++;
        if (isinf(f->T[id])) {
            hent e = {T, ++f->seq, id};
            if (heap_push(&f->h, e)) return -1;
            f->cur_seq[id] = e.seq; f->T[id] = T; f->live++;
        } else if (T < f->T[id]) {
            hent e = {T, ++f->seq, id};
            if (heap_push(&f->h, e)) return -1;
            f->cur_seq[id] = e.seq; f->T[id] = T;
        }
    }
    return 0;
}

/* FastMarching.py:82-89 */
static int f2_pop(front2d *f, int *x, int *y) {
    while (f->h.n) {
        hent e = heap_pop(&f->h);
        if (f->closed[e.idx] || f->cur_seq[e.idx] != e.seq) continue; /* stale */
        f->live--;
        *x = (int)(e.idx % f->cols); *y = (int)(e.idx / f->cols);
        return 1;
    }
    return 0;
}

static int f2_init(front2d *f, const double *cost, int rows, int cols, double *T, int sx, int sy) {
    int64_t n = (int64_t)rows * cols;
    memset(f, 0, sizeof(*f));
    f->rows = rows; f->cols = cols; f->cost = cost; f->T = T;
    f->closed = (uint8_t *)calloc((size_t)n, 1);
    f->cur_seq = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    if (!f->closed || !f->cur_seq) return -1;
    for (int64_t i = 0; i < n; ++i) { T[i] = INFINITY; f->closed[i] = isinf(cost[i]) ? 1 : 0; }
    T[(int64_t)sy * cols + sx] = 0.0;
    f->closed[(int64_t)sy * cols + sx] = 1;
    return f2_update(f, sx, sy);
}
static void f2_free(front2d *f) { free(f->closed); free(f->cur_seq); free(f->h.a); }

/*
 * Single-front FMM from goal; early exit when `start` is popped
 * (sx<0 => full field).  order_out (optional, length rows*cols) receives the
 * linear index of every popped node in pop order; *npop the count;
 * *evals the number of local-solver evaluations.
 */
int orc_fmm2d(const double *cost, int rows, int cols, int gx, int gy, int sx, int sy,
              double *T, int64_t *order_out, int64_t *npop, int64_t *evals) {
    front2d f;
    int64_t k = 0;
    if (f2_init(&f, cost, rows, cols, T, gx, gy)) { f2_free(&f); return ORC_ERR_ALLOC; }
    int x, y;
    while (f2_pop(&f, &x, &y)) {
        f.closed[(int64_t)y * cols + x] = 1;
        if (order_out) order_out[k] = (int64_t)y * cols + x;
        ++k;
        if (f2_update(&f, x, y)) { f2_free(&f); return ORC_ERR_ALLOC; }
        if (x == sx && y == sy) break;
    }
    if (npop) *npop = k;
    if (evals) *evals = f.evals;
    f2_free(&f);
    return ORC_OK;
}

/* FastMarching.py:114-162 */
int orc_bifmm2d(const double *cost, int rows, int cols, int gx, int gy, int sx, int sy,
                double *TG, double *TS, int32_t *join, int64_t *npop) {
    front2d G, S;
    int rc = ORC_ERR_NOJOIN;
    int64_t k = 0;
    if (f2_init(&G, cost, rows, cols, TG, gx, gy)) { f2_free(&G); return ORC_ERR_ALLOC; }
    if (f2_init(&S, cost, rows, cols, TS, sx, sy)) { f2_free(&G); f2_free(&S); return ORC_ERR_ALLOC; }
    int tgx = gx, tgy = gy, tsx = sx, tsy = sy;   /* nodeTargetG / nodeTargetS */
    while (G.live > 0 || S.live > 0) {
        if (G.live > 0) {
            f2_pop(&G, &tgx, &tgy);
            G.closed[(int64_t)tgy * cols + tgx] = 1;
            f2_update(&G, tgx, tgy);
        }
        if (S.live > 0) {
            f2_pop(&S, &tsx, &tsy);
            S.closed[(int64_t)tsy * cols + tsx] = 1;
            f2_update(&S, tsx, tsy);
        }
        ++k;
        if (S.closed[(int64_t)tgy * cols + tgx] == 1) { join[0] = tgx; join[1] = tgy; rc = ORC_OK; break; }
        if (G.closed[(int64_t)tsy * cols + tsx] == 1) { join[0] = tsx; join[1] = tsy; rc = ORC_OK; break; }
    }
    int64_t n = (int64_t)rows * cols;
    for (int64_t i = 0; i < n; ++i) {
        if (isnan(TG[i])) TG[i] = INFINITY;
        if (isnan(TS[i])) TS[i] = INFINITY;
    }
    if (npop) *npop = k;
    f2_free(&G); f2_free(&S);
    return rc;
}

/* ------------------------------------------------------------- 2D tracer -- */
/* Normalised inf-aware gradient at one node, FastMarching.py:262-297. */
static void grad_node2d(const double *c, int m, int n, int i, int j, double *gnx, double *gny) {
#define C2(jj, ii) c[(int64_t)(jj) * n + (ii)]
    double Gx, Gy;
    if (j == 0) Gy = C2(1, i) - C2(0, i);
    else if (j == m - 1) Gy = C2(j, i) - C2(j - 1, i);
    else if (isinf(C2(j + 1, i))) {
        if (isinf(C2(j - 1, i))) Gy = 0; else Gy = C2(j, i) - C2(j - 1, i);
    } else {
        if (isinf(C2(j - 1, i))) Gy = C2(j + 1, i) - C2(j, i);
        else Gy = (C2(j + 1, i) - C2(j - 1, i)) / 2;
    }
    if (i == 0) Gx = C2(j, 1) - C2(j, 0);
    else if (i == n - 1) Gx = C2(j, i) - C2(j, i - 1);
    else if (isinf(C2(j, i + 1))) {
        if (isinf(C2(j, i - 1))) Gx = 0; else Gx = C2(j, i) - C2(j, i - 1);
    } else {
        if (isinf(C2(j, i - 1))) Gx = C2(j, i + 1) - C2(j, i);
        else Gx = (C2(j, i + 1) - C2(j, i - 1)) / 2;
    }
    double nrm = sqrt(pw2(Gx) + pw2(Gy));              /* Gx[j,i]**2: scalar pow */
    *gnx = Gx / nrm;
    *gny = Gy / nrm;
#undef C2
}

/* FastMarching.py:242-300.  has_point==0 => whole map.  Outputs are full-size,
 * zero outside the window exactly as the reference. */
void orc_gradient2d(const double *cost, int m, int n, int has_point, double px, double py,
                    double *Gnx, double *Gny) {
    int jmin = 0, imin = 0, jmax = m, imax = n;
    if (has_point) {
        int a = (int)py + 3, b = (int)px + 3;
        jmax = m < a ? m : a; imax = n < b ? n : b;
        a = (int)(py - 3); b = (int)(px - 3);
        jmin = a > 0 ? a : 0; imin = b > 0 ? b : 0;
    }
    memset(Gnx, 0, sizeof(double) * (size_t)m * n);
    memset(Gny, 0, sizeof(double) * (size_t)m * n);
    for (int i = imin; i < imax; ++i)
        for (int j = jmin; j < jmax; ++j)
            grad_node2d(cost, m, n, i, j, &Gnx[(int64_t)j * n + i], &Gny[(int64_t)j * n + i]);
}

/* Windowed normalised gradient value at node (i,j) as the tracer sees it:
 * zero outside the 6x6 window of FastMarching.py:250-252. */
static void grad_win2d(const double *T, int m, int n, double px, double py, int i, int j,
                       double *gx, double *gy) {
    int a = (int)py + 3, b = (int)px + 3;
    int jmax = m < a ? m : a, imax = n < b ? n : b;
    a = (int)(py - 3); b = (int)(px - 3);
    int jmin = a > 0 ? a : 0, imin = b > 0 ? b : 0;
    if (i < imin || i >= imax || j < jmin || j >= jmax) { *gx = 0; *gy = 0; return; }
    grad_node2d(T, m, n, i, j, gx, gy);
}

/* FastMarching.py:305-338 on a (virtual) map given by a node accessor. */
typedef double (*node_fn)(void *ctx, int j, int i);
static double interp2d_fn(double px, double py, int m, int n, node_fn f, void *ctx, int *oob) {
    double fi = trunc(px), fj = trunc(py);
    if (!(fi >= 0) || !(fj >= 0) || fi >= 4294967296.0 || fj >= 4294967296.0) { *oob = 1; return NAN; }
    uint32_t i = (uint32_t)fi, j = (uint32_t)fj;
    double a = px - i, b = py - j;
    /* in-range requirement of the generic branch: j+1 < m and i+1 < n
     * (the `i == n` / `j == m` branches of the reference index past the array) */
    if ((int64_t)i + 1 >= n || (int64_t)j + 1 >= m) { *oob = 1; return NAN; }
    double m00 = f(ctx, j, i), m01 = f(ctx, j, i + 1), m10 = f(ctx, j + 1, i), m11 = f(ctx, j + 1, i + 1);
    double a00 = m00, a10 = m01 - m00, a01 = m10 - m00;
    double a11 = m11 + m00 - m01 - m10;
    if (a == 0) {
        if (b == 0) return a00;
        return a00 + a01 * b;
    }
    if (b == 0) return a00 + a10 * a;
    return a00 + a10 * a + a01 * b + a11 * a * b;
}
typedef struct { const double *p; int n; } plainmap;
static double plain_node(void *ctx, int j, int i) { plainmap *q = (plainmap *)ctx; return q->p[(int64_t)j * q->n + i]; }
double orc_interp2d(const double *map, int m, int n, double px, double py, int *oob) {
    plainmap q = {map, n};
    int o = 0;
    double v = interp2d_fn(px, py, m, n, plain_node, &q, &o);
    if (oob) *oob = o;
    return v;
}
typedef struct { const double *T; int m, n; double px, py; int comp; } gradmap;
static double grad_nodefn(void *ctx, int j, int i) {
    gradmap *g = (gradmap *)ctx; double gx, gy;
    grad_win2d(g->T, g->m, g->n, g->px, g->py, i, j, &gx, &gy);
    return g->comp ? gy : gx;
}

/* Python round(): half to even.  rint() under the default rounding mode. */
static double pyround(double v) { return rint(v); }

/*
 * FastMarching.py:164-236.  out: (cap,2) doubles [x,y]; returns the number of rows
 * written, status in *status.  nsteps = round(15000/tau).
 */
int64_t orc_trace2d(const double *T, int m, int n, const double *init, const double *end,
                    double tau, double *out, int64_t cap, int *status) {
    int64_t K = 0;
    int64_t nsteps = (int64_t)pyround(15000 / tau);
    *status = ORC_TRACE_OK;
#define PUSH(x, y) do { if (K < cap) { out[2 * K] = (x); out[2 * K + 1] = (y); } ++K; } while (0)
    PUSH(init[0], init[1]);
    for (int64_t k = 0; k < nsteps; ++k) {
        double px = out[2 * (K - 1)], py = out[2 * (K - 1) + 1];
        if (isnan(px) || isnan(py)) { *status = ORC_TRACE_VALUEERROR; return K; }   /* int(nan) at :250 */
        if (isinf(px) || isinf(py)) { *status = ORC_TRACE_OVERFLOW; return K; }
        gradmap g = {T, m, n, px, py, 0};
        int oob = 0;
        double dx = interp2d_fn(px, py, m, n, grad_nodefn, &g, &oob);
        g.comp = 1;
        double dy = interp2d_fn(px, py, m, n, grad_nodefn, &g, &oob);
        if (oob) { *status = ORC_TRACE_INDEXERROR; return K; }
        if (isnan(dx) || isnan(dy)) {
            /* :178-218.  Under numpy >= 2 the child scan raises OverflowError at
             * np.uint32([.., -1]) (nearN is a list so `+` concatenates) and the
             * bare except returns the pruned path + nearN, without `end`. */
            *status = ORC_TRACE_EARLY;
            double n0 = pyround(px), n1 = pyround(py);
            for (;;) {
                if (n0 < 0 || n1 < 0 || n0 >= n || n1 >= m) return K;       /* IndexError -> except */
                if (!isinf(T[(int64_t)n1 * n + (int64_t)n0])) break;
                --K;                                                        /* np.delete(gamma,-1) */
                if (K == 0) return K;                                       /* gamma[-1] IndexError -> except */
                n0 = pyround(out[2 * (K - 1)]); n1 = pyround(out[2 * (K - 1) + 1]);
            }
            while (K > 0) {
                double ex = out[2 * (K - 1)] - n0, ey = out[2 * (K - 1) + 1] - n1;
                if (!(sqrt(ex * ex + ey * ey) < 1)) break;
                --K;
            }
            PUSH(n0, n1);
            return K;
        }
        double nx, ny;
        if (sqrt(dx * dx + dy * dy) < 0.01) {          /* np.linalg.norm: sqrt(dot) */
            double s = sqrt(pw2(dx) + pw2(dy));
            nx = dx / s; ny = dy / s;
        } else {
            nx = dx / sqrt(pw2(dx) + pw2(dy));
            ny = dy / sqrt(pw2(nx) + pw2(dy));         /* sic: uses the already-normalised dx (:226-227) */
        }
        PUSH(px - tau * nx, py - tau * ny);
        {
            double ex = out[2 * (K - 1)] - end[0], ey = out[2 * (K - 1) + 1] - end[1];
            if (sqrt(ex * ex + ey * ey) < 1.5) break;
        }
    }
    PUSH(end[0], end[1]);
#undef PUSH
    return K;
}

/* ------------------------------------------------------------- 3D solver -- */
typedef struct {
    int ny, nx, nz;
    const double *cost;
    double *T;
    uint8_t *closed;
    int64_t *cur_seq;
    int64_t seq, live, evals;
    heap_t h;
} front3d;

#define ID3(f, x, y, z) (((int64_t)(y) * (f)->nx + (x)) * (f)->nz + (z))
static inline double f3_T(const front3d *f, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= f->nx || y >= f->ny || z >= f->nz) return INFINITY;
    return f->T[ID3(f, x, y, z)];
}

/* FastMarching3D.py:59-75 -- descending-dimension quadratic solver.
 * Tarray order is [Tx, Ty, Tz]; sums are right-associated like sumlist(). */
double orc_solve3d(double Tx, double Ty, double Tz, double C) {
    double a[3] = {Tx, Ty, Tz};
    int n = 3;
    double Tr = INFINITY;
    while (Tr == INFINITY) {
        if (n == 0) return NAN;                 /* reference: max([]) raises ValueError */
        int im = 0;
        for (int i = 1; i < n; ++i) if (a[i] > a[im]) im = i;      /* max(): first maximum */
        double Tmax = a[im];
        double sumT = 0;
        for (int i = 0; i < n; ++i) { double d = Tmax - a[i]; sumT = sumT + pw2(d); }
        if (pw2(C) > sumT) {
            double S, Q;
            if (n == 3) { S = a[0] + (a[1] + a[2]); Q = a[0] * a[0] + (a[1] * a[1] + a[2] * a[2]); }
            else if (n == 2) { S = a[0] + a[1]; Q = a[0] * a[0] + a[1] * a[1]; }
            else { S = a[0]; Q = a[0] * a[0]; }
            Tr = (S + sqrt(n * pw2(C) + pw2(S) - n * Q)) / n;   /* Q: ndarray**2 -> exact squares */
        }
        for (int i = im; i + 1 < n; ++i) a[i] = a[i + 1];           /* Tarray.remove(Tmax) */
        --n;
    }
    return Tr;
}

/* FastMarching3D.py:19-101 */
static int f3_update(front3d *f, int nx_, int ny_, int nz_) {
    static const int off[6][3] = {{0, 0, -1}, {0, 0, 1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    for (int i = 0; i < 6; ++i) {
        int x = nx_ + off[i][0], y = ny_ + off[i][1], z = nz_ + off[i][2];
        if (x < 0 || y < 0 || z < 0 || x >= f->nx || y >= f->ny || z >= f->nz) continue;
        int64_t id = ID3(f, x, y, z);
        if (f->closed[id]) continue;
        double C = f->cost[id];
        double Tx1 = f3_T(f, x - 1, y, z), Tx2 = f3_T(f, x + 1, y, z);
        double Ty1 = f3_T(f, x, y - 1, z), Ty2 = f3_T(f, x, y + 1, z);
        double Tz1 = f3_T(f, x, y, z - 1), Tz2 = f3_T(f, x, y, z + 1);
        double Tx = Tx1 < Tx2 ? Tx1 : Tx2;
        double Ty = Ty1 < Ty2 ? Ty1 : Ty2;
        double Tz = Tz1 < Tz2 ? Tz1 : Tz2;
        double T = orc_solve3d(Tx, Ty, Tz, C);
        f->evals++;
        if (isinf(f->T[id])) {
            hent e = {T, ++f->seq, id};
            if (heap_push(&f->h, e)) return -1;
            f->cur_seq[id] = e.seq; f->T[id] = T; f->live++;
        } else if (T < f->T[id]) {
            hent e = {T, ++f->seq, id};
            if (heap_push(&f->h, e)) return -1;
            f->cur_seq[id] = e.seq; f->T[id] = T;
        }
    }
    return 0;
}

/* FastMarching3D.py:126-145.  start[0]<0 => full field. */
int orc_fmm3d(const double *cost, int ny, int nx, int nz, const int32_t *goal, const int32_t *start,
              double *T, int64_t *order_out, int64_t *npop, int64_t *evals) {
    front3d f;
    int64_t n = (int64_t)ny * nx * nz, k = 0;
    memset(&f, 0, sizeof(f));
    f.ny = ny; f.nx = nx; f.nz = nz; f.cost = cost; f.T = T;
    f.closed = (uint8_t *)calloc((size_t)n, 1);
    f.cur_seq = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    if (!f.closed || !f.cur_seq) { free(f.closed); free(f.cur_seq); return ORC_ERR_ALLOC; }
    for (int64_t i = 0; i < n; ++i) { T[i] = INFINITY; f.closed[i] = isinf(cost[i]) ? 1 : 0; }
    T[ID3(&f, goal[0], goal[1], goal[2])] = 0.0;
    f.closed[ID3(&f, goal[0], goal[1], goal[2])] = 1;
    f3_update(&f, goal[0], goal[1], goal[2]);
    while (f.h.n) {
        hent e = heap_pop(&f.h);
        if (f.closed[e.idx] || f.cur_seq[e.idx] != e.seq) continue;
        f.live--;
        int z = (int)(e.idx % nz), x = (int)((e.idx / nz) % nx), y = (int)(e.idx / ((int64_t)nz * nx));
        f.closed[e.idx] = 1;
        if (order_out) order_out[k] = e.idx;
        ++k;
        f3_update(&f, x, y, z);
        if (x == start[0] && y == start[1] && z == start[2]) break;
    }
    if (npop) *npop = k;
    if (evals) *evals = f.evals;
    free(f.closed); free(f.cur_seq); free(f.h.a);
    return ORC_OK;
}

/* ------------------------------------------------------------- 3D tracer -- */
/* np.gradient (unit spacing, edge_order=1) of T[y][x][z] along one axis at one node. */
static double npgrad3(const double *T, int ny, int nx, int nz, int y, int x, int z, int axis) {
#define T3(yy, xx, zz) T[((int64_t)(yy) * nx + (xx)) * nz + (zz)]
    int len = axis == 0 ? ny : (axis == 1 ? nx : nz);
    int p = axis == 0 ? y : (axis == 1 ? x : z);
    int dy = axis == 0, dx = axis == 1, dz = axis == 2;
    if (len < 2) return NAN;   /* np.gradient raises; never reached by the planner */
    if (p == 0) return T3(y + dy, x + dx, z + dz) - T3(y, x, z);
    if (p == len - 1) return T3(y, x, z) - T3(y - dy, x - dx, z - dz);
    return (T3(y + dy, x + dx, z + dz) - T3(y - dy, x - dx, z - dz)) / 2.0;
#undef T3
}

typedef double (*node3_fn)(void *ctx, int j, int i, int k);
/* FastMarching3D.py:275-314 (generic branch; the edge branches index past the array) */
static double interp3d_fn(double px, double py, double pz, int m, int n, int o, node3_fn f, void *ctx, int *oob) {
    double fi = trunc(px), fj = trunc(py), fk = trunc(pz);
    if (!(fi >= 0) || !(fj >= 0) || !(fk >= 0) || fi >= 4294967296.0 || fj >= 4294967296.0 || fk >= 4294967296.0) { *oob = 1; return NAN; }
    uint32_t i = (uint32_t)fi, j = (uint32_t)fj, k = (uint32_t)fk;
    if ((int64_t)i + 1 >= n || (int64_t)j + 1 >= m || (int64_t)k + 1 >= o) { *oob = 1; return NAN; }
    double a = px - i, b = py - j, c = pz - k;
    double m000 = f(ctx, j, i, k), m010 = f(ctx, j, i + 1, k), m100 = f(ctx, j + 1, i, k), m001 = f(ctx, j, i, k + 1);
    double m110 = f(ctx, j + 1, i + 1, k), m011 = f(ctx, j, i + 1, k + 1), m101 = f(ctx, j + 1, i, k + 1), m111 = f(ctx, j + 1, i + 1, k + 1);
    double a0 = m000;
    double a1 = m010 - m000;
    double a2 = m100 - m000;
    double a3 = m001 - m000;
    double a4 = m110 + m000 - m010 - m100;
    double a5 = m011 + m000 - m010 - m001;
    double a6 = m101 + m000 - m100 - m001;
    double a7 = m111 + m000 - m100 - m001 - m010;     /* sic (:290) */
    return a0 + a1 * a + a2 * b + a3 * c + a4 * a * b + a5 * a * c + a6 * b * c + a7 * a * b * c;
}
typedef struct { const double *p; int n, o; } plain3;
static double plain3_node(void *ctx, int j, int i, int k) { plain3 *q = (plain3 *)ctx; return q->p[((int64_t)j * q->n + i) * q->o + k]; }
double orc_interp3d(const double *map, int m, int n, int o, double px, double py, double pz, int *oob) {
    plain3 q = {map, n, o};
    int ob = 0;
    double v = interp3d_fn(px, py, pz, m, n, o, plain3_node, &q, &ob);
    if (oob) *oob = ob;
    return v;
}
typedef struct { const double *T; int ny, nx, nz, axis; } grad3;
static double grad3_node(void *ctx, int j, int i, int k) { grad3 *g = (grad3 *)ctx; return npgrad3(g->T, g->ny, g->nx, g->nz, j, i, k, g->axis); }

/*
 * FastMarching3D.py:198-271.  out: (cap,3) doubles [x,y,z].
 */
int64_t orc_trace3d(const double *T, int ny, int nx, int nz, const double *init, const double *end,
                    double tau, double *out, int64_t cap, int *status) {
    int64_t K = 0;
    int64_t nsteps = (int64_t)pyround(15000 / tau);
    *status = ORC_TRACE_OK;
#define PUSH3(x, y, z) do { if (K < cap) { out[3 * K] = (x); out[3 * K + 1] = (y); out[3 * K + 2] = (z); } ++K; } while (0)
#define TT(xx, yy, zz) T[((int64_t)(yy) * nx + (xx)) * nz + (zz)]
    PUSH3(init[0], init[1], init[2]);
    for (int64_t k = 0; k < nsteps; ++k) {
        double px = out[3 * (K - 1)], py = out[3 * (K - 1) + 1], pz = out[3 * (K - 1) + 2];
        double dx, dy, dz;
        int oob = 0;
        int bad = !isfinite(px) || !isfinite(py) || !isfinite(pz);
        if (bad) {
            /* np.uint32(np.fix(inf|nan)) is garbage, the interpolation yields NaN and the
             * fallback's int(round(.)) raises (:213-215): OverflowError for inf, ValueError for nan;
             * components are converted in x,y,z order */
            double q[3] = {px, py, pz};
            for (int c = 0; c < 3; ++c) {
                if (isinf(q[c])) { *status = ORC_TRACE_OVERFLOW; return K; }
                if (isnan(q[c])) { *status = ORC_TRACE_VALUEERROR; return K; }
            }
        }
        grad3 g = {T, ny, nx, nz, 1};
        dx = interp3d_fn(px, py, pz, ny, nx, nz, grad3_node, &g, &oob);
        g.axis = 0;
        dy = interp3d_fn(px, py, pz, ny, nx, nz, grad3_node, &g, &oob);
        g.axis = 2;
        dz = interp3d_fn(px, py, pz, ny, nx, nz, grad3_node, &g, &oob);
        if (oob) { *status = ORC_TRACE_INDEXERROR; return K; }
        if (isnan(dx) || isnan(dy) || isnan(dz)) {
            int64_t n0 = (int64_t)pyround(px), n1 = (int64_t)pyround(py), n2 = (int64_t)pyround(pz);
            for (;;) {
                if (n0 < 0 || n1 < 0 || n2 < 0 || n0 >= nx || n1 >= ny || n2 >= nz) { *status = ORC_TRACE_INDEXERROR; return K; }
                if (!isinf(TT(n0, n1, n2))) break;
                --K;                                             /* gamma = gamma[:-1] */
                if (K == 0) { *status = ORC_TRACE_INDEXERROR; return K; }
                double qx = out[3 * (K - 1)], qy = out[3 * (K - 1) + 1], qz = out[3 * (K - 1) + 2];
                n0 = (int64_t)pyround(qx); n1 = (int64_t)pyround(qy); n2 = (int64_t)pyround(qz);
            }
            while (K > 0) {
                double ex = out[3 * (K - 1)] - (double)n0, ey = out[3 * (K - 1) + 1] - (double)n1, ez = out[3 * (K - 1) + 2] - (double)n2;
                if (!(sqrt(ex * ex + ey * ey + ez * ez) < 1)) break;
                --K;
            }
            PUSH3((double)n0, (double)n1, (double)n2);
            px = (double)n0; py = (double)n1; pz = (double)n2;     /* gamma[-1] is now the node */
            double currentT = TT(n0, n1, n2);
            static const int ch[6][3] = {{0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 0, -1}, {0, 0, 1}};
            for (int i = 0; i < 6; ++i) {
                int64_t c0 = n0 + ch[i][0], c1 = n1 + ch[i][1], c2 = n2 + ch[i][2];
                /* python negative indices wrap; >= size raises IndexError */
                if (c0 >= nx || c1 >= ny || c2 >= nz) { *status = ORC_TRACE_INDEXERROR; return K; }
                int64_t w0 = c0 < 0 ? c0 + nx : c0, w1 = c1 < 0 ? c1 + ny : c1, w2 = c2 < 0 ? c2 + nz : c2;
                double tv = TT(w0, w1, w2);
                if (tv < currentT) {
                    currentT = tv;
                    dx = (double)(n0 - c0) / tau; dy = (double)(n1 - c1) / tau; dz = (double)(n2 - c2) / tau;
                }
            }
        }
        double norm = sqrt(pw2(dx) + pw2(dy) + pw2(dz));
        if (norm < 0.01) {
            double dnx = dx / norm, dny = dy / norm, dnz = dz / norm;
            PUSH3(px - tau * dnx, py - tau * dny, pz - tau * dnz);
        } else {
            PUSH3(px - tau * dx, py - tau * dy, pz - tau * dz);
        }
        {
            double ex = out[3 * (K - 1)] - end[0], ey = out[3 * (K - 1) + 1] - end[1], ez = out[3 * (K - 1) + 2] - end[2];
            if (sqrt(ex * ex + ey * ey + ez * ez) < 1.5) break;
        }
    }
    PUSH3(end[0], end[1], end[2]);
#undef PUSH3
#undef TT
    return K;
}

/* libm pow(x, 2.0) over an array: what `x**2` on a NumPy scalar evaluates to (test helper for the device port) */
void orc_pow2_array(const double *x, double *out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = pw2(x[i]); }

int orc_version(void) { return 1; }
