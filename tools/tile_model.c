/*
 * tools/tile_model.c -- host-side *algorithm model* of the GPU tile solver.
 *
 * Design tool, not product and not oracle: it simulates the asynchronous
 * active-tile Fast Iterative Method that eikonal2d.cu implements (one worker per
 * tile visit, in-tile ordered Gauss-Seidel sweeps until a sweep changes nothing,
 * neighbour activation only when a changed edge value undercuts the neighbour's
 * adjacent value) with a discrete-event clock, so that queue discipline, tile size
 * and sweep policy can be compared offline by evaluations per cell, visits per
 * tile and critical-path length.  Result is checked against a heap FMM.
 *
 *   gcc -O2 -o /tmp/tile_model tools/tile_model.c -lm
 *   /tmp/tile_model N tile workers prio map seed
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int N, TS, NT;          /* map side, tile side, tiles per side */
static double *cost, *T;

static inline double eik(double a, double b, double c) {
    double m = a < b ? a : b, d = a - b;
    if (!(fabs(d) <= c)) return m + c;
    return .5 * (a + b + sqrt(2 * (c * c) - d * d));
}

/* ---- reference heap FMM (for checking) ---- */
typedef struct { double t; int idx; } he;
static he *hp; static int hn;
static void hpush(he e) { int i = hn++; while (i > 0) { int p = (i - 1) / 2; if (hp[p].t <= e.t) break; hp[i] = hp[p]; i = p; } hp[i] = e; }
static he hpop(void) { he top = hp[0], e = hp[--hn]; int i = 0; for (;;) { int l = 2 * i + 1, r = l + 1, m = i; double bt = e.t; if (l < hn && hp[l].t < bt) { m = l; bt = hp[l].t; } if (r < hn && hp[r].t < bt) { m = r; } if (m == i) break; hp[i] = hp[m]; i = m; } if (hn) hp[i] = e; return top; }
static void fmm(double *F, int sx, int sy) {
    char *closed = calloc((size_t)N * N, 1);
    hp = malloc(sizeof(he) * (size_t)N * N * 4); hn = 0;
    for (int i = 0; i < N * N; ++i) { F[i] = INFINITY; if (isinf(cost[i])) closed[i] = 1; }
    F[sy * N + sx] = 0; hpush((he){0, sy * N + sx});
    while (hn) {
        he e = hpop(); if (closed[e.idx] && e.t != 0) continue; if (e.t > F[e.idx]) continue; closed[e.idx] = 1;
        int x = e.idx % N, y = e.idx / N;
        static const int o[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
        for (int k = 0; k < 4; ++k) {
            int cx = x + o[k][0], cy = y + o[k][1];
            if (cx < 0 || cy < 0 || cx >= N || cy >= N) continue;
            int id = cy * N + cx; if (closed[id]) continue;
#define FT(xx, yy) (((xx) < 0 || (yy) < 0 || (xx) >= N || (yy) >= N) ? INFINITY : F[(yy) * N + (xx)])
            double a = fmin(FT(cx - 1, cy), FT(cx + 1, cy)), b = fmin(FT(cx, cy - 1), FT(cx, cy + 1));
            double v = eik(a, b, cost[id]);
            if (v < F[id]) { F[id] = v; hpush((he){v, id}); }
        }
    }
    free(closed); free(hp);
}

/* ---- tile machinery ---- */
enum { IDLE = 0, QUEUED = 1 };
static int *state; static double *prio;     /* per tile */
static long total_evals, total_visits, total_sweeps;
static int *visits;

/* queue: either FIFO ring or binary heap on prio */
static int *ring; static long qh, qt; static int use_prio;
static he *ph; static int pn;
static void q_push(int t, double p) {
    if (!use_prio) { ring[qt++ % (NT * NT)] = t; return; }
    he e = {p, t}; int i = pn++; while (i > 0) { int q = (i - 1) / 2; if (ph[q].t <= e.t) break; ph[i] = ph[q]; i = q; } ph[i] = e;
}
static int q_pop(void) {
    if (!use_prio) { if (qh == qt) return -1; return ring[qh++ % (NT * NT)]; }
    while (pn) {
        he top = ph[0], e = ph[--pn]; int i = 0;
        for (;;) { int l = 2 * i + 1, r = l + 1, m = i; double bt = e.t; if (l < pn && ph[l].t < bt) { m = l; bt = ph[l].t; } if (r < pn && ph[r].t < bt) m = r; if (m == i) break; ph[i] = ph[m]; i = m; }
        if (pn) ph[i] = e;
        if (state[top.idx] == QUEUED && prio[top.idx] == top.t) return top.idx;   /* skip stale */
    }
    return -1;
}
static void activate(int t, double p) {
    if (state[t] == QUEUED) { if (use_prio && p < prio[t]) { prio[t] = p; q_push(t, p); } return; }
    state[t] = QUEUED; prio[t] = p; q_push(t, p);
}

/* one visit; returns #sweeps.  buf is (TS+2)^2 with halo */
static int first_dir_policy = 1;
static int visit(int t, int act[4], double actp[4]) {
    int tx = t % NT, ty = t / NT, P = TS + 2;
    static double *buf = NULL, *cb = NULL, *old = NULL;
    if (!buf) { buf = malloc(sizeof(double) * P * P); cb = malloc(sizeof(double) * TS * TS); old = malloc(sizeof(double) * P * P); }
    for (int j = -1; j <= TS; ++j) for (int i = -1; i <= TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j;
        buf[(j + 1) * P + i + 1] = (x < 0 || y < 0 || x >= N || y >= N) ? INFINITY : T[y * N + x];
    }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j;
        cb[j * TS + i] = (x >= N || y >= N) ? INFINITY : cost[y * N + x];
    }
    memcpy(old, buf, sizeof(double) * P * P);
    /* choose first direction from halo minima */
    int d0 = 0;
    if (first_dir_policy) {
        double mL = INFINITY, mR = INFINITY, mU = INFINITY, mD = INFINITY;
        for (int k = 1; k <= TS; ++k) { mL = fmin(mL, buf[k * P]); mR = fmin(mR, buf[k * P + TS + 1]); mU = fmin(mU, buf[k]); mD = fmin(mD, buf[(TS + 1) * P + k]); }
        int sx = (mL <= mR) ? 0 : 1, sy = (mU <= mD) ? 0 : 1;   /* 0 => ascending */
        d0 = sx | (sy << 1);
    }
    int nsw = 0;
    for (int it = 0; it < 64; ++it) {
        int d = d0 ^ ((it & 1) ? 1 : 0) ^ ((it & 2) ? 2 : 0);      /* d0, flip x, flip y, flip both */
        /* order: d0; d0^1; d0^3; d0^2 would be gray; keep simple */
        if ((it & 3) == 2) d = d0 ^ 3; else if ((it & 3) == 3) d = d0 ^ 2;
        int sx = d & 1, sy = (d >> 1) & 1, changed = 0;
        for (int jj = 0; jj < TS; ++jj) for (int ii = 0; ii < TS; ++ii) {
            int i = sx ? TS - 1 - ii : ii, j = sy ? TS - 1 - jj : jj;
            double c = cb[j * TS + i]; if (isinf(c)) continue;
            double *p = &buf[(j + 1) * P + i + 1];
            double a = fmin(p[-1], p[1]), b = fmin(p[-P], p[P]);
            double v = eik(a, b, c);
            total_evals++;
            if (v < *p) { *p = v; changed = 1; }
        }
        ++nsw;
        if (!changed) break;
    }
    /* write back + edge activation */
    for (int k = 0; k < 4; ++k) { act[k] = 0; actp[k] = INFINITY; }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j; if (x >= N || y >= N) continue;
        double v = buf[(j + 1) * P + i + 1];
        if (v < old[(j + 1) * P + i + 1]) {
            T[y * N + x] = v;
            if (i == 0 && v < buf[(j + 1) * P]) { act[0] = 1; actp[0] = fmin(actp[0], v); }
            if (i == TS - 1 && v < buf[(j + 1) * P + TS + 1]) { act[1] = 1; actp[1] = fmin(actp[1], v); }
            if (j == 0 && v < buf[i + 1]) { act[2] = 1; actp[2] = fmin(actp[2], v); }
            if (j == TS - 1 && v < buf[(TS + 1) * P + i + 1]) { act[3] = 1; actp[3] = fmin(actp[3], v); }
        }
    }
    total_sweeps += nsw; total_visits++; visits[t]++;
    return nsw;
}


/* ---- alternative in-tile scheme: warp-level cell FIM with per-row active bitmasks ----
 * lane == row; each step every lane with a non-empty mask picks one active cell of its row,
 * evaluates it against the values as of the previous step (lockstep), and on improvement
 * activates the neighbours that could benefit (value > new).  Converged when all masks are 0. */
static int pick_policy = 1;
static int inner_mode = 0;
static long total_steps;
static int visit2(int t, int act[4], double actp[4], int *steps_out) {
    int tx = t % NT, ty = t / NT, P = TS + 2;
    static double *buf = NULL, *cb = NULL, *nb = NULL;
    static unsigned *mask, *dirty, *addm; static int *last, *dirn;
    if (!buf) { buf = malloc(sizeof(double) * P * P); nb = malloc(sizeof(double) * P * P); cb = malloc(sizeof(double) * TS * TS);
        mask = malloc(4 * TS); dirty = malloc(4 * TS); addm = malloc(4 * TS); last = malloc(4 * TS); dirn = malloc(4 * TS); }
    for (int j = -1; j <= TS; ++j) for (int i = -1; i <= TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j;
        buf[(j + 1) * P + i + 1] = (x < 0 || y < 0 || x >= N || y >= N) ? INFINITY : T[y * N + x];
    }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j;
        cb[j * TS + i] = (x >= N || y >= N) ? INFINITY : cost[y * N + x];
    }
#define B(j, i) buf[((j) + 1) * P + (i) + 1]
    for (int j = 0; j < TS; ++j) { mask[j] = 0; dirty[j] = 0; last[j] = 0; dirn[j] = 1; }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        if (isinf(cb[j * TS + i])) continue;
        double v = B(j, i);
        int a = 0;
        if (i == 0 && B(j, -1) < v) a = 1;
        if (i == TS - 1 && B(j, TS) < v) a = 1;
        if (j == 0 && B(-1, i) < v) a = 1;
        if (j == TS - 1 && B(TS, i) < v) a = 1;
        /* seed: a zero-valued cell (source) activates its neighbours */
        if ((i > 0 && B(j, i - 1) == 0) || (i < TS - 1 && B(j, i + 1) == 0) || (j > 0 && B(j - 1, i) == 0) || (j < TS - 1 && B(j + 1, i) == 0)) if (v > 0) a = 1;
        if (a) mask[j] |= 1u << i;
    }
    int steps = 0;
    for (;;) {
        int any = 0; for (int j = 0; j < TS; ++j) if (mask[j]) any = 1;
        if (!any) break;
        ++steps;
        memcpy(nb, buf, sizeof(double) * P * P);
        for (int j = 0; j < TS; ++j) addm[j] = 0;
        for (int j = 0; j < TS; ++j) {
            unsigned m = mask[j]; if (!m) continue;
            int k;
            if (pick_policy == 0) k = __builtin_ctz(m);
            else {
                unsigned hi = (last[j] >= 31) ? 0 : (m & ~((2u << last[j]) - 1u)) | (m & (1u << last[j]));
                unsigned lo = m & ((1u << last[j]) - 1u) ; lo |= (m & (1u << last[j]));
                if (dirn[j] > 0) { if (hi) k = __builtin_ctz(hi); else { k = 31 - __builtin_clz(lo); dirn[j] = -1; } }
                else { if (lo) k = 31 - __builtin_clz(lo); else { k = __builtin_ctz(hi); dirn[j] = 1; } }
            }
            last[j] = k; mask[j] &= ~(1u << k);
            double c = cb[j * TS + k];
            double a = fmin(B(j, k - 1), B(j, k + 1)), b = fmin(B(j - 1, k), B(j + 1, k));
            double v = eik(a, b, c); total_evals++;
            if (v < B(j, k)) {
                nb[(j + 1) * P + k + 1] = v; dirty[j] |= 1u << k;
                if (k > 0 && B(j, k - 1) > v && !isinf(cb[j * TS + k - 1])) addm[j] |= 1u << (k - 1);
                if (k < TS - 1 && B(j, k + 1) > v && !isinf(cb[j * TS + k + 1])) addm[j] |= 1u << (k + 1);
                if (j > 0 && B(j - 1, k) > v && !isinf(cb[(j - 1) * TS + k])) addm[j - 1] |= 1u << k;
                if (j < TS - 1 && B(j + 1, k) > v && !isinf(cb[(j + 1) * TS + k])) addm[j + 1] |= 1u << k;
            }
        }
        memcpy(buf, nb, sizeof(double) * P * P);
        for (int j = 0; j < TS; ++j) mask[j] |= addm[j];
        if (steps > 100000) { fprintf(stderr, "runaway\n"); exit(1); }
    }
    for (int k = 0; k < 4; ++k) { act[k] = 0; actp[k] = INFINITY; }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        if (!(dirty[j] >> i & 1)) continue;
        int x = tx * TS + i, y = ty * TS + j; if (x >= N || y >= N) continue;
        double v = B(j, i);
        T[y * N + x] = v;
        if (i == 0 && v < B(j, -1)) { act[0] = 1; actp[0] = fmin(actp[0], v); }
        if (i == TS - 1 && v < B(j, TS)) { act[1] = 1; actp[1] = fmin(actp[1], v); }
        if (j == 0 && v < B(-1, i)) { act[2] = 1; actp[2] = fmin(actp[2], v); }
        if (j == TS - 1 && v < B(TS, i)) { act[3] = 1; actp[3] = fmin(actp[3], v); }
    }
#undef B
    total_steps += steps; total_visits++; visits[t]++;
    *steps_out = steps;
    return 1;
}

int main(int argc, char **argv) {
    N = argc > 1 ? atoi(argv[1]) : 1024; TS = argc > 2 ? atoi(argv[2]) : 32;
    int W = argc > 3 ? atoi(argv[3]) : 1776; use_prio = argc > 4 ? atoi(argv[4]) : 0;
    int mapkind = argc > 5 ? atoi(argv[5]) : 0; unsigned seed = argc > 6 ? atoi(argv[6]) : 0;
    first_dir_policy = argc > 7 ? atoi(argv[7]) : 1;
    inner_mode = argc > 8 ? atoi(argv[8]) : 0; pick_policy = argc > 9 ? atoi(argv[9]) : 1;
    NT = (N + TS - 1) / TS;
    cost = malloc(sizeof(double) * N * N); T = malloc(sizeof(double) * N * N);
    srand(seed + 1);
    for (int i = 0; i < N * N; ++i) cost[i] = 1.0 + 4.0 * (rand() / (double)RAND_MAX);
    if (mapkind == 1) {   /* planner-like: plateau 1.0 + blobs of 301 with graded rim */
        for (int i = 0; i < N * N; ++i) cost[i] = 1.0;
        int nb = N * N / 20000 + 3;
        for (int b = 0; b < nb; ++b) {
            int cx = rand() % N, cy = rand() % N, r = N / 40 + rand() % (N / 20 + 1);
            for (int y = cy - 2 * r; y <= cy + 2 * r; ++y) for (int x = cx - 2 * r; x <= cx + 2 * r; ++x) {
                if (x < 0 || y < 0 || x >= N || y >= N) continue;
                double d = sqrt((double)(x - cx) * (x - cx) + (double)(y - cy) * (y - cy));
                double v = d <= r ? 301.0 : (d <= 2 * r ? 1.0 + 10.0 * (2 * r - d) / r : 1.0);
                if (v > cost[y * N + x]) cost[y * N + x] = v;
            }
        }
    }
    for (int i = 0; i < N; ++i) { cost[i] = cost[(N - 1) * N + i] = cost[i * N] = cost[i * N + N - 1] = INFINITY; }
    int sx = N / 4, sy = N / 4;
    cost[sy * N + sx] = fmin(cost[sy * N + sx], 5.0);

    state = calloc(NT * NT, sizeof(int)); prio = malloc(sizeof(double) * NT * NT); visits = calloc(NT * NT, sizeof(int));
    ring = malloc(sizeof(int) * NT * NT); ph = malloc(sizeof(he) * NT * NT * 16);
    for (int i = 0; i < N * N; ++i) T[i] = INFINITY;
    T[sy * N + sx] = 0;
    activate((sy / TS) * NT + sx / TS, 0);

    /* discrete-event simulation: W workers; a visit costs nsw*(2*TS-1)+LOAD steps */
    double *wfree = calloc(W, sizeof(double));
    typedef struct { double t; int tile; int act[4]; double actp[4]; } ev;
    /* pending completions kept in a simple array (W small enough) */
    ev *run = malloc(sizeof(ev) * W); int nrun = 0;
    double now = 0, LOAD = 20;
    long maxq = 0;
    for (;;) {
        /* start as many as possible */
        while (nrun < W) {
            int t = q_pop(); if (t < 0) break;
            state[t] = IDLE;     /* popped: further activations re-queue (model of RUNNING|DIRTY) */
            ev e; e.tile = t;
            if (inner_mode == 0) { int nsw = visit(t, e.act, e.actp); e.t = now + nsw * (2 * TS - 1) + LOAD; total_steps += nsw * (2 * TS - 1); }
            else { int st; visit2(t, e.act, e.actp, &st); e.t = now + st + LOAD; }
            run[nrun++] = e;
        }
        if (nrun > maxq) maxq = nrun;
        if (nrun == 0) break;
        /* advance to earliest completion */
        int bi = 0; for (int i = 1; i < nrun; ++i) if (run[i].t < run[bi].t) bi = i;
        ev e = run[bi]; run[bi] = run[--nrun]; now = e.t;
        int tx = e.tile % NT, ty = e.tile / NT;
        if (e.act[0] && tx > 0) activate(e.tile - 1, e.actp[0]);
        if (e.act[1] && tx < NT - 1) activate(e.tile + 1, e.actp[1]);
        if (e.act[2] && ty > 0) activate(e.tile - NT, e.actp[2]);
        if (e.act[3] && ty < NT - 1) activate(e.tile + NT, e.actp[3]);
    }
    /* NOTE: the model applies a visit's writes at pop time (instant), activations at completion. */
    double *F = malloc(sizeof(double) * N * N); fmm(F, sx, sy);
    double maxrel = 0; long nfin = 0, infmis = 0;
    for (int i = 0; i < N * N; ++i) {
        if (isinf(F[i]) != isinf(T[i])) infmis++;
        else if (!isinf(F[i]) && F[i] > 0) { double r = fabs(T[i] - F[i]) / F[i]; if (r > maxrel) maxrel = r; nfin++; }
    }
    int vmax = 0; for (int i = 0; i < NT * NT; ++i) if (visits[i] > vmax) vmax = visits[i];
    printf("N=%d TS=%d W=%d prio=%d map=%d: maxrel=%.3g infmis=%ld | evals/cell=%.2f visits/tile=%.2f (max %d) sweeps/visit=%.2f steps/visit=%.1f | crit steps=%.0f (%.1f us @125cyc/1.9GHz) peak running=%ld\n",
           N, TS, W, use_prio, mapkind, maxrel, infmis, (double)total_evals / nfin, (double)total_visits / (NT * NT), vmax,
           (double)total_sweeps / total_visits, (double)total_steps / total_visits, now, now * 125 / 1.9e3, maxq);
    return 0;
}
