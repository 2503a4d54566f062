/*
 * tools/sched_model.c -- discrete-event model of the GPU tile scheduler (design tool).
 *
 * Unlike tile_model.c it mirrors the real kernel's timing: a visit samples tile+halo when it
 * STARTS, takes steps*CYC_STEP + CYC_OVH cycles, and its results (and activations) become
 * visible when it COMPLETES; a tile activated while running becomes DIRTY and is re-queued.
 * Policies: 0 FIFO ring; 1 exact priority (min activating value); 2 bucketed levels with a
 * speculation window (workers idle rather than run more than WINDOW levels ahead).
 *   gcc -O2 -o /tmp/sched_model tools/sched_model.c -lm
 *   /tmp/sched_model N workers policy delta_tiles window [mapfile.bin]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TS 32
static int N, NT;
static double *cost, *T;
static double CYC_STEP = 620, CYC_OVH = 30000, CYC_POP = 1500;
static int VARIANT = 0; static double STEP2_FACTOR = 1.25; static int SLICE = 1 << 30; static int PUBSTEP = 0; static double CYC_PUB = 9000;

static inline double eik(double a, double b, double c) {
    double m = a < b ? a : b, d = a - b;
    if (!(fabs(d) <= c)) return m + c;
    return .5 * (a + b + sqrt(2 * (c * c) - d * d));
}
typedef struct { double t; int idx; } he;
static void fmm(double *F, int sx, int sy) {
    char *closed = calloc((size_t)N * N, 1);
    he *hp = malloc(sizeof(he) * (size_t)N * N * 4); int hn = 0;
    for (int i = 0; i < N * N; ++i) { F[i] = INFINITY; if (isinf(cost[i])) closed[i] = 1; }
    F[sy * N + sx] = 0; hp[hn++] = (he){0, sy * N + sx};
    while (hn) {
        he top = hp[0], e = hp[--hn]; int i = 0;
        for (;;) { int l = 2 * i + 1, r = l + 1, m = i; double bt = e.t; if (l < hn && hp[l].t < bt) { m = l; bt = hp[l].t; } if (r < hn && hp[r].t < bt) m = r; if (m == i) break; hp[i] = hp[m]; i = m; }
        if (hn) hp[i] = e;
        if (top.t > F[top.idx]) continue; if (closed[top.idx] && top.t != 0) continue; closed[top.idx] = 1;
        int x = top.idx % N, y = top.idx / N;
        static const int o[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
        for (int k = 0; k < 4; ++k) {
            int cx = x + o[k][0], cy = y + o[k][1];
            if (cx < 0 || cy < 0 || cx >= N || cy >= N) continue;
            int id = cy * N + cx; if (closed[id]) continue;
#define FT(xx, yy) (((xx) < 0 || (yy) < 0 || (xx) >= N || (yy) >= N) ? INFINITY : F[(yy) * N + (xx)])
            double a = fmin(FT(cx - 1, cy), FT(cx + 1, cy)), b = fmin(FT(cx, cy - 1), FT(cx, cy + 1));
            double v = eik(a, b, cost[id]);
            if (v < F[id]) { F[id] = v; int j = hn++; he ne = {v, id}; while (j > 0) { int p = (j - 1) / 2; if (hp[p].t <= ne.t) break; hp[j] = hp[p]; j = p; } hp[j] = ne; }
        }
    }
    free(closed); free(hp);
}

enum { IDLE, QUEUED, RUNNING, DIRTY };
static int *state; static double *prio; static unsigned *saved_mask;
static long total_evals, total_visits, total_steps, armed_sum, armed_n, hist[5], lanes_active_sum;

/* in-tile cell FIM (same as visit2 of tile_model.c); works on a private buffer */
typedef struct { double buf[(TS + 2) * (TS + 2)]; unsigned dirty[TS]; int steps; int unfinished; unsigned left[TS]; int has_mid; double midbuf[(TS + 2) * (TS + 2)]; unsigned middirty[TS]; } visit_t;
static void run_visit(int t, visit_t *V, int sx, int sy) {
    int tx = t % NT, ty = t / NT, P = TS + 2;
    static double cb[TS * TS], nb[(TS + 2) * (TS + 2)];
    unsigned mask[TS], addm[TS]; int lastA[TS], lastB[TS], dirn[TS];
    double *buf = V->buf;
    for (int j = -1; j <= TS; ++j) for (int i = -1; i <= TS; ++i) {
        int x = tx * TS + i, y = ty * TS + j;
        buf[(j + 1) * P + i + 1] = (x < 0 || y < 0 || x >= N || y >= N) ? INFINITY : T[y * N + x];
    }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) { int x = tx * TS + i, y = ty * TS + j; cb[j * TS + i] = (x >= N || y >= N) ? INFINITY : cost[y * N + x]; }
#define B(j, i) buf[((j) + 1) * P + (i) + 1]
    /* arm in natural orientation first (bit i of mask[j] = cell row j col i) */
    unsigned arm[TS]; for (int j = 0; j < TS; ++j) arm[j] = 0;
    int nh = 0, nv = 0;
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
        if (isinf(cb[j * TS + i])) continue;
        double v = B(j, i); int a = 0;
        if (i == 0 && B(j, -1) < v) { a = 1; nh++; }
        if (i == TS - 1 && B(j, TS) < v) { a = 1; nh++; }
        if (j == 0 && B(-1, i) < v) { a = 1; nv++; }
        if (j == TS - 1 && B(TS, i) < v) { a = 1; nv++; }
        int gx = tx * TS + i, gy = ty * TS + j;
        if ((abs(gx - sx) + abs(gy - sy)) == 1 && v > 0) a = 1;
        if (a) arm[j] |= 1u << i;
    }
    int transposed = (VARIANT & 2) && nv > nh;     /* lanes own columns when the inflow is mostly vertical */
    /* line l, position k  ->  cell (j,i) = transposed ? (k,l) : (l,k) */
#define CJ(l, k) (transposed ? (k) : (l))
#define CI(l, k) (transposed ? (l) : (k))
    V->has_mid = 0;
    for (int l = 0; l < TS; ++l) { mask[l] = 0; V->dirty[l] = 0; lastA[l] = 0; lastB[l] = TS - 1; dirn[l] = 1; }
    for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) if (arm[j] >> i & 1) { if (transposed) mask[i] |= 1u << j; else mask[j] |= 1u << i; }
    for (int l = 0; l < TS; ++l) { mask[l] |= saved_mask[t * TS + l]; saved_mask[t * TS + l] = 0; }   /* armed cells left by a time-sliced visit */
    int steps = 0; double cyc = 0;
    for (;;) {
        int any = 0; for (int l = 0; l < TS; ++l) if (mask[l]) any = 1;
        if (!any) break;
        ++steps;
        memcpy(nb, buf, sizeof(nb));
        for (int l = 0; l < TS; ++l) addm[l] = 0;
        int two = 0;
        for (int l = 0; l < TS; ++l) {
            unsigned m = mask[l]; if (!m) continue;
            { int pc = __builtin_popcount(m); armed_sum += pc; armed_n++; hist[pc > 4 ? 4 : pc]++; lanes_active_sum++; }
            int ks[2], nk = 0;
            if (VARIANT & 1) {   /* two cursors: A sweeps up (cyclic), B sweeps down (cyclic) */
                unsigned hi = m & (~0u << lastA[l]); int kA = hi ? __builtin_ctz(hi) : __builtin_ctz(m);
                unsigned lo = m & ((2u << lastB[l]) - 1u); int kB = lo ? 31 - __builtin_clz(lo) : 31 - __builtin_clz(m);
                lastA[l] = kA; lastB[l] = kB; ks[nk++] = kA; if (kB != kA) { ks[nk++] = kB; two = 1; }
            } else {
                unsigned hi = m & (~0u << lastA[l]), lo = m & ((2u << lastA[l]) - 1u); int k;
                if (dirn[l] > 0) { if (hi) k = __builtin_ctz(hi); else { k = 31 - __builtin_clz(lo); dirn[l] = -1; } }
                else { if (lo) k = 31 - __builtin_clz(lo); else { k = __builtin_ctz(hi); dirn[l] = 1; } }
                lastA[l] = k; ks[nk++] = k;
            }
            for (int q = 0; q < nk; ++q) {
                int k = ks[q]; mask[l] &= ~(1u << k);
                int j = CJ(l, k), i = CI(l, k);
                double c = cb[j * TS + i];
                double v = eik(fmin(B(j, i - 1), B(j, i + 1)), fmin(B(j - 1, i), B(j + 1, i)), c); total_evals++;
                if (v < B(j, i)) {
                    nb[(j + 1) * P + i + 1] = v; V->dirty[j] |= 1u << i;
                    /* neighbours in (line,pos) space */
                    int dj[4] = {0, 0, -1, 1}, di[4] = {-1, 1, 0, 0};
                    for (int d = 0; d < 4; ++d) {
                        int jj = j + dj[d], ii = i + di[d];
                        if (jj < 0 || ii < 0 || jj >= TS || ii >= TS) continue;
                        if (!(B(jj, ii) > v) || isinf(cb[jj * TS + ii])) continue;
                        if (VARIANT & 4) {   /* exact filter: v must become the axis-minimum of the neighbour */
                            int oj = jj + dj[d], oi = ii + di[d];
                            if (oj >= -1 && oi >= -1 && oj <= TS && oi <= TS && !(v < B(oj, oi))) continue;
                        }
                        if (transposed) addm[ii] |= 1u << jj; else addm[jj] |= 1u << ii;
                    }
                }
            }
        }
        memcpy(buf, nb, sizeof(nb));
        for (int l = 0; l < TS; ++l) mask[l] |= addm[l];
        cyc += two ? STEP2_FACTOR : 1.0;
        if (PUBSTEP && steps == PUBSTEP) { V->has_mid = 1; memcpy(V->midbuf, buf, sizeof(V->midbuf)); memcpy(V->middirty, V->dirty, sizeof(V->middirty)); }
        if (steps >= SLICE) break;          /* time slice: publish what we have, go again later */
    }
    V->unfinished = 0; for (int l = 0; l < TS; ++l) if (mask[l]) V->unfinished = 1;
    for (int l = 0; l < TS; ++l) V->left[l] = mask[l];
    V->steps = (int)(cyc + 0.5); total_steps += steps; total_visits++;
}

/* queue abstraction */
static int policy; static double DELTA; static int WINDOW;
static int *ring; static long qh, qt;
static he *ph; static int pn;
static void q_push(int t) {
    if (policy == 0) { ring[qt++ % (NT * NT * 4)] = t; return; }
    he e = {prio[t], t}; int i = pn++; while (i > 0) { int q = (i - 1) / 2; if (ph[q].t <= e.t) break; ph[i] = ph[q]; i = q; } ph[i] = e;
}
static double q_min(void) { return pn ? ph[0].t : INFINITY; }
static int q_pop(double min_running) {
    if (policy == 0) { if (qh == qt) return -1; return ring[qh++ % (NT * NT * 4)]; }
    if (!pn) return -1;
    if (policy == 2) {   /* window: do not run more than WINDOW levels ahead of the lowest queued-or-running level */
        double base = fmin(q_min(), min_running);
        if (floor(ph[0].t / DELTA) > floor(base / DELTA) + WINDOW) return -1;
    }
    he top = ph[0], e = ph[--pn]; int i = 0;
    for (;;) { int l = 2 * i + 1, r = l + 1, m = i; double bt = e.t; if (l < pn && ph[l].t < bt) { m = l; bt = ph[l].t; } if (r < pn && ph[r].t < bt) m = r; if (m == i) break; ph[i] = ph[m]; i = m; }
    if (pn) ph[i] = e;
    return top.idx;
}
static void activate(int t, double p) {
    if (state[t] == IDLE) { state[t] = QUEUED; prio[t] = p; q_push(t); }
    else if (state[t] == RUNNING) { state[t] = DIRTY; prio[t] = p; }
    else if (state[t] == DIRTY) { if (p < prio[t]) prio[t] = p; }
    /* QUEUED: keeps its first bucket (as in the planned GPU queue) */
}

int main(int argc, char **argv) {
    N = argc > 1 ? atoi(argv[1]) : 1024; int W = argc > 2 ? atoi(argv[2]) : 1776; policy = argc > 3 ? atoi(argv[3]) : 0;
    double dtiles = argc > 4 ? atof(argv[4]) : 1.0; WINDOW = argc > 5 ? atoi(argv[5]) : 2;
    if (getenv("PUBSTEP")) PUBSTEP = atoi(getenv("PUBSTEP")); if (getenv("SLICE")) SLICE = atoi(getenv("SLICE")); if (getenv("OVH")) CYC_OVH = atof(getenv("OVH")); if (getenv("STEP")) CYC_STEP = atof(getenv("STEP"));
    if (getenv("VARIANT")) VARIANT = atoi(getenv("VARIANT")); if (getenv("STEP2")) STEP2_FACTOR = atof(getenv("STEP2"));
    NT = (N + TS - 1) / TS;
    cost = malloc(sizeof(double) * N * N); T = malloc(sizeof(double) * N * N);
    if (argc > 6) { FILE *f = fopen(argv[6], "rb"); if (!f || fread(cost, 8, (size_t)N * N, f) != (size_t)N * N) { fprintf(stderr, "map read failed\n"); return 1; } fclose(f); }
    else { srand(1); for (int i = 0; i < N * N; ++i) cost[i] = 1.0 + 4.0 * (rand() / (double)RAND_MAX); for (int i = 0; i < N; ++i) cost[i] = cost[(N - 1) * N + i] = cost[i * N] = cost[i * N + N - 1] = INFINITY; }
    double cmin = INFINITY; for (int i = 0; i < N * N; ++i) if (cost[i] < cmin) cmin = cost[i];
    DELTA = dtiles * TS * cmin;
    int sx = N / 4, sy = N / 4;
    while (isinf(cost[sy * N + sx]) || cost[sy * N + sx] > 2) { sx++; }
    state = calloc(NT * NT, sizeof(int)); prio = malloc(sizeof(double) * NT * NT); saved_mask = calloc((size_t)NT * NT * TS, sizeof(unsigned));
    ring = malloc(sizeof(int) * NT * NT * 4); ph = malloc(sizeof(he) * NT * NT * 8);
    for (int i = 0; i < N * N; ++i) T[i] = INFINITY;
    T[sy * N + sx] = 0;
    activate((sy / TS) * NT + sx / TS, 0);
    typedef struct { double t; int tile; visit_t *V; double p; int mid; double tend; } ev;
    ev *run = malloc(sizeof(ev) * W); int nrun = 0;
    visit_t *pool = malloc(sizeof(visit_t) * W); int *freev = malloc(sizeof(int) * W); int nfree = W; for (int i = 0; i < W; ++i) freev[i] = i;
    double now = 0; long maxrun = 0; double busy = 0;
    for (;;) {
        for (;;) {
            if (nrun >= W) break;
            double minrun = INFINITY; for (int i = 0; i < nrun; ++i) if (run[i].p < minrun) minrun = run[i].p;
            int t = q_pop(minrun); if (t < 0) break;
            state[t] = RUNNING;
            ev e; e.tile = t; e.p = prio[t]; e.V = &pool[freev[--nfree]];
            run_visit(t, e.V, sx, sy);
            double dur = e.V->steps * CYC_STEP + CYC_OVH + CYC_POP + (e.V->has_mid ? CYC_PUB : 0);
            e.tend = now + dur; busy += dur; e.mid = 0;
            if (e.V->has_mid) { e.mid = 1; e.t = now + CYC_POP + CYC_OVH * 0.3 + PUBSTEP * CYC_STEP + CYC_PUB; } else e.t = e.tend;
            run[nrun++] = e;
        }
        if (nrun > maxrun) maxrun = nrun;
        if (nrun == 0) break;
        int bi = 0; for (int i = 1; i < nrun; ++i) if (run[i].t < run[bi].t) bi = i;
        ev e = run[bi]; now = e.t;
        int t = e.tile, tx = t % NT, ty = t / NT, P = TS + 2;
        int is_mid = e.mid;
        if (is_mid) { run[bi].mid = 0; run[bi].t = e.tend; } else run[bi] = run[--nrun];
        double *buf = is_mid ? e.V->midbuf : e.V->buf; unsigned *dd = is_mid ? e.V->middirty : e.V->dirty;
        int act[4] = {0, 0, 0, 0}; double ap[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        for (int j = 0; j < TS; ++j) for (int i = 0; i < TS; ++i) {
            if (!(dd[j] >> i & 1)) continue;
            int x = tx * TS + i, y = ty * TS + j; double v = buf[(j + 1) * P + i + 1];
            T[y * N + x] = v;
            if (i == 0 && v < buf[(j + 1) * P]) { act[0] = 1; ap[0] = fmin(ap[0], v); }
            if (i == TS - 1 && v < buf[(j + 1) * P + TS + 1]) { act[1] = 1; ap[1] = fmin(ap[1], v); }
            if (j == 0 && v < buf[i + 1]) { act[2] = 1; ap[2] = fmin(ap[2], v); }
            if (j == TS - 1 && v < buf[(TS + 1) * P + i + 1]) { act[3] = 1; ap[3] = fmin(ap[3], v); }
        }
        if (is_mid) {
            if (act[0] && tx > 0) activate(t - 1, ap[0]);
            if (act[1] && tx < NT - 1) activate(t + 1, ap[1]);
            if (act[2] && ty > 0) activate(t - NT, ap[2]);
            if (act[3] && ty < NT - 1) activate(t + NT, ap[3]);
            continue;
        }
        for (int l = 0; l < TS; ++l) saved_mask[t * TS + l] |= e.V->left[l];
        freev[nfree++] = (int)(e.V - pool);
        if (act[0] && tx > 0) activate(t - 1, ap[0]);
        if (act[1] && tx < NT - 1) activate(t + 1, ap[1]);
        if (act[2] && ty > 0) activate(t - NT, ap[2]);
        if (act[3] && ty < NT - 1) activate(t + NT, ap[3]);
        if (state[t] == DIRTY || e.V->unfinished) { if (state[t] != DIRTY) prio[t] = e.p; state[t] = QUEUED; q_push(t); } else state[t] = IDLE;
    }
    double *F = malloc(sizeof(double) * N * N); fmm(F, sx, sy);
    double maxrel = 0; long nfin = 0, infmis = 0;
    for (int i = 0; i < N * N; ++i) { if (isinf(F[i]) != isinf(T[i])) infmis++; else if (!isinf(F[i]) && F[i] > 0) { double r = fabs(T[i] - F[i]) / F[i]; if (r > maxrel) maxrel = r; nfin++; } }
    printf("N=%d W=%d policy=%d delta=%.1f tiles window=%d: maxrel=%.2g infmis=%ld | evals/cell=%.1f visits/tile=%.2f steps/visit=%.0f | time=%.2f ms (1.9GHz) peak running=%ld avg busy workers=%.0f\n",
           N, W, policy, dtiles, WINDOW, maxrel, infmis, (double)total_evals / nfin, (double)total_visits / (NT * NT), (double)total_steps / total_visits, now / 1.9e6, maxrun, busy / now);
    printf("  armed cells per active lane: mean %.2f; share with 1/2/3/4+ armed: %.2f %.2f %.2f %.2f; active lanes per step %.1f\n", (double)armed_sum / armed_n,
           (double)hist[1] / armed_n, (double)hist[2] / armed_n, (double)hist[3] / armed_n, (double)hist[4] / armed_n, (double)lanes_active_sum / total_steps);
    return 0;
}
