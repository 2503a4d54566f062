/*
 * fm_b200.h -- C ABI of the B200-native Fast Marching replacement
 * (libfm_b200.so, built from planning_motion_planning_b200/csrc/ for sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of
 * esa-prl/planning-motion_planning: the Eikonal solve of src/FastMarching and the
 * gradient-descent path extraction over its result.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repo):
 *
 *   fmb_solve2d_*        src/FastMarching/FastMarching.py:92-112   computeTmap (full field)
 *                        src/FastMarching/FastMarching.py:114-162  biComputeTmap (two solves, nq = 2)
 *                        src/FastMarching/FastMarching.py:17-29,44-80  getEikonal / updateNode
 *   fmb_solve3d_*        src/FastMarching/FastMarching3D.py:126-145 computeTmap
 *                        src/FastMarching/FastMarching3D.py:19-101  updateNode (descending-dimension solver)
 *   fmb_trace2d_*        src/FastMarching/FastMarching.py:164-236  getPathGDM
 *                        src/FastMarching/FastMarching.py:242-338  computeGradient / interpolatePoint
 *   fmb_trace3d_*        src/FastMarching/FastMarching3D.py:198-314 getPathGDM / interpolatePoint
 *   fmb_truncate2d_* /   the early-exit semantics of the reference loops
 *   fmb_truncate3d_*     (FastMarching.py:108-109,150-155; FastMarching3D.py:141-142): turn a full
 *                        field + pop ranks into the PARTIAL field the reference returns
 *
 * Conventions
 *  - Plain C, no torch types.  Every data pointer is a DEVICE pointer on the
 *    current CUDA device (e.g. torch.Tensor.data_ptr()); `stream` is a
 *    cudaStream_t passed as void* (0 = legacy default stream).
 *  - Arrays are indexed [y][x] (2D) / [y][x][z] (3D, z contiguous) exactly like the
 *    reference's NumPy arrays; nodes are [x,y] / [x,y,z] (reference convention).
 *    `pitch` arguments are row strides in ELEMENTS; `qstride` is the element
 *    stride between consecutive queries of a batch (0 = all queries share it).
 *  - All calls are asynchronous on `stream` unless stated; they never allocate
 *    device memory: the caller supplies a workspace of fmb_workspace_bytes_*().
 *  - Return value: 0 = ok, nonzero = FMB_E_* ; fmb_last_error() (thread local)
 *    gives a message.  Device-side failures (watchdog, step cap) are reported by
 *    fmb_finish(), which synchronises the stream.
 *  - There is NO CPU fallback anywhere behind this interface.
 */
#ifndef FM_B200_H
#define FM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMB_OK 0
#define FMB_E_INVALID 1       /* bad argument */
#define FMB_E_CUDA 2          /* CUDA runtime error */
#define FMB_E_WORKSPACE 3     /* workspace too small */
#define FMB_E_WATCHDOG 4      /* device-side watchdog fired (would have hung) */
#define FMB_E_STEPCAP 5       /* in-tile iteration cap exceeded */
#define FMB_E_NOJOIN 6        /* bi-solve: fronts never meet (reference raises NameError) */

/* tracer status per path (observable behaviour of the reference's getPathGDM) */
#define FMB_TRACE_OK 0          /* `end` appended, normal return */
#define FMB_TRACE_EARLY 1       /* 2D NaN fallback: bare `except` returns the path so far (FastMarching.py:217-218) */
#define FMB_TRACE_VALUEERROR 2  /* NaN waypoint: reference raises ValueError */
#define FMB_TRACE_INDEXERROR 3  /* stencil left the array: reference raises IndexError */
#define FMB_TRACE_OVERFLOW 4    /* inf waypoint: reference raises OverflowError */

/* solver counters, filled by fmb_finish() */
typedef struct fmb_stats {
    uint64_t tile_visits;   /* tile activations processed */
    uint64_t steps;         /* lock-step iterations summed over visits */
    uint64_t evals;         /* local-solver evaluations (getEikonal / 3D quadratic) */
    uint64_t pushes;        /* work-queue pushes */
    uint64_t cells_written; /* T values stored to HBM */
    double solve_kernel_ms; /* device time of the last persistent solve kernel on this thread (CUDA events) */
    double init_kernel_ms;  /* device time of the init (fill + seed) launches that preceded it */
    uint64_t cyc_wait, cyc_load, cyc_relax, cyc_store; /* warp-cycles per phase, summed over worker warps (2D) */
    uint64_t reserved[1];
    uint64_t cyc_check;     /* sweep engine: cycles of the check passes (part of cyc_relax), thread 0 of every CTA */
    uint64_t noop_visits;   /* sweep engine: visits whose first check pass found the tile already at its fixed point */
    uint64_t rounds;        /* sweep engine: rounds of four sweeps */
    uint64_t continuations; /* sweep engine: re-activations of a running tile served in place */
} fmb_stats;

/* Tunables of the solvers: process-wide, read by every solve call (no getenv() on the call path; the FMB_*
 * environment variables only give the INITIAL values, read once at first use).  0 / -1 = automatic. */
typedef struct fmb_options {
    int32_t engine2d;     /* 0 auto, 1 warp-per-tile armed-cell visits (round 1), 2 CTA-per-tile Jacobi visits, 3 four-warp sweep visits,
                             4 / 5 warp-per-tile sweep visits for batches (costs staged in shared memory / read from global),
                             6 = 1 without the shared-memory cost tile (best-first batches: more resident warps; their default) */
    int32_t cta_cells;    /* cells per thread of the CTA engine: 0 auto, 1, 2 or 4 */
    int32_t tile_w2d;     /* tile width of the warp engine: 16 or 32 */
    int32_t tile_z3d;     /* 3D tile depth: 16 or 32 */
    int32_t best_first;   /* per-query best-first order: -1 auto, 0 off, 1 on */
    int32_t windowed;     /* windowed FIFO order of one large map: -1 auto, 0 off, 1 on */
    int32_t window;       /* window in levels (-1 = default) */
    int32_t worker_div;   /* tiles per worker used to size the grid (0 = automatic) */
    int32_t max_blocks;   /* cap on the persistent grid (0 = none) */
    int32_t watchdog_ms;  /* device watchdog, time without progress */
    int32_t step_cap;     /* in-tile iteration cap */
    int32_t engine3d;     /* 0 auto, 1 warp-per-tile visits (round 1), 3 eight-warp sweep visits */
    int32_t level_div;    /* windowed order: levels per tile crossing at the source's cost (0 = default) */
    int32_t win_running;  /* windowed order: running tiles hold their level (-1 auto, 0 off, 1 on) */
    int32_t check_passes; /* sweep engine: Jacobi check passes tried before another round of sweeps (0 = default) */
    int32_t pipeline;     /* sweep engine, one map: pipelined visits (-1 auto, 0 off, 1 on) */
    int32_t precheck;     /* sweep engine: open every visit with a check pass (-1 auto, 0, 1) */
    int32_t causal_slack; /* sweep engines, local causal order: a tile only waits for a neighbour whose priority lies more than
                             this many percent of one tile crossing (at the seed's cost) below its own (0 = default, -1 = none) */
    int32_t tma;          /* 2D sweep engine: stage interior tiles with TMA tensor copies (-1 auto = on, 0 off, 1 on) */
    int32_t ring2;        /* 2D sweep engine, local causal order: also wait for a queued / running tile of the SECOND ring whose priority
                             lies more than this many percent of one tile crossing below mine (0 = default: 200 in 2D, off in 3D; -1 = off) */
    int32_t variant;      /* sweep engines: bit 0 = straight-line sweep step (predication, no vote / branches), bit 1 (2D) = the cell's current value
                             is re-read right before the store, bit 3 (3D) = octant rule: a cell is evaluated only by the sweep whose upwind side carries
                             its lower neighbour on every axis, sweeps without such a cell skip the round (36 -> 12 evaluations per cell, +10 % time)
                             (0 = default: 3 in 2D, 1 in 3D; -1 = the branching step) */
    int32_t concurrent_solves; /* 2D sweep engine, one map per solve: how many solves the caller keeps in flight on this GPU (one stream
                             each).  A solve in causal order is a chain of dependent visits that leaves most issue slots idle; its
                             persistent grid then takes 1/n of the resident CTA slots, so all n solves are resident at once
                             (0 / 1 = default: two CTAs per SM, the fastest grid for one solve alone) */
    int32_t replay_sparse; /* early-exit emulation: replay only the relaxations the narrow band depends on (-1 auto: from 2^21 cells on,
                             0 always the dense replay, 1 always the sparse one; it falls back to the dense one on the device) */
} fmb_options;
void fmb_get_options(fmb_options *out);
int fmb_set_options(const fmb_options *in);

/* Self-test: counts (into *d_bad, zeroed by the caller) the inputs of d_x[0..n) inside the range of the library's
 * branch-free fp64 square root whose result differs in any bit from sqrt.rn.f64.  Must stay 0. */
int fmb_debug_sqrt_check(const double *d_x, int64_t n, uint64_t *d_bad, void *stream);
/* The same for the branch-free fp64 division of the tracer's fast step: d_bad2[0] += pairs (a, b) the fast form accepts
 * whose quotient differs in any bit from div.rn.f64 (must stay 0), d_bad2[1] += accepted pairs (both zeroed by the caller). */
int fmb_debug_div_check(const double *d_a, const double *d_b, int64_t n, uint64_t *d_bad2, void *stream);

int fmb_version(void);
const char *fmb_last_error(void);
/* number of SMs of the current device (grid sizing / reporting) */
int fmb_sm_count(void);

/* ---- 2D Eikonal solve (full field per query) ------------------------------
 * nq independent queries of identical shape rows x cols.
 *   d_cost   cost map(s); +inf = obstacle; out-of-domain is treated as +inf
 *   d_T      output field(s); fully overwritten (unreached cells = +inf)
 *   d_seeds  int32 [nq][2] = [x,y] of the source (T = 0) of each query
 *   d_ws     workspace of at least fmb_workspace_bytes_2d(rows, cols, nq) bytes
 */
size_t fmb_workspace_bytes_2d(int rows, int cols, int nq);
int fmb_solve2d_f64(const double *d_cost, int64_t cost_pitch, int64_t cost_qstride,
                    double *d_T, int64_t T_pitch, int64_t T_qstride,
                    int rows, int cols, int nq, const int32_t *d_seeds,
                    void *d_ws, size_t ws_bytes, void *stream);
int fmb_solve2d_f32(const float *d_cost, int64_t cost_pitch, int64_t cost_qstride,
                    float *d_T, int64_t T_pitch, int64_t T_qstride,
                    int rows, int cols, int nq, const int32_t *d_seeds,
                    void *d_ws, size_t ws_bytes, void *stream);

/* One 2D solve whose cost map is still on the host: the upload runs on `copy_stream` in bands of rows ordered by their
 * distance from the goal while the solve already runs on `stream` (its CTAs wait per band on device flags the copy stream
 * sets; every copy is queued before the kernel, so nothing waits for work queued behind it).  h_cost: dense [rows][cols]
 * PAGE-LOCKED host memory (FMB_E_INVALID otherwise); d_cost: device buffer of the same shape, complete when `stream`
 * reaches the end of the call's work; goal_xy: HOST int32[2].  Replaces "upload, then FastMarching.computeTmap's full
 * field" of the drop-in (FastMarching.py:92-112 as intended) when the caller's map is page-locked: 4096^2 on B200 hides
 * the 2.5 ms upload behind the 7.3 ms solve.  d_ws: fmb_workspace_bytes_2d_h2d bytes; fmb_finish(d_ws, ...) as usual. */
size_t fmb_workspace_bytes_2d_h2d(int rows, int cols);
int fmb_solve2d_h2d_f64(const double *h_cost, double *d_cost, int rows, int cols, const int32_t *goal_xy, double *d_T,
                        void *d_ws, size_t ws_bytes, void *stream, void *copy_stream);

/* Resume a single 2D solve from the CURRENT contents of d_T (no re-initialisation): used by the
 * row-slab domain decomposition of very large maps, where halo rows received from the neighbour
 * slabs are written into d_T between calls (they carry cost = +inf locally, so they are inputs
 * only).  d_seed: int32[2] [x,y] of a source inside this array, or an out-of-range node for none.
 * activate: bit 0 = queue the first tile row, bit 1 = the last tile row, bit 2 = every tile.
 * halo_rows: bit 0 / bit 1 = array row 0 / rows-1 is such a halo row (its tiles re-arm all cells).
 * The reference has no counterpart (single process, FastMarching.py:92-112). */
int fmb_resolve2d_f64(const double *d_cost, int64_t cost_pitch, double *d_T, int64_t T_pitch,
                      int rows, int cols, const int32_t *d_seed, int activate, int halo_rows,
                      void *d_ws, size_t ws_bytes, void *stream);

/* ---- 3D Eikonal solve ------------------------------------------------------
 * volumes are dense [ny][nx][nz] arrays with z contiguous.
 * d_seeds int32 [nq][3] = [x,y,z].
 */
size_t fmb_workspace_bytes_3d(int ny, int nx, int nz, int nq);
int fmb_solve3d_f64(const double *d_cost, int64_t cost_qstride, double *d_T, int64_t T_qstride,
                    int ny, int nx, int nz, int nq, const int32_t *d_seeds,
                    void *d_ws, size_t ws_bytes, void *stream);
int fmb_solve3d_f32(const float *d_cost, int64_t cost_qstride, float *d_T, int64_t T_qstride,
                    int ny, int nx, int nz, int nq, const int32_t *d_seeds,
                    void *d_ws, size_t ws_bytes, void *stream);

/* Polish pass: re-relaxes EVERY cell of the converged field(s) in d_T with the reference's own rounding of the squares
 * it takes on NumPy scalars (FastMarching3D.py:68-71: `**2` = libm pow(x, 2.0), not correctly rounded; csrc/pow2_glibc.cuh
 * reproduces it bit for bit) until the field is the exact fixed point of THAT update.  The values move by an ulp in a
 * fraction of a percent of the cells: irrelevant for the 1e-9 tolerance, decisive for exact ties of the pop order
 * (early exit of FastMarching3D.computeTmap :141-142 on uniform-cost volumes).  Same arguments as fmb_solve3d_f64,
 * d_T = its result; asynchronous, fmb_finish() reports. */
/* The whole solve with that exact arithmetic from the start (same result as solve + polish; faster on large volumes,
 * where the polish pass re-relaxes nearly every cell anyway). */
int fmb_solve3d_exact_f64(const double *d_cost, int64_t cost_qstride, double *d_T, int64_t T_qstride,
                          int ny, int nx, int nz, int nq, const int32_t *d_seeds,
                          void *d_ws, size_t ws_bytes, void *stream);
int fmb_polish3d_f64(const double *d_cost, int64_t cost_qstride, double *d_T, int64_t T_qstride,
                     int ny, int nx, int nz, int nq, const int32_t *d_seeds,
                     void *d_ws, size_t ws_bytes, void *stream);

/* Synchronise `stream`, report device-side failures of the solves issued with
 * this workspace, and (optionally) return the counters.  Synchronous. */
int fmb_finish(void *d_ws, size_t ws_bytes, void *stream, fmb_stats *stats_or_null);

/* ---- gradient-descent path extraction --------------------------------------
 * npaths paths; path p walks field  d_T + field_of_path[p]*T_qstride
 * (field_of_path == NULL: p itself).
 *   d_init, d_end  double [npaths][D] waypoints in cell units, [x,y(,z)]
 *   d_out          double [npaths][cap][D]; first row = init
 *   d_count        int32  [npaths] rows written
 *   d_status       int32  [npaths] FMB_TRACE_*
 * max_steps = round(15000/tau) reproduces the reference; cap >= max_steps + 2.
 */
int fmb_trace2d_f64(const double *d_T, int64_t T_pitch, int64_t T_qstride, int rows, int cols,
                    int npaths, const int32_t *d_field_of_path,
                    const double *d_init, const double *d_end, double tau, int max_steps,
                    double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status, void *stream);
int fmb_trace3d_f64(const double *d_T, int64_t T_qstride, int ny, int nx, int nz,
                    int npaths, const int32_t *d_field_of_path,
                    const double *d_init, const double *d_end, double tau, int max_steps,
                    double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status, void *stream);

/* ---- early-exit emulation ------------------------------------------------------
 * The reference returns partial fields: cells accepted within the first k pops are
 * final, narrow-band cells hold the tentative value of their last relaxation, the
 * rest is +inf.  Given the full field d_F, the costs, and d_rank (int32 pop rank of
 * every cell: source 0, unreached INT32_MAX; a stable ascending sort of d_F), these
 * rebuild that partial field for truncation after k pops.  Dense arrays.
 * The kernel replays every relaxation of the first k pops in pop order, in parallel (csrc/truncate.cuh).
 * d_list: int32 scratch with room for one entry per cell (receives the cell popped r-th for r <= k);
 * d_counters: int32[2], zeroed by the caller; [0] is the ticket counter of the replay, on return
 * [1] = dependency waits that hit the safety limit (0 in every test; non-zero means an inexact value).
 * d_memo: scratch of 4 (2D) / 6 (3D) doubles per cell, contents ignored on entry: the tentative value a
 * cell held after each of its neighbours popped.
 */
int fmb_truncate2d_f64(const double *d_F, const double *d_cost, const int32_t *d_rank, int rows, int cols,
                       int32_t k, double *d_out, int32_t *d_list, int32_t *d_counters, double *d_memo, void *stream);
int fmb_truncate3d_f64(const double *d_F, const double *d_cost, const int32_t *d_rank, int ny, int nx, int nz,
                       int32_t k, double *d_out, int32_t *d_list, int32_t *d_counters, double *d_memo, void *stream);

/* One refinement step of the reference's pop order among exactly equal T values (the reference
 * keeps its narrow band with bisect_left + insert, FastMarching.py:65-67,76-78: last inserted pops
 * first).  Inputs: full field, costs, current int32 pop ranks and insertion times, tie-group id of
 * every cell (dense rank of its T value); outputs the new insertion times and a packed 64-bit key
 * whose ascending stable sort gives the next ranks.  Iterated to a fixed point by the caller. */
int fmb_tie_keys2d_f64(const double *d_T, const double *d_cost, const int32_t *d_rank, const int32_t *d_tau,
                       const int32_t *d_group, int rows, int cols, int32_t seed_index, int32_t transposed,
                       int32_t *d_tau_new, int64_t *d_key, void *stream);
/* The same order in ONE launch, without iteration: only strictly upwind neighbours (smaller T) take part
 * in a cell's final update and they pop before the cell's tie group starts, so insertion times and
 * in-group ranks are settled group by group in ascending T by a ticketed sweep with dependency waits.
 * d_members lists the cells in ascending T (stable), d_gstart / d_gsize give every cell the slice of its
 * tie group in that list (size 1 for unreached cells; groups larger than 4096 cells are the caller's cue to
 * use the iterated form); outputs d_rank (unreached = INT32_MAX) and d_tau; d_key: int64 scratch per cell;
 * d_scratch: 2*rows*cols + 2 int32, on return its last entry counts waits that hit the safety limit.
 * transposed != 0: d_T is the transpose of the caller's map (an F-ordered input solved as its C-ordered
 * transpose); the child order of updateNode (FastMarching.py:46-54) is not symmetric in x and y. */
int fmb_tie_order2d_f64(const double *d_T, const double *d_cost, const int32_t *d_members, const int32_t *d_gstart,
                        const int32_t *d_gsize, int rows, int cols, int32_t seed_index, int32_t transposed, int32_t *d_rank,
                        int32_t *d_tau, int64_t *d_key, int32_t *d_scratch, void *stream);
/* 3D form (FastMarching3D.py:22-33 child order z-1, z+1, x-1, x+1, y+1, y-1; :77-95 the same bisect_left
 * insertion): d_T is [ny][nx][nz], seed_index = (y*nx + x)*nz + z. */
int fmb_tie_order3d_f64(const double *d_T, const double *d_cost, const int32_t *d_members, const int32_t *d_gstart,
                        const int32_t *d_gsize, int ny, int nx, int nz, int32_t seed_index, int32_t *d_rank, int32_t *d_tau,
                        int64_t *d_key, int32_t *d_scratch, void *stream);

/* Where the two fronts of biComputeTmap meet (FastMarching.py:141-155): k = min over cells of
 * max(rankG, rankS) = the round in which the reference's alternating loop breaks, and the join node
 * (the cell G popped in round k if it attains the minimum -- G is tested first, :150-152 -- else S's).
 * d_rankG / d_rankS: int32 pop ranks of the two full fields (unreached = INT32_MAX).
 * d_out: int32[4], 8-byte aligned; on return [0] = k, [1] = flat index of the join cell (both INT32_MAX
 * when the fronts never meet: the reference raises NameError, :161); [2..3] scratch. */
int fmb_bi_join(const int32_t *d_rankG, const int32_t *d_rankS, int64_t total, int32_t *d_out, void *stream);

/* ---- 2D cost-map construction (the step right before the 2D solve) --------
 * Replaces Coupled_motion_planner.py:37-80 (surface_normal), :83-95 (image_filling), :97-109
 * (structural_disk) and the inline pipeline of main() :1144-1216: slope obstacles from the DEM,
 * hole filling, opening by a disk of r_open, dilate / fill / erode by r_close, distance band
 * (dilation by r_expand x exact Euclidean distance), 1 + 300*obstacle + 10*band, 50x50 box blur
 * with 300 outside, +inf map limits.
 *   d_dem     n x n zero-based DEM (the planner's Zs after :1101), C order
 *   d_grid    n doubles: np.linspace(0, size, n) (the meshgrid axis of :41-43)
 *   d_cost    out, n x n in [y][x] order == what the planner passes to biComputeTmap (cMap.T, :1226)
 *   d_obst_raw / d_obst / d_pre   optional outs: obstacle map after the first filling, final obstacle
 *             map (uint8), cost before the blur ([y][x]); NULL to skip
 * Asynchronous on `stream`; fmb_costmap2d_finish synchronises and returns the number of cells with a
 * positive distance-band value (0 means the reference would raise ValueError at :1198). */
size_t fmb_workspace_bytes_costmap2d(int n);
int fmb_costmap2d_f64(const double *d_dem, const double *d_grid, int n, double resolution, double slope_max,
                      int r_open, int r_close, int r_expand, double *d_cost, uint8_t *d_obst_raw, uint8_t *d_obst,
                      double *d_pre, void *d_ws, size_t ws_bytes, void *stream);
int fmb_costmap2d_finish(void *d_ws, size_t ws_bytes, void *stream, int32_t *n_positive);

/* ---- 3D arm-workspace cost volume (feeds fmb_solve3d) ----------------------
 * Replaces GetObstMap (Coupled_motion_planner.py:319-358), TunnelCost (:505-725) and the product
 * Cmap1*Cmap2 (:1627).  The host wrapper prepares the small tables with the reference's scalar
 * expressions; the device resolves the order-dependent scatter ("first writer wins", +inf always
 * wins) with sequence numbers and writes the volumes.  All pointers are device pointers.
 *   d_Zs        DEM crop, zs_rows x zs_cols (the planner's ZsMap, :1523-1526)
 *   d_frames    (npose + 1) x 12 doubles: rows 0..2 of the base frame Toa per pose (:530-533), then
 *               the frame of the closing half sphere (:661-664)
 *   d_li, d_lk  np.linspace(-R, R, nX / nZ) of the tube cross-section (:546-547), R = rlim + 2 resX
 *   d_norm, d_val  nX x nZ tables: sqrt(i**2 + k**2) (:560) and the graded cost (:573)
 *   d_lr, d_hval   nK radii of the half sphere (:676) and their cost (:697)
 *   d_angles    cos(theta)[100], sin(theta)[100], cos(sigma)[90], sin(sigma)[90] (:668-681)
 *   fin / ini   sample node and initial end-effector node [x, y, z] (never blocked, :576, :641, :720)
 * Outputs (any may be NULL): d_cmap = terrain * tunnel, d_tunnel (TunnelCost's Cmap, shape sY x sX x sZ),
 * d_terrain (GetObstMap's finalMap, shape sX x sY x sZ).  Asynchronous on `stream`. */
typedef struct fmb_costvolume_desc {
    const double *d_Zs; int32_t zs_rows, zs_cols;
    double resX, resY, resZ, xm, ym;
    int32_t sX, sY, sZ;
    const double *d_frames; int32_t npose;
    const double *d_li, *d_lk; int32_t nX, nZ;
    const double *d_norm, *d_val; double rlim;
    const double *d_lr, *d_hval; int32_t nK;
    const double *d_angles; double shell;
    int64_t fin[3], ini[3];
} fmb_costvolume_desc;
size_t fmb_workspace_bytes_costvolume(int sX, int sY, int sZ);
int fmb_costvolume_f64(const fmb_costvolume_desc *desc, double *d_cmap, double *d_tunnel, double *d_terrain,
                       void *d_ws, size_t ws_bytes, void *stream);

/* ---- path post-processing (the step right after the tracers) -------------------------
 * Batched, row counts read from the tracer's device-side counters (d_count of fmb_trace*_f64), asynchronous.
 *
 * fmb_path_stitch2d_f64: Coupled_motion_planner.py:1232-1234 -- out = resolution * (vstack(flipud(pathS), pathG[1:]) + 1)
 *   d_pathS / d_pathG [npairs][cap][2], d_out [npairs][2*cap][2], d_count_out[p] = countS[p] + countG[p] - 1.
 * fmb_path_post3d_f64: Coupled_motion_planner.py:1641-1671 -- per-axis scaling (scale3 = resX, resY, resZ; host array),
 *   scipy.signal.savgol_filter(., 11, 3) with its default mode='interp' (cubic fitted to the first / last 11 samples at
 *   the ends), + offset3 (host array: Xmin, Ymin, Zmin), last row := d_last[p] (device [npaths][3], or NULL), then
 *   interp1d(range(n), .)(np.linspace(0, n - 1, m)) (np.interp semantics).  d_out [npaths][m][3];
 *   d_status[p] = 1 when the path has fewer than 11 rows (scipy raises ValueError; the rows are NaN then).
 *   d_ws: fmb_workspace_bytes_pathpost() bytes. */
/* fmb_path_pack_f64: the written rows of npaths tracer slabs ([npaths][cap][dim], d_count rows each) packed into one dense
 * array, path p at row d_offsets[p] (int64, e.g. the exclusive prefix sum of d_count): d_out[row] = scale * (waypoint + shift)
 * (scale = 1, shift = 0 copies the waypoints as traced; scale = resolution, shift = 1 gives the planner's metres, :1234).
 * Lets a batch hand K x dim waypoints to the host instead of cap-row slabs. */
int fmb_path_pack_f64(const double *d_paths, const int32_t *d_count, const int64_t *d_offsets, int64_t cap, int npaths, int dim,
                      double scale, double shift, double *d_out, void *stream);
size_t fmb_workspace_bytes_pathpost(void);
/* *d_flag = 1 when the two device arrays differ in any bit, else 0 (asynchronous).  The drop-in keeps the device copy
 * of a field it returned; when the caller hands that array to getPathGDM (Coupled_motion_planner.py:1229-1230) the tracer
 * starts on the copy while the array is uploaded again, and this check decides whether the path may be used. */
int fmb_fields_differ_f64(const double *d_a, const double *d_b, int64_t n, int32_t *d_flag, void *stream);
/* The same question without moving the array: fmb_trace2d_logged_f64 = fmb_trace2d_f64 that also logs which cells of the
 * field the path read (the origin node of every 16 x 16 gradient block it staged: d_blocks [npaths][cap_blocks][2] = bx, by;
 * d_nblocks[p] = count, cap_blocks + 1 on overflow, -1 for fields too small for blocks), and fmb_windows_differ_f64
 * compares exactly those 18 x 18 windows of the traced device field with the caller's array, which the kernel reads in
 * place: h_field = page-locked host memory addressed by the device through the same pointer (cudaHostAlloc under
 * unified addressing; FMB_E_INVALID otherwise), element (y, x) at h_field[y * hs_y + x * hs_x].  *d_flag = 1 on any
 * difference, overflow or missing log.  One path (the log of path 0 is used). */
int fmb_trace2d_logged_f64(const double *d_T, int64_t T_pitch, int64_t T_qstride, int rows, int cols, int npaths,
                           const int32_t *d_field_of_path, const double *d_init, const double *d_end, double tau,
                           int max_steps, double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status,
                           int32_t *d_blocks, int32_t *d_nblocks, int cap_blocks, void *stream);
int fmb_windows_differ_f64(const double *d_field, int64_t pitch, int rows, int cols, const double *h_field, int64_t hs_y, int64_t hs_x,
                           const int32_t *d_blocks, const int32_t *d_nblocks, int cap_blocks, int32_t *d_flag, void *stream);
int fmb_path_stitch2d_f64(const double *d_pathS, const int32_t *d_countS, const double *d_pathG, const int32_t *d_countG,
                          int64_t cap, int npairs, double resolution, double *d_out, int32_t *d_count_out, void *stream);
int fmb_path_post3d_f64(const double *d_paths, const int32_t *d_count, int64_t cap, int npaths, const double *scale3,
                        const double *offset3, const double *d_last, int m, double *d_out, int32_t *d_status,
                        void *d_ws, size_t ws_bytes, void *stream);

/* ---- the early-exit front end inside the library ------------------------------------------------------------
 * Pop ranks of a FULL field (what fmb_truncate* and fmb_bi_join consume), incl. the reference's LIFO order among
 * exactly equal values (FastMarching.py:65-67,76-78 / FastMarching3D.py:77-95 bisect_left + insert): a stable radix
 * sort of T, the tie groups, and the ordered sweep of fmb_tie_order* -- one call, caller-supplied workspace of
 * fmb_workspace_bytes_pop_ranks(cells) bytes, asynchronous, no host synchronisation (whether ties exist at all and
 * whether a tie group is too large for the sweep is decided on the device).  d_T / d_cost dense; seed_index = flat
 * index of the source.  3D: focus_index >= 0 restricts the question to the tie group of that cell (the early exit
 * only depends on the order inside the group of `start`), -1 = every group.
 * fmb_pop_ranks_status (synchronises): out4 = {ties exist, largest tie group, waits that hit the safety limit
 * (must be 0), a group exceeded 4096 cells (the ranks are then the plain stable order -- not the reference's on
 * that degenerate map)}. */
size_t fmb_workspace_bytes_pop_ranks(int64_t cells);
int fmb_pop_ranks2d_f64(const double *d_T, const double *d_cost, int rows, int cols, int32_t seed_index, int32_t transposed,
                        int32_t *d_rank, void *d_ws, size_t ws_bytes, void *stream);
int fmb_pop_ranks3d_f64(const double *d_T, const double *d_cost, int ny, int nx, int nz, int32_t seed_index, int32_t focus_index,
                        int32_t *d_rank, void *d_ws, size_t ws_bytes, void *stream);
int fmb_pop_ranks_status(const void *d_ws, void *stream, int32_t *out4);

/* biComputeTmap in ONE call (FastMarching.py:114-162): both full fields, both rank sets, the join, both partial
 * fields.  d_cost dense [rows][cols]; goal_xy / start_xy HOST int32[2]; transposed: see fmb_tie_order2d_f64.
 * d_TG / d_TS: out, the partial fields the reference returns; d_join: device int32[16], [0] = k (pops per front when
 * the loop breaks), [1] = flat index of nodeJoin, both INT32_MAX when the fronts never meet (reference: NameError);
 * [4..7] / [8..11] = fmb_pop_ranks_status of the G / S front, [12] / [13] = waits of the G / S replay that hit the
 * safety limit (must be 0), [14] = relaxations the replay of the G front ran in its sparse form (the dependency cone of
 * the narrow band; the dense form replays all 4 (k + 1)), [15] = bit 0 / 1: the G / S front fell back to the dense form,
 * bits 8..: rounds the G front's cone expansion took (diagnostics).
 * stream2: optional second stream for the S front (NULL = everything on `stream`); on return `stream` has joined it.
 * Asynchronous; fmb_finish(d_ws, ...) reports device-side failures of the solve. */
size_t fmb_workspace_bytes_bisolve2d(int rows, int cols);
int fmb_bisolve2d_f64(const double *d_cost, int rows, int cols, const int32_t *goal_xy, const int32_t *start_xy, int32_t transposed,
                      double *d_TG, double *d_TS, int32_t *d_join, void *d_ws, size_t ws_bytes, void *stream, void *stream2);
/* The same with the cost map still on the host: h_cost = dense [rows][cols] PAGE-LOCKED host memory, uploaded into d_cost
 * in bands of rows on stream2 (required, != stream) behind the two solves that consume it (see fmb_solve2d_h2d_f64). */
int fmb_bisolve2d_h2d_f64(const double *h_cost, double *d_cost, int rows, int cols, const int32_t *goal_xy, const int32_t *start_xy,
                          int32_t transposed, double *d_TG, double *d_TS, int32_t *d_join, void *d_ws, size_t ws_bytes, void *stream,
                          void *stream2);
/* Single front with the reference's early exit in ONE call: FastMarching.py:92-112 (as intended) / FastMarching3D.py:126-145.
 * The partial field after `start` is accepted; the full field when start is outside the array, unreached, or == goal
 * (closed before the loop, never popped).  3D: the fast solve decides where it can -- the field in the reference's own
 * arithmetic (fmb_solve3d_exact_f64; 4-8x dearer) differs from it by ~3e-12, so only a (near-)tie of T[start] with
 * another cell can make the accepted set depend on it; that second solve is launched behind a device flag and its
 * kernels return at once when no cell lies within 1e-9 of T[start].
 * d_info: device int32[16], [0] = k (INT32_MAX = no truncation), [4..7] rank status, [12] replay waits at the limit,
 * [14] (3D) = 1 when the exact solve ran. */
size_t fmb_workspace_bytes_until2d(int rows, int cols);
size_t fmb_workspace_bytes_until3d(int ny, int nx, int nz);
int fmb_solve2d_until_f64(const double *d_cost, int rows, int cols, const int32_t *goal_xy, const int32_t *start_xy, int32_t transposed,
                          double *d_T, int32_t *d_info, void *d_ws, size_t ws_bytes, void *stream);
int fmb_solve3d_until_f64(const double *d_cost, int ny, int nx, int nz, const int32_t *goal_xyz, const int32_t *start_xyz,
                          double *d_T, int32_t *d_info, void *d_ws, size_t ws_bytes, void *stream);

/* ---- batch entry for a native host ---------------------------------------------------
 * Replaces, for N queries at once, the call sequence of the reference's C++ host: MotionPlanning.cpp:31-54
 * runPyFunction (one query per call into the embedded interpreter) followed by :66-91 returnPyArrayDouble/Int
 * (raw PyArray_DATA pointers into module globals, Coupled_motion_planner.py:1402-1406, 1678-1693, whose references
 * are leaked).  HOST pointers in; the result is OWNED by the caller and released with fmb_plan2d_free(); all device
 * memory and the stream live only inside the call (`device` = CUDA device to use, -1 = the current one).
 * Synchronous, re-entrant (one call per host thread).
 *   h_cost    [rows][cost_pitch] fp64 cost map shared by every query (+inf = obstacle)
 *   h_goals / h_starts  int32 [nq][2] = [x, y]
 *   per query: full-field solve from the goal (FastMarching.py:92-112), path start -> goal (FastMarching.py:164-236,
 *   tau, max_steps <= 0 = round(15000 / tau)), waypoints scaled as the planner does, resolution * (cell + 1)
 *   (Coupled_motion_planner.py:1234; resolution = 1 and subtract 1 for cell units).
 * Result: query q owns rows offsets[q] .. offsets[q + 1] of `waypoints` ([row][2] = x, y); status[q] = FMB_TRACE_*. */
typedef struct fmb_plan2d_result {
    int32_t nq;
    int64_t *offsets;       /* [nq + 1] */
    double *waypoints;      /* [offsets[nq]][2] */
    int32_t *status;        /* [nq] */
    double solve_ms, trace_ms;   /* device time of the solves / tracers of all chunks (CUDA events) */
} fmb_plan2d_result;
int fmb_plan_batch2d_host(const double *h_cost, int64_t cost_pitch, int rows, int cols, const int32_t *h_goals,
                          const int32_t *h_starts, int nq, double tau, int max_steps, double resolution, int device,
                          fmb_plan2d_result **out);
void fmb_plan2d_free(fmb_plan2d_result *r);

#ifdef __cplusplus
}
#endif
#endif /* FM_B200_H */
