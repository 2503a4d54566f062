// fm_capi.cu -- extern "C" boundary of libfm_b200.so (declared in include/fm_b200.h).
//
// Host side only launches kernels; no host-side numerics, no CPU fallback.
#include "../../include/fm_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "eikonal2d.cuh"
#include "eikonal2d_cta.cuh"
#include "eikonal2d_sweep.cuh"
#include "eikonal2d_wsweep.cuh"
#include "eikonal3d.cuh"
#include "eikonal3d_sweep.cuh"
#include "trace2d.cuh"
#include "trace3d.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return FMB_E_CUDA;
}
#define CK(call, where) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, where); } while (0)

// CUDA events bracketing the most recent solve of every (host thread, device, stream): timing evidence for
// bench.py.  Events belong to the device that was current when they were created, so they are kept per
// device; one host thread may drive several streams, so they are also kept per stream (at most TM_SLOTS
// streams per device and thread, oldest evicted).  Timing can never fail a solve: every error is swallowed.
struct Timing { cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr; bool armed = false; unsigned long long age = 0; };
constexpr int TM_DEVS = 16, TM_SLOTS = 16;
thread_local Timing g_tm[TM_DEVS][TM_SLOTS];
thread_local unsigned long long g_tm_clock = 0;
Timing *timing_slot(cudaStream_t st, bool create) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TM_DEVS) { cudaGetLastError(); return nullptr; }
    Timing *row = g_tm[dev], *victim = &row[0];
    for (int i = 0; i < TM_SLOTS; ++i) {
        if (row[i].e0 && row[i].st == st) { row[i].age = ++g_tm_clock; return &row[i]; }
        if (row[i].age < victim->age) victim = &row[i];
    }
    if (!create) return nullptr;
    if (!victim->e0) {
        if (cudaEventCreate(&victim->e0) != cudaSuccess || cudaEventCreate(&victim->e1) != cudaSuccess ||
            cudaEventCreate(&victim->e2) != cudaSuccess) { victim->e0 = nullptr; cudaGetLastError(); return nullptr; }
    }
    victim->st = st; victim->armed = false; victim->age = ++g_tm_clock;
    return victim;
}
void timing_begin(cudaStream_t st) {
    Timing *t = timing_slot(st, true);
    if (t) { t->armed = false; if (cudaEventRecord(t->e0, st) != cudaSuccess) cudaGetLastError(); }
}
void timing_mid(cudaStream_t st) {
    Timing *t = timing_slot(st, false);
    if (t && cudaEventRecord(t->e1, st) != cudaSuccess) cudaGetLastError();
}
void timing_end(cudaStream_t st) {
    Timing *t = timing_slot(st, false);
    if (t) { if (cudaEventRecord(t->e2, st) == cudaSuccess) t->armed = true; else cudaGetLastError(); }
}

// ---- tunables: one process-wide options struct (fmb_get_options / fmb_set_options); the FMB_*
// environment variables only give its INITIAL values, read once
std::mutex g_opt_mu;
fmb_options g_opt;
bool g_opt_init = false;
int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
void opt_defaults_locked() {
    if (g_opt_init) return;
    memset(&g_opt, 0, sizeof(g_opt));
    g_opt.engine2d = env_int("FMB_ENGINE2D", 0);
    g_opt.cta_cells = env_int("FMB_CTA_CELLS", 0);
    g_opt.tile_w2d = env_int("FMB_TW2D", 32);
    g_opt.tile_z3d = env_int("FMB_TZ3D", 16);
    g_opt.best_first = env_int("FMB_BEST_FIRST", -1);
    g_opt.windowed = env_int("FMB_WINDOWED", -1);
    g_opt.window = env_int("FMB_WINDOW", -1);
    g_opt.worker_div = env_int("FMB_WORKER_DIV", 0);
    g_opt.max_blocks = env_int("FMB_MAX_BLOCKS", 0);
    g_opt.watchdog_ms = env_int("FMB_WATCHDOG_MS", 20000);
    g_opt.step_cap = env_int("FMB_STEP_CAP", 1 << 20);
    g_opt.engine3d = env_int("FMB_ENGINE3D", 0);
    g_opt.level_div = env_int("FMB_LEVEL_DIV", 0);
    g_opt.win_running = env_int("FMB_WIN_RUNNING", -1);
    g_opt.check_passes = env_int("FMB_CHECK_PASSES", 0);
    g_opt.pipeline = env_int("FMB_PIPELINE", -1);
    g_opt.precheck = env_int("FMB_PRECHECK", -1);
    g_opt.causal_slack = env_int("FMB_CAUSAL_SLACK", 0);
    g_opt.tma = env_int("FMB_TMA", -1);
    g_opt.ring2 = env_int("FMB_RING2", 0);
    g_opt.variant = env_int("FMB_VARIANT", 0);
    g_opt.concurrent_solves = env_int("FMB_CONCURRENT_SOLVES", 0);
    g_opt.replay_sparse = env_int("FMB_REPLAY_SPARSE", -1);
    g_opt_init = true;
}
fmb_options opt() {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    opt_defaults_locked();
    return g_opt;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

unsigned pow2_at_least(long long v) {
    unsigned p = 1024;
    while ((long long)p < v && p < (1u << 30)) p <<= 1;
    return p;
}

// workspace layout: [QueueCtl | pad to 256] [tile_state: ntiles ints] [ring: slots ints]
struct WsLayout {
    size_t ctl_off, state_off, ring_off, prio_off, win_off, level_off, runprio_off, total;
    unsigned ring_slots;
};
WsLayout ws_layout(long long ntiles) {
    WsLayout L;
    L.ctl_off = 0;
    L.state_off = 256;
    size_t st = ((size_t)ntiles * sizeof(int) + 255) & ~(size_t)255;
    L.ring_off = L.state_off + st;
    L.ring_slots = pow2_at_least(ntiles);
    L.prio_off = (L.ring_off + (size_t)L.ring_slots * sizeof(int) + 255) & ~(size_t)255;
    // windowed order: [inv_delta double | hint int | pad] [lev_count: WIN_LEVELS ints] [tile_level: ntiles ints]
    L.win_off = (L.prio_off + (size_t)ntiles * sizeof(unsigned long long) + 255) & ~(size_t)255;
    L.level_off = L.win_off + 256 + (size_t)fmb::WIN_LEVELS * sizeof(int);
    L.runprio_off = (L.level_off + (size_t)ntiles * sizeof(int) + 255) & ~(size_t)255;
    L.total = L.runprio_off + (size_t)ntiles * sizeof(unsigned long long);
    return L;
}

constexpr int MIN_TW = 16;       // workspace is sized for the narrowest tile so the tile width can be tuned at run time
constexpr int WARPS = 4;         // worker warps per CTA of the persistent solver

long long tiles2d(int rows, int cols, int tw) {
    return (long long)((cols + tw - 1) / tw) * ((rows + fmb::TILE_H - 1) / fmb::TILE_H);
}

template <typename real, int TW>
void launch_init2d(const fmb::Problem2D<real> &P, const WsLayout &L, cudaStream_t st, int resume_activate, long long fill_blocks) {
    if (resume_activate < 0) {
        fmb::init_fill2d_kernel<real><<<(unsigned)fill_blocks, 256, 0, st>>>(P, (int)L.ring_slots);
    } else {
        fmb::init_resume2d_kernel<real><<<64, 256, 0, st>>>(P, (int)L.ring_slots);
        if (resume_activate & 7)
            fmb::activate_rows2d_kernel<real><<<(P.ntx * P.nty + 127) / 128, 128, 0, st>>>(P, resume_activate);
    }
    fmb::init_seed2d_kernel<real, TW><<<(P.nq + 127) / 128, 128, 0, st>>>(P);     // out-of-range seed = no seed
}

template <typename real, int TW, bool BEST, bool CG = false>
int launch_solve2d(fmb::Problem2D<real> P, const WsLayout &L, cudaStream_t st, int resume_activate = -1) {
    using TL = fmb::Tile2D<real, TW>;
    const size_t smem = (CG ? sizeof(real) * TL::T_ELEMS : TL::WARP_BYTES) * WARPS;
    auto kern = fmb::solve2d_kernel<real, TW, WARPS, BEST, CG>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(solve2d)");
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem), "occupancy(solve2d)");
    if (per_sm < 1) return fail(FMB_E_CUDA, "solve2d kernel does not fit on an SM%s");
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    long long blocks = (long long)per_sm * sm_count();
    // One map: more workers than ~half the tiles only add speculative re-visits (measured: 400^2
    // and 90x90x28 run faster with fewer warps); batches keep one worker per tile up to the machine.
    // Windowed order: the solve is bound by the chain of tile visits, not by throughput -- one warp per
    // 28 tiles finishes as fast as the whole machine with a sixth of the work (4096^2: 22.4 ms, 13.6
    // evaluations per cell instead of 90), which leaves the other SM slots to concurrent solves.
    const fmb_options O = opt();
    const int div = O.worker_div > 0 ? O.worker_div : (P.windowed ? 28 : (P.nq == 1 ? 2 : 1));
    const long long need = (ntiles + (long long)WARPS * div - 1) / ((long long)WARPS * div);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (O.max_blocks > 0 && blocks > O.max_blocks) blocks = O.max_blocks;

    const long long cells = (long long)P.rows * P.cols * P.nq;
    long long fill_blocks = (cells + 256 * 8 - 1) / (256 * 8);
    if (fill_blocks > (long long)sm_count() * 16) fill_blocks = (long long)sm_count() * 16;
    if (fill_blocks < 1) fill_blocks = 1;
    cudaGetLastError();          // a stale error of the caller must not be reported as ours
    timing_begin(st);
    launch_init2d<real, TW>(P, L, st, resume_activate, fill_blocks);
    timing_mid(st);
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(P);
    cudaError_t le = cudaGetLastError();
    timing_end(st);
    CK(le, "launch solve2d");
    return FMB_OK;
}

// CTA-per-tile engine (eikonal2d_cta.cuh): 1024 / R threads per tile visit, 32 x 32 tiles
template <typename real, int R, bool BEST>
int launch_solve2d_cta(fmb::Problem2D<real> P, const WsLayout &L, cudaStream_t st, int resume_activate = -1) {
    using TL = fmb::CtaTile2D<real>;
    constexpr int NT = 1024 / R;
    const size_t smem = TL::BYTES;
    auto kern = fmb::solve2d_cta_kernel<real, R, BEST>;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem), "occupancy(solve2d_cta)");
    if (per_sm < 1) return fail(FMB_E_CUDA, "solve2d_cta kernel does not fit on an SM%s");
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    long long blocks = (long long)per_sm * sm_count();
    const fmb_options O = opt();
    const int div = O.worker_div > 0 ? O.worker_div : (P.nq == 1 ? 2 : 1);
    const long long need = (ntiles + div - 1) / div;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (O.max_blocks > 0 && blocks > O.max_blocks) blocks = O.max_blocks;
    const long long cells = (long long)P.rows * P.cols * P.nq;
    long long fill_blocks = (cells + 256 * 8 - 1) / (256 * 8);
    if (fill_blocks > (long long)sm_count() * 16) fill_blocks = (long long)sm_count() * 16;
    if (fill_blocks < 1) fill_blocks = 1;
    cudaGetLastError();
    timing_begin(st);
    launch_init2d<real, 32>(P, L, st, resume_activate, fill_blocks);
    timing_mid(st);
    kern<<<(unsigned)blocks, NT, smem, st>>>(P);
    cudaError_t le = cudaGetLastError();
    timing_end(st);
    CK(le, "launch solve2d_cta");
    return FMB_OK;
}

// TMA tensor maps of a solve's T and cost arrays (3D: x, y, query), built per call on the host (cuTensorMapEncodeTiled is
// reached through cudaGetDriverEntryPoint: the library does not link libcuda).  Returns false when the arrays cannot be
// described (unaligned base / pitch, driver without the entry point): the kernel then stages with cp.async.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}
bool make_tmap3d(CUtensorMap *m, const void *base, int rows, int cols, long long pitch, long long qstride, int nq, int box_w, int box_h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (((size_t)base % 16) != 0 || (pitch % 2) != 0 || (qstride % 2) != 0 || cols < box_w || rows < box_h) return false;
    const bool per_q = qstride > 0 && nq > 1;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(per_q ? nq : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)(per_q ? qstride : (long long)rows * pitch) * 8};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Sweep engine (eikonal2d_sweep.cuh): 4 warps per tile visit, one diagonal-wavefront sweep each
template <typename real, bool BEST>
int launch_solve2d_sweep(fmb::Problem2D<real> P, const WsLayout &L, cudaStream_t st, int resume_activate = -1) {
    const size_t smem = fmb::Sweep2DSmem::bytes<real>();
    auto kern = fmb::solve2d_sweep_kernel<real, BEST>;
    fmb::TmaMaps2D tm;
    memset(&tm, 0, sizeof(tm));
    const fmb_options O0 = opt();
    int use_tma = 0;
    if (sizeof(real) == 8 && O0.tma != 0)
        use_tma = make_tmap3d(&tm.T, P.T, P.rows, P.cols, P.T_pitch, P.T_qstride, P.nq, fmb::Sweep2DSmem::PT, fmb::TILE_H + 2) &&
                  make_tmap3d(&tm.C, P.cost, P.rows, P.cols, P.cost_pitch, P.cost_qstride, P.nq, fmb::Sweep2DSmem::PC, fmb::TILE_H);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem), "occupancy(solve2d_sweep)");
    if (per_sm < 1) return fail(FMB_E_CUDA, "solve2d_sweep kernel does not fit on an SM%s");
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    long long blocks = (long long)per_sm * sm_count();
    const fmb_options O = opt();
    const int div = O.worker_div > 0 ? O.worker_div : (P.nq == 1 ? 2 : 1);
    const long long need = (ntiles + div - 1) / div;
    if (blocks > need) blocks = need;
    // One map in causal order is a chain of dependent visits: two CTAs per SM per query run it as fast as the whole
    // machine (measured 4096^2: 10.3 ms with 296 CTAs, 10.1 ms with 1036) and leave the other slots to concurrent
    // solves on other streams -- a persistent grid that fills every slot would serialise them.
    if (!P.best_first && P.windowed == 2 && O.max_blocks <= 0) {
        long long lean = (long long)2 * sm_count() * P.nq;
        // n solves in flight on n streams (fmb_options.concurrent_solves): 1/n of the resident slots each, so that no
        // solve waits for another one's persistent CTAs to retire (4096^2, 444 slots: 3 x 296 CTAs 5.2 ms per solve,
        // 3 x 148 3.9 ms, 4 x 111 3.5 ms; one solve alone: 7.3 ms with 296 CTAs, 7.6 ms with 148)
        if (O.concurrent_solves > 1) {
            const long long share = (long long)per_sm * sm_count() / O.concurrent_solves;
            if (share >= 1 && share < lean) lean = share;
        }
        if (blocks > lean) blocks = lean;
    }
    if (blocks < 1) blocks = 1;
    if (O.max_blocks > 0 && blocks > O.max_blocks) blocks = O.max_blocks;      // (an explicit cap also lifts the lean default)
    const long long cells = (long long)P.rows * P.cols * P.nq;
    long long fill_blocks = (cells + 256 * 8 - 1) / (256 * 8);
    if (fill_blocks > (long long)sm_count() * 16) fill_blocks = (long long)sm_count() * 16;
    if (fill_blocks < 1) fill_blocks = 1;
    cudaGetLastError();
    timing_begin(st);
    launch_init2d<real, 32>(P, L, st, resume_activate, fill_blocks);
    timing_mid(st);
    kern<<<(unsigned)blocks, 128, smem, st>>>(P, tm, use_tma);
    cudaError_t le = cudaGetLastError();
    timing_end(st);
    CK(le, "launch solve2d_sweep");
    return FMB_OK;
}

// Warp-per-tile sweep engine for batches (eikonal2d_wsweep.cuh): 4 independent warps per CTA
template <typename real, bool BEST, bool STAGE_C>
int launch_solve2d_wsweep(fmb::Problem2D<real> P, const WsLayout &L, cudaStream_t st) {
    using TL = fmb::Tile2D<real, 32>;
    const size_t smem = sizeof(real) * (TL::T_ELEMS + (STAGE_C ? TL::C_ELEMS : 0)) * WARPS;
    auto kern = fmb::solve2d_wsweep_kernel<real, BEST, STAGE_C>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(solve2d_wsweep)");
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem), "occupancy(solve2d_wsweep)");
    if (per_sm < 1) return fail(FMB_E_CUDA, "solve2d_wsweep kernel does not fit on an SM%s");
    const long long ntiles = (long long)P.nq * P.ntx * P.nty;
    long long blocks = (long long)per_sm * sm_count();
    const fmb_options O = opt();
    const int div = O.worker_div > 0 ? O.worker_div : (P.nq == 1 ? 2 : 1);
    const long long need = (ntiles + (long long)WARPS * div - 1) / ((long long)WARPS * div);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (O.max_blocks > 0 && blocks > O.max_blocks) blocks = O.max_blocks;
    const long long cells = (long long)P.rows * P.cols * P.nq;
    long long fill_blocks = (cells + 256 * 8 - 1) / (256 * 8);
    if (fill_blocks > (long long)sm_count() * 16) fill_blocks = (long long)sm_count() * 16;
    if (fill_blocks < 1) fill_blocks = 1;
    cudaGetLastError();
    timing_begin(st);
    launch_init2d<real, 32>(P, L, st, -1, fill_blocks);
    timing_mid(st);
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(P);
    cudaError_t le = cudaGetLastError();
    timing_end(st);
    CK(le, "launch solve2d_wsweep");
    return FMB_OK;
}

template <typename real>
int solve2d(const real *d_cost, int64_t cost_pitch, int64_t cost_qstride, real *d_T, int64_t T_pitch,
            int64_t T_qstride, int rows, int cols, int nq, const int32_t *d_seeds, void *d_ws, size_t ws_bytes,
            void *stream, int resume_activate = -1, int arm_rows = 0, const int *band_ready = nullptr, int band_shift = 0) {
    if (!d_cost || !d_T || !d_seeds || !d_ws) return fail(FMB_E_INVALID, "null pointer argument%s");
    if (rows < 1 || cols < 1 || nq < 1) return fail(FMB_E_INVALID, "rows, cols and nq must be positive%s");
    if (cost_pitch < cols || T_pitch < cols) return fail(FMB_E_INVALID, "pitch smaller than cols%s");
    const fmb_options O = opt();
    // Work order first (it decides the engine): best-first per query pays off when many queries share the GPU and a
    // query's tile table is small enough to scan per claim (batched planning: a throughput regime).
    const long long tiles_per_q32 = tiles2d(rows, cols, 32);
    const int best_first = resume_activate >= 0 ? 0 : (O.best_first >= 0 ? O.best_first : ((nq >= 8 && tiles_per_q32 <= 1024) ? 1 : 0));
    // Engine: the four-warp sweep visit (eikonal2d_sweep.cuh) where the solve is a chain of dependent visits (one map,
    // a few maps); the warp-per-tile visit (eikonal2d.cuh) for batches, where throughput per warp counts (measured
    // 4096 x 512^2: 64 ms against 142 ms), without the shared-memory cost tile (engine 6: 20 resident warps per SM
    // instead of 12, 55 ms).
    int engine = O.engine2d > 0 ? O.engine2d : (best_first ? 6 : 3);
    // engine 6 = engine 1 with the cost read from global memory instead of a shared-memory tile (best-first batches)
    const bool cost_global = engine == 6 && best_first;
    if (engine == 6) engine = 1;
    int tw = (engine >= 2 || cost_global) ? 32 : O.tile_w2d;       // engines 2-6 use 32 x 32 tiles
    if (tw != 16 && tw != 32) return fail(FMB_E_INVALID, "tile_w2d must be 16 or 32%s");
    const long long ntiles = tiles2d(rows, cols, tw) * nq;
    if (ntiles >= (1LL << 30)) return fail(FMB_E_INVALID, "too many tiles for one launch%s");
    if (ws_bytes < fmb_workspace_bytes_2d(rows, cols, nq)) return fail(FMB_E_WORKSPACE, "workspace too small%s");
    WsLayout L = ws_layout(ntiles);
    char *ws = (char *)d_ws;
    fmb::Problem2D<real> P;
    P.cost = d_cost; P.cost_pitch = cost_pitch; P.cost_qstride = cost_qstride;
    P.T = d_T; P.T_pitch = T_pitch; P.T_qstride = T_qstride;
    P.rows = rows; P.cols = cols; P.nq = nq;
    P.ntx = (cols + tw - 1) / tw; P.nty = (rows + fmb::TILE_H - 1) / fmb::TILE_H;
    P.seeds = d_seeds;
    P.tile_state = (int *)(ws + L.state_off);
    P.q.ctl = (fmb::QueueCtl *)(ws + L.ctl_off);
    P.q.ring = (int *)(ws + L.ring_off);
    P.q.ring_mask = L.ring_slots - 1;
    P.q.watchdog_cycles = (long long)(O.watchdog_ms > 0 ? O.watchdog_ms : 20000) * 2000000LL;   // ~2 GHz
    P.step_cap = O.step_cap > 0 ? O.step_cap : 1 << 20;
    P.tile_prio = (unsigned long long *)(ws + L.prio_off);
    P.best_first = best_first;
    P.arm_rows = arm_rows;
    // windowed FIFO (deferral of tiles far ahead of the lowest queued level): one large map only
    // one map: the sweep engine runs the local causal order at every size; the older engines keep the windowed FIFO
    // for maps of >= 16384 tiles
    P.windowed = (!P.best_first && resume_activate < 0)
                     ? (O.windowed >= 0 ? O.windowed : (engine == 3 ? 2 : ((nq == 1 && ntiles >= 16384) ? 1 : 0))) : 0;
    if (P.windowed == 2 && engine != 3 && engine < 4) P.windowed = 1;       // the local causal order exists in the sweep engines only
    if (engine >= 4 && P.windowed == 1) P.windowed = 0;       // (the warp-sweep engine has FIFO, causal and best-first orders)
    if (P.windowed == 1 && nq != 1) P.windowed = 0;           // the level window is kept per launch, not per query
    P.win_window = O.window >= 0 ? O.window : 2;
    P.check_passes = O.check_passes > 0 ? O.check_passes : (engine == 3 ? 2 : 4);
    P.precheck = O.precheck >= 0 ? O.precheck : 0;
    P.pipeline = O.pipeline >= 0 ? O.pipeline : 0;      // early publish: measured slower (more rounds per visit)
    P.win_div = engine == 3 ? (O.level_div > 0 ? O.level_div : 1) : 1;
    P.win_running = engine == 3 ? (O.win_running >= 0 ? O.win_running : 1) : 0;
    P.win_inv_delta = (double *)(ws + L.win_off);
    P.win_hint = (int *)(ws + L.win_off + 8);
    P.lev_count = (int *)(ws + L.win_off + 256);
    P.tile_level = (int *)(ws + L.level_off);
    P.run_prio = (unsigned long long *)(ws + L.runprio_off);
    P.slack = (double *)(ws + L.win_off + 16);
    // measured 4096^2 (planner-like map): 10.7 ms without the second-ring rule, 9.3 with it (visits per tile 3.8 -> 2.1);
    // 7.0 ms with the straight-line sweep step as well (510 -> 376 cycles per step) and two check passes instead of four
    P.hop_frac = O.ring2 > 0 ? 0.01 * O.ring2 : (O.ring2 < 0 ? 0.0 : 2.0);
    P.variant = O.variant > 0 ? O.variant : (O.variant < 0 ? 0 : 3);
    P.band_ready = band_ready; P.band_shift = band_shift;
    if (band_ready && !(engine == 3 && !P.best_first && (nq == 1 || cost_qstride == 0) && resume_activate < 0))
        return fail(FMB_E_INVALID, "a solve on a cost map that is still arriving needs the sweep engine on one (shared) map%s");
    P.slack_frac = O.causal_slack > 0 ? 0.01 * O.causal_slack : 0.0;      // measured 4096^2: 10.4 / 10.6 / 11.9 / 15.9 ms at 0 / 25 / 50 / 100 %
    cudaStream_t st = (cudaStream_t)stream;
    if (engine >= 4 && resume_activate < 0) {
        if (engine == 4) return P.best_first ? launch_solve2d_wsweep<real, true, true>(P, L, st) : launch_solve2d_wsweep<real, false, true>(P, L, st);
        return P.best_first ? launch_solve2d_wsweep<real, true, false>(P, L, st) : launch_solve2d_wsweep<real, false, false>(P, L, st);
    }
    if (engine == 3 || engine >= 4) {
        if (P.best_first) return launch_solve2d_sweep<real, true>(P, L, st);
        return launch_solve2d_sweep<real, false>(P, L, st, resume_activate);
    }
    if (engine == 2) {
        const int R = O.cta_cells > 0 ? O.cta_cells : 2;
        if (R != 1 && R != 2 && R != 4) return fail(FMB_E_INVALID, "cta_cells must be 1, 2 or 4%s");
        if (P.best_first) {
            if (R == 1) return launch_solve2d_cta<real, 1, true>(P, L, st);
            if (R == 2) return launch_solve2d_cta<real, 2, true>(P, L, st);
            return launch_solve2d_cta<real, 4, true>(P, L, st);
        }
        if (R == 1) return launch_solve2d_cta<real, 1, false>(P, L, st, resume_activate);
        if (R == 2) return launch_solve2d_cta<real, 2, false>(P, L, st, resume_activate);
        return launch_solve2d_cta<real, 4, false>(P, L, st, resume_activate);
    }
    if (P.best_first) {
        if (tw == 16) return launch_solve2d<real, 16, true>(P, L, st);
        if (cost_global) return launch_solve2d<real, 32, true, true>(P, L, st);
        return launch_solve2d<real, 32, true>(P, L, st);
    }
    if (tw == 16) return launch_solve2d<real, 16, false>(P, L, st, resume_activate);
    return launch_solve2d<real, 32, false>(P, L, st, resume_activate);
}

}  // namespace

namespace {
// bitwise comparison of the branch-free square root with sqrt.rn.f64 (self-test entry, tests/test_gpu_parity.py)
__global__ void sqrt_check_kernel(const double *x, long long n, unsigned long long *bad) {
    unsigned long long mine = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (!fmb::sqrt_fast_ok(v)) continue;
        if (__double_as_longlong(fmb::sqrt_rn_fast(v)) != __double_as_longlong(__dsqrt_rn(v))) ++mine;
    }
    if (mine) atomicAdd(bad, mine);
}
// bitwise comparison of the branch-free division with div.rn.f64; d_bad[1] counts the pairs the fast form accepts
__global__ void div_check_kernel(const double *a, const double *b, long long n, unsigned long long *bad) {
    unsigned long long mine = 0, accepted = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bool ok = true;
        const double q = fmb::ddiv_rn_fast(a[i], b[i], ok);
        if (!ok) continue;
        ++accepted;
        if (__double_as_longlong(q) != __double_as_longlong(__ddiv_rn(a[i], b[i]))) ++mine;
    }
    if (mine) atomicAdd(bad, mine);
    if (accepted) atomicAdd(bad + 1, accepted);
}
}  // namespace

extern "C" {

int fmb_debug_div_check(const double *d_a, const double *d_b, int64_t n, uint64_t *d_bad2, void *stream) {
    if (!d_a || !d_b || !d_bad2 || n < 1) return fail(FMB_E_INVALID, "bad argument%s");
    div_check_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, (unsigned long long *)d_bad2);
    CK(cudaGetLastError(), "launch div_check");
    return FMB_OK;
}

int fmb_debug_sqrt_check(const double *d_x, int64_t n, uint64_t *d_bad, void *stream) {
    if (!d_x || !d_bad || n < 1) return fail(FMB_E_INVALID, "bad argument%s");
    sqrt_check_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(d_x, n, (unsigned long long *)d_bad);
    CK(cudaGetLastError(), "launch sqrt_check");
    return FMB_OK;
}

int fmb_version(void) { return 100; }
const char *fmb_last_error(void) { return g_err; }
int fmb_sm_count(void) { return sm_count(); }
void fmb_get_options(fmb_options *out) { if (out) *out = opt(); }
int fmb_set_options(const fmb_options *in) {
    if (!in) return fail(FMB_E_INVALID, "null options%s");
    if (in->cta_cells != 0 && in->cta_cells != 1 && in->cta_cells != 2 && in->cta_cells != 4) return fail(FMB_E_INVALID, "cta_cells must be 0, 1, 2 or 4%s");
    if (in->tile_w2d != 16 && in->tile_w2d != 32) return fail(FMB_E_INVALID, "tile_w2d must be 16 or 32%s");
    if (in->tile_z3d != 16 && in->tile_z3d != 32) return fail(FMB_E_INVALID, "tile_z3d must be 16 or 32%s");
    std::lock_guard<std::mutex> lk(g_opt_mu);
    opt_defaults_locked();
    g_opt = *in;
    return FMB_OK;
}

size_t fmb_workspace_bytes_2d(int rows, int cols, int nq) {
    if (rows < 1 || cols < 1 || nq < 1) return 0;
    return ws_layout(tiles2d(rows, cols, MIN_TW) * nq).total;
}

int fmb_solve2d_f64(const double *d_cost, int64_t cost_pitch, int64_t cost_qstride, double *d_T, int64_t T_pitch,
                    int64_t T_qstride, int rows, int cols, int nq, const int32_t *d_seeds, void *d_ws,
                    size_t ws_bytes, void *stream) {
    return solve2d<double>(d_cost, cost_pitch, cost_qstride, d_T, T_pitch, T_qstride, rows, cols, nq, d_seeds, d_ws,
                           ws_bytes, stream);
}
int fmb_solve2d_f32(const float *d_cost, int64_t cost_pitch, int64_t cost_qstride, float *d_T, int64_t T_pitch,
                    int64_t T_qstride, int rows, int cols, int nq, const int32_t *d_seeds, void *d_ws,
                    size_t ws_bytes, void *stream) {
    return solve2d<float>(d_cost, cost_pitch, cost_qstride, d_T, T_pitch, T_qstride, rows, cols, nq, d_seeds, d_ws,
                          ws_bytes, stream);
}

// ---- upload of the cost map overlapped with the solve --------------------------------------------------------------
namespace {
int band_shift_for(int rows) {              // ~16 bands, a power of two of at least 32 rows (tile rows never straddle a band)
    int sh = 5;
    while ((rows >> sh) > 16) ++sh;
    return sh;
}
int *pinned_one() {
    static int *one = nullptr;
    if (!one) { if (cudaHostAlloc((void **)&one, sizeof(int), cudaHostAllocDefault) != cudaSuccess) { one = nullptr; cudaGetLastError(); } else *one = 1; }
    return one;
}
struct BandEvents {                              // released on every return path (destroying a recorded event is deferred by the runtime)
    cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
    ~BandEvents() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
};
size_t band_flag_bytes(int rows) { return 4 * (size_t)((rows >> 5) + 2) + 256; }
// Queues the upload of a page-locked host map in bands of rows on the copy stream `cs`, each band followed by the write of
// its device flag: first the bands that hold the seed rows (event e[1] marks their arrival: the seed kernel reads the cost
// there), then the others by their distance from the nearest seed; e[2] marks the end.  The flags are cleared on `st` first
// (e[0]).  Every copy is queued BEFORE the caller launches the solve: where streams do not overlap (a serialising
// profiler, CUDA_LAUNCH_BLOCKING) the kernel then finds all flags set instead of waiting for work queued behind it.
int enqueue_band_upload(const double *h_cost, double *d_cost, int rows, int cols, const int *seed_rows, int nseeds, int *flags,
                        cudaStream_t st, cudaStream_t cs, BandEvents &ev, int *shift_out) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, h_cost) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return fail(FMB_E_INVALID, "h_cost must be page-locked host memory (cudaHostAlloc / cudaHostRegister)%s");
    }
    int *one = pinned_one();
    if (!one) return fail(FMB_E_CUDA, "cudaHostAlloc(flag source) failed%s");
    const int sh = band_shift_for(rows), band_rows = 1 << sh, nb = (rows + band_rows - 1) >> sh;
    *shift_out = sh;
    for (cudaEvent_t &x : ev.e) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming), "cudaEventCreate");
    CK(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)nb, st), "cudaMemsetAsync(band flags)");
    CK(cudaEventRecord(ev.e[0], st), "cudaEventRecord");
    CK(cudaStreamWaitEvent(cs, ev.e[0], 0), "cudaStreamWaitEvent");
    std::vector<char> sent((size_t)nb, 0);
    auto send = [&](int b) -> int {
        if (b < 0 || b >= nb || sent[b]) return FMB_OK;
        sent[b] = 1;
        const int r0 = b << sh, nr = (r0 + band_rows <= rows) ? band_rows : rows - r0;
        CK(cudaMemcpyAsync(d_cost + (size_t)r0 * cols, h_cost + (size_t)r0 * cols, sizeof(double) * (size_t)nr * cols,
                           cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(cost band)");
        CK(cudaMemcpyAsync(flags + b, one, sizeof(int), cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(band flag)");
        return FMB_OK;
    };
    for (int k = 0; k < nseeds; ++k) { const int rc = send(seed_rows[k] >> sh); if (rc) return rc; }
    CK(cudaEventRecord(ev.e[1], cs), "cudaEventRecord");
    for (int d = 0; d < nb; ++d)                 // rings of bands around every seed: b + 1, b - 1, b + 2, ... (a front's next
        for (int k = 0; k < nseeds; ++k) {       // rows arrive before it needs them, whichever seed it started from)
            const int b = seed_rows[k] >> sh;
            int rc = send(b + d); if (rc) return rc;
            rc = send(b - d); if (rc) return rc;
        }
    CK(cudaEventRecord(ev.e[2], cs), "cudaEventRecord");
    return FMB_OK;
}
}  // namespace
size_t fmb_workspace_bytes_2d_h2d(int rows, int cols) {
    if (rows < 1 || cols < 1) return 0;
    return ((fmb_workspace_bytes_2d(rows, cols, 1) + 255) & ~(size_t)255) + 256 + band_flag_bytes(rows);
}
int fmb_solve2d_h2d_f64(const double *h_cost, double *d_cost, int rows, int cols, const int32_t *goal_xy, double *d_T,
                        void *d_ws, size_t ws_bytes, void *stream, void *copy_stream) {
    if (!h_cost || !d_cost || !goal_xy || !d_T || !d_ws || !copy_stream) return fail(FMB_E_INVALID, "null pointer argument%s");
    if (rows < 1 || cols < 1) return fail(FMB_E_INVALID, "bad shape%s");
    if (goal_xy[0] < 0 || goal_xy[0] >= cols || goal_xy[1] < 0 || goal_xy[1] >= rows) return fail(FMB_E_INVALID, "goal outside the map%s");
    if (ws_bytes < fmb_workspace_bytes_2d_h2d(rows, cols)) return fail(FMB_E_WORKSPACE, "workspace too small%s");
    cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
    char *ws = (char *)d_ws;
    const size_t solve_bytes = fmb_workspace_bytes_2d(rows, cols, 1);
    size_t o = (solve_bytes + 255) & ~(size_t)255;
    int32_t *seeds = (int32_t *)(ws + o); o += 256;
    int *flags = (int *)(ws + o);
    BandEvents ev;
    int sh = 0;
    const int seed_row = goal_xy[1];
    int rc = enqueue_band_upload(h_cost, d_cost, rows, cols, &seed_row, 1, flags, st, cs, ev, &sh);
    if (rc) return rc;
    CK(cudaMemcpyAsync(seeds, goal_xy, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(seeds)");
    CK(cudaStreamWaitEvent(st, ev.e[1], 0), "cudaStreamWaitEvent");        // the seed kernel reads the cost at the goal
    rc = solve2d<double>(d_cost, cols, 0, d_T, cols, (int64_t)rows * cols, rows, cols, 1, seeds, d_ws, solve_bytes, stream, -1, 0,
                         flags, sh);
    CK(cudaStreamWaitEvent(st, ev.e[2], 0), "cudaStreamWaitEvent");        // d_cost is complete for whatever follows on `stream`
    return rc;
}

int fmb_resolve2d_f64(const double *d_cost, int64_t cost_pitch, double *d_T, int64_t T_pitch, int rows, int cols,
                      const int32_t *d_seed, int activate, int halo_rows, void *d_ws, size_t ws_bytes, void *stream) {
    if (activate < 0 || activate > 7 || halo_rows < 0 || halo_rows > 3) return fail(FMB_E_INVALID, "bad activate / halo_rows%s");
    return solve2d<double>(d_cost, cost_pitch, 0, d_T, T_pitch, 0, rows, cols, 1, d_seed, d_ws, ws_bytes, stream,
                           activate, halo_rows);
}

int fmb_finish(void *d_ws, size_t ws_bytes, void *stream, fmb_stats *stats) {
    if (!d_ws || ws_bytes < 256) return fail(FMB_E_INVALID, "bad workspace%s");
    CK(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
    fmb::QueueCtl h;
    CK(cudaMemcpy(&h, d_ws, sizeof(h), cudaMemcpyDeviceToHost), "cudaMemcpy(ctl)");
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->tile_visits = h.visits; stats->steps = h.steps; stats->evals = h.evals;
        stats->pushes = h.pushes; stats->cells_written = h.cells_written;
        stats->reserved[0] = h.pad[0];     /* deferrals of the windowed order */
        stats->cyc_check = h.pad[1]; stats->noop_visits = h.noop_visits; stats->rounds = h.rounds; stats->continuations = h.continuations;
        stats->cyc_wait = h.cyc_wait; stats->cyc_load = h.cyc_load; stats->cyc_relax = h.cyc_relax; stats->cyc_store = h.cyc_store;
        Timing *tm = timing_slot((cudaStream_t)stream, false);
        if (tm && tm->armed) {
            float a = 0.f, b = 0.f;
            if (cudaEventElapsedTime(&a, tm->e0, tm->e1) == cudaSuccess && cudaEventElapsedTime(&b, tm->e1, tm->e2) == cudaSuccess) {
                stats->init_kernel_ms = a; stats->solve_kernel_ms = b;
            } else cudaGetLastError();
        }
    }
    // a failure is reported ONCE: by this call for the solve it finishes, or -- carried over by the init kernels of later
    // solves (QueueCtl::sticky) -- for an earlier solve that was queued on this workspace and never finished
    if (!h.abort && h.sticky_magic == fmb::STICKY_MAGIC && h.sticky) h.abort = h.sticky;
    if (h.abort) {
        const int zero = 0;
        CK(cudaMemcpy((char *)d_ws + offsetof(fmb::QueueCtl, abort), &zero, sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(abort)");
        CK(cudaMemcpy((char *)d_ws + offsetof(fmb::QueueCtl, sticky), &zero, sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(sticky)");
    }
    if (h.abort == fmb::DEV_WATCHDOG) return fail(FMB_E_WATCHDOG, "device watchdog fired: a queue wait exceeded FMB_WATCHDOG_MS%s");
    if (h.abort == fmb::DEV_COSTRANGE) return fail(FMB_E_INVALID, "a finite cost lies outside the supported range [1e-140, 1e140] (fp32: [1e-15, 1e15])%s");
    if (h.abort == fmb::DEV_STEPCAP) return fail(FMB_E_STEPCAP, "in-tile iteration cap (FMB_STEP_CAP) exceeded%s");
    if (h.abort) return fail(FMB_E_CUDA, "unknown device-side failure%s");
    if (h.pending != 0) return fail(FMB_E_CUDA, "solver left with pending tiles%s");
    return FMB_OK;
}

namespace {
int trace2d_impl(const double *d_T, int64_t T_pitch, int64_t T_qstride, int rows, int cols, int npaths,
                 const int32_t *d_field_of_path, const double *d_init, const double *d_end, double tau,
                 int max_steps, double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status, int32_t *d_blocks,
                 int32_t *d_nblocks, int cap_blocks, void *stream) {
    if (!d_T || !d_init || !d_end || !d_out || !d_count || !d_status) return fail(FMB_E_INVALID, "null pointer argument%s");
    if (rows < 2 || cols < 2 || npaths < 1 || max_steps < 0) return fail(FMB_E_INVALID, "bad shape%s");
    if (cap < (int64_t)max_steps + 2) return fail(FMB_E_INVALID, "cap must be >= max_steps + 2%s");
    fmb::TraceArgs2D<double> A;
    A.T = d_T; A.T_pitch = T_pitch; A.T_qstride = T_qstride; A.rows = rows; A.cols = cols; A.npaths = npaths;
    A.field_of_path = d_field_of_path; A.init = d_init; A.end = d_end; A.tau = tau; A.max_steps = max_steps;
    A.out = d_out; A.cap = cap; A.count = d_count; A.status = d_status;
    A.blocks = d_blocks; A.nblocks = d_nblocks; A.cap_blocks = cap_blocks;
    constexpr int TW_ = 4;
    fmb::trace2d_kernel<double, TW_><<<(npaths + TW_ - 1) / TW_, TW_ * 32, TW_ * fmb::TRACE2D_SMEM_PER_WARP, (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError(), "launch trace2d");
    return FMB_OK;
}
}  // namespace

int fmb_trace2d_f64(const double *d_T, int64_t T_pitch, int64_t T_qstride, int rows, int cols, int npaths,
                    const int32_t *d_field_of_path, const double *d_init, const double *d_end, double tau,
                    int max_steps, double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status, void *stream) {
    return trace2d_impl(d_T, T_pitch, T_qstride, rows, cols, npaths, d_field_of_path, d_init, d_end, tau, max_steps, d_out, cap,
                        d_count, d_status, nullptr, nullptr, 0, stream);
}
int fmb_trace2d_logged_f64(const double *d_T, int64_t T_pitch, int64_t T_qstride, int rows, int cols, int npaths,
                           const int32_t *d_field_of_path, const double *d_init, const double *d_end, double tau,
                           int max_steps, double *d_out, int64_t cap, int32_t *d_count, int32_t *d_status,
                           int32_t *d_blocks, int32_t *d_nblocks, int cap_blocks, void *stream) {
    if (!d_blocks || !d_nblocks || cap_blocks < 1) return fail(FMB_E_INVALID, "bad block log%s");
    return trace2d_impl(d_T, T_pitch, T_qstride, rows, cols, npaths, d_field_of_path, d_init, d_end, tau, max_steps, d_out, cap,
                        d_count, d_status, d_blocks, d_nblocks, cap_blocks, stream);
}
int fmb_windows_differ_f64(const double *d_field, int64_t pitch, int rows, int cols, const double *h_field, int64_t hs_y, int64_t hs_x,
                           const int32_t *d_blocks, const int32_t *d_nblocks, int cap_blocks, int32_t *d_flag, void *stream) {
    if (!d_field || !h_field || !d_blocks || !d_nblocks || !d_flag || cap_blocks < 1) return fail(FMB_E_INVALID, "bad argument%s");
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, h_field) != cudaSuccess || pa.type != cudaMemoryTypeHost || pa.devicePointer != (void *)h_field) {
        cudaGetLastError();
        return fail(FMB_E_INVALID, "h_field must be page-locked host memory the device addresses by the same pointer%s");
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), st), "cudaMemsetAsync(flag)");
    int blocks = cap_blocks < 2 * sm_count() ? cap_blocks : 2 * sm_count();
    fmb::windows_differ_kernel<<<blocks, 128, 0, st>>>(d_field, pitch, rows, cols, h_field, hs_y, hs_x, d_blocks, d_nblocks, cap_blocks, d_flag);
    CK(cudaGetLastError(), "launch windows_differ");
    return FMB_OK;
}

}  // extern "C"

#include "fm_capi3d.inc"
#include "fm_capi_costmap.inc"
#include "fm_capi_pathpost.inc"
#include "fm_capi_ranks.inc"
#include "fm_capi_host.inc"
