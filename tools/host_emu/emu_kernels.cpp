// emu_kernels.cpp -- runs the device code of planning_motion_planning_b200/csrc on the
// CPU through cuda_emu.h.  TEST TOOLING ONLY (see cuda_emu.h); built by tests/ as
// tools/host_emu/libfm_emu.so and never loaded by the product package.
//
//   g++ -O1 -g -std=c++17 -DFMB_HOST_EMU -ffp-contract=off -fPIC -shared \
//       -I tools/host_emu -o tools/host_emu/libfm_emu.so tools/host_emu/emu_kernels.cpp
#include <algorithm>
#include "cuda_emu.h"

#include "../../planning_motion_planning_b200/csrc/eikonal2d.cuh"
#include "../../planning_motion_planning_b200/csrc/eikonal2d_cta.cuh"
#include "../../planning_motion_planning_b200/csrc/eikonal2d_sweep.cuh"
#include "../../planning_motion_planning_b200/csrc/eikonal2d_wsweep.cuh"
#include "../../planning_motion_planning_b200/csrc/eikonal3d_sweep.cuh"
#include "../../planning_motion_planning_b200/csrc/eikonal3d.cuh"
#include "../../planning_motion_planning_b200/csrc/pow2_glibc.cuh"
#include "../../planning_motion_planning_b200/csrc/trace2d.cuh"
#include "../../planning_motion_planning_b200/csrc/trace3d.cuh"
#include "../../planning_motion_planning_b200/csrc/truncate.cuh"
#include "../../planning_motion_planning_b200/csrc/tiekeys.cuh"
#include "../../planning_motion_planning_b200/csrc/costmap2d.cuh"
#include "../../planning_motion_planning_b200/csrc/costvolume.cuh"
#include "../../include/fm_b200.h"

namespace {
unsigned pow2_at_least(long long v) { unsigned p = 1024; while ((long long)p < v) p <<= 1; return p; }
constexpr int WARPS = 4;

template <typename real, int TW>
int run2d(const real *cost, long long cost_qstride, real *T, int rows, int cols, int nq, const int *seeds, int nblocks,
          unsigned long long *stats) {
    fmb::Problem2D<real> P;
    P.cost = cost; P.cost_pitch = cols; P.cost_qstride = cost_qstride;
    P.T = T; P.T_pitch = cols; P.T_qstride = (long long)rows * cols;
    P.rows = rows; P.cols = cols; P.nq = nq;
    P.ntx = (cols + TW - 1) / TW; P.nty = (rows + fmb::TILE_H - 1) / fmb::TILE_H;
    P.seeds = seeds;
    const long long ntiles = (long long)nq * P.ntx * P.nty;
    std::vector<int> state(ntiles), ring(pow2_at_least(ntiles));
    fmb::QueueCtl ctl = {};
    P.tile_state = state.data(); P.q.ctl = &ctl; P.q.ring = ring.data(); P.q.ring_mask = (unsigned)ring.size() - 1;
    P.q.watchdog_cycles = 1LL << 40; P.step_cap = 1 << 20;
    std::vector<unsigned long long> prio(ntiles);
    P.tile_prio = prio.data();
    P.best_first = getenv("FMB_BEST_FIRST") ? atoi(getenv("FMB_BEST_FIRST")) : 0;
    P.arm_rows = 0;
    std::vector<int> lev_count(fmb::WIN_LEVELS), tile_level(ntiles);
    int win_hint = 0; double win_inv_delta = 1.0;
    P.windowed = (!P.best_first && nq == 1 && getenv("FMB_WINDOWED")) ? atoi(getenv("FMB_WINDOWED")) : 0;
    P.win_window = getenv("FMB_WINDOW") ? atoi(getenv("FMB_WINDOW")) : 16; P.win_div = 1; P.win_running = 0; P.check_passes = 1; P.precheck = 0; P.pipeline = getenv("FMB_PIPELINE") ? atoi(getenv("FMB_PIPELINE")) : 1;
    P.lev_count = lev_count.data(); P.tile_level = tile_level.data(); P.win_hint = &win_hint; P.win_inv_delta = &win_inv_delta;
    emu::launch(2, 64, 0, [&] { fmb::init_fill2d_kernel<real>(P, (int)ring.size()); });
    emu::launch(1, 32 * ((nq + 31) / 32), 0, [&] { fmb::init_seed2d_kernel<real, TW>(P); });
    const bool cg = getenv("FMB_COST_GLOBAL") && atoi(getenv("FMB_COST_GLOBAL"));     // engine2d = 6: cost read from global memory
    if (P.best_first && cg) emu::launch(nblocks, WARPS * 32, sizeof(real) * fmb::Tile2D<real, TW>::T_ELEMS * WARPS, [&] { fmb::solve2d_kernel<real, TW, WARPS, true, true>(P); });
    else if (P.best_first) emu::launch(nblocks, WARPS * 32, fmb::Tile2D<real, TW>::WARP_BYTES * WARPS, [&] { fmb::solve2d_kernel<real, TW, WARPS, true>(P); });
    else emu::launch(nblocks, WARPS * 32, fmb::Tile2D<real, TW>::WARP_BYTES * WARPS, [&] { fmb::solve2d_kernel<real, TW, WARPS, false>(P); });
    if (stats) { stats[0] = ctl.visits; stats[1] = ctl.steps; stats[2] = ctl.evals; stats[3] = ctl.pushes; stats[4] = ctl.cells_written; }
    return ctl.abort ? ctl.abort : (ctl.pending != 0 ? -1 : 0);
}

// CTA-per-tile engine (eikonal2d_cta.cuh): same set-up, 1024 / R threads per block
template <typename real, int R>
int run2d_cta(const real *cost, long long cost_qstride, real *T, int rows, int cols, int nq, const int *seeds, int nblocks,
              int best_first, int windowed, int window, unsigned long long *stats) {
    constexpr int TW = 32;
    fmb::Problem2D<real> P;
    P.cost = cost; P.cost_pitch = cols; P.cost_qstride = cost_qstride;
    P.T = T; P.T_pitch = cols; P.T_qstride = (long long)rows * cols;
    P.rows = rows; P.cols = cols; P.nq = nq;
    P.ntx = (cols + TW - 1) / TW; P.nty = (rows + fmb::TILE_H - 1) / fmb::TILE_H;
    P.seeds = seeds;
    const long long ntiles = (long long)nq * P.ntx * P.nty;
    std::vector<int> state(ntiles), ring(pow2_at_least(ntiles));
    fmb::QueueCtl ctl = {};
    P.tile_state = state.data(); P.q.ctl = &ctl; P.q.ring = ring.data(); P.q.ring_mask = (unsigned)ring.size() - 1;
    P.q.watchdog_cycles = 1LL << 40; P.step_cap = 1 << 20;
    std::vector<unsigned long long> prio(ntiles);
    P.tile_prio = prio.data();
    P.best_first = best_first;
    P.arm_rows = 0;
    std::vector<int> lev_count(fmb::WIN_LEVELS), tile_level(ntiles);
    int win_hint = 0; double win_inv_delta = 1.0;
    P.windowed = (!best_first && nq == 1) ? windowed : 0;
    P.check_passes = 4; P.precheck = 0; P.pipeline = getenv("FMB_PIPELINE") ? atoi(getenv("FMB_PIPELINE")) : 1; P.win_window = window; P.win_div = R == 0 ? 2 : 1; P.win_running = R == 0 ? 1 : 0;
    P.lev_count = lev_count.data(); P.tile_level = tile_level.data(); P.win_hint = &win_hint; P.win_inv_delta = &win_inv_delta;
    std::vector<unsigned long long> run_prio(ntiles);
    P.run_prio = run_prio.data();
    double slack[2] = {0.0, 0.0}; P.slack = slack; P.slack_frac = getenv("FMB_EMU_SLACK") ? atof(getenv("FMB_EMU_SLACK")) : 0.0;
    P.hop_frac = getenv("FMB_EMU_RING2") ? atof(getenv("FMB_EMU_RING2")) : 1.0;
    P.variant = getenv("FMB_EMU_VARIANT") ? atoi(getenv("FMB_EMU_VARIANT")) : 0;
    emu::launch(2, 64, 0, [&] { fmb::init_fill2d_kernel<real>(P, (int)ring.size()); });
    emu::launch(1, 32 * ((nq + 31) / 32), 0, [&] { fmb::init_seed2d_kernel<real, TW>(P); });
    if (R == 0 && getenv("FMB_EMU_WSWEEP") && atoi(getenv("FMB_EMU_WSWEEP"))) {          // warp-per-tile sweep engine (1: costs staged, 2: costs from global)
        const bool stage = atoi(getenv("FMB_EMU_WSWEEP")) == 1;
        using TL = fmb::Tile2D<real, 32>;
        const size_t smem = sizeof(real) * (TL::T_ELEMS + (stage ? TL::C_ELEMS : 0)) * 4;
        if (P.windowed == 1) P.windowed = 0;
        if (best_first) { if (stage) emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_wsweep_kernel<real, true, true>(P); }); else emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_wsweep_kernel<real, true, false>(P); }); }
        else { if (stage) emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_wsweep_kernel<real, false, true>(P); }); else emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_wsweep_kernel<real, false, false>(P); }); }
    } else if (R == 0) {          // sweep engine
        const size_t smem = fmb::Sweep2DSmem::bytes<real>();
        if (best_first) emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_sweep_kernel<real, true>(P); });
        else emu::launch(nblocks, 128, smem, [&] { fmb::solve2d_sweep_kernel<real, false>(P); });
    } else {
        const size_t smem = fmb::CtaTile2D<real>::BYTES;
        constexpr int RR = R ? R : 1;
        if (best_first) emu::launch(nblocks, 1024 / RR, smem, [&] { fmb::solve2d_cta_kernel<real, RR, true>(P); });
        else emu::launch(nblocks, 1024 / RR, smem, [&] { fmb::solve2d_cta_kernel<real, RR, false>(P); });
    }
    if (stats) { stats[0] = ctl.visits; stats[1] = ctl.steps; stats[2] = ctl.evals; stats[3] = ctl.pushes; stats[4] = ctl.cells_written; }
    return ctl.abort ? ctl.abort : (ctl.pending != 0 ? -1 : 0);
}

template <typename real, int TZ>
int run3d(const real *cost, long long cost_qstride, real *T, int ny, int nx, int nz, int nq, const int *seeds, int nblocks,
          unsigned long long *stats, bool sweep = false) {
    fmb::Problem3D<real> P;
    P.cost = cost; P.cost_qstride = cost_qstride; P.T = T; P.T_qstride = (long long)ny * nx * nz;
    P.ny = ny; P.nx = nx; P.nz = nz; P.nq = nq;
    P.nty = (ny + fmb::T3Y - 1) / fmb::T3Y; P.ntx = (nx + fmb::T3X - 1) / fmb::T3X; P.ntz = (nz + TZ - 1) / TZ;
    P.seeds = seeds;
    const long long ntiles = (long long)nq * P.nty * P.ntx * P.ntz;
    std::vector<int> state(ntiles), ring(pow2_at_least(ntiles));
    fmb::QueueCtl ctl = {};
    P.tile_state = state.data(); P.q.ctl = &ctl; P.q.ring = ring.data(); P.q.ring_mask = (unsigned)ring.size() - 1;
    P.q.watchdog_cycles = 1LL << 40; P.step_cap = 1 << 20;
    std::vector<unsigned long long> prio(ntiles), run_prio(ntiles);
    P.tile_prio = prio.data(); P.run_prio = run_prio.data(); P.causal = sweep ? 1 : 0; P.check_passes = 4; P.arm_all = 0;
    double slack[2] = {0.0, 0.0}; P.slack = slack; P.slack_frac = getenv("FMB_EMU_SLACK") ? atof(getenv("FMB_EMU_SLACK")) : 0.0;
    P.hop_frac = getenv("FMB_EMU_RING2") ? atof(getenv("FMB_EMU_RING2")) : 1.0;
    P.variant = getenv("FMB_EMU_VARIANT") ? atoi(getenv("FMB_EMU_VARIANT")) : 0;
    P.enable = nullptr;
    using TL16 = fmb::Tile3D<real, 16>;
    const size_t smem_sweep = sizeof(real) * (TL16::T_ELEMS + TL16::C_ELEMS + 8) + 32 * sizeof(unsigned) + 4 * sizeof(int);
    emu::launch(2, 64, 0, [&] { fmb::init_fill3d_kernel<real>(P, (int)ring.size()); });
    emu::launch(1, 32 * ((nq + 31) / 32), 0, [&] { fmb::init_seed3d_kernel<real, TZ>(P); });
    if (sweep) emu::launch(nblocks, 256, smem_sweep, [&] { fmb::solve3d_sweep_kernel<real, false>(P); });
    else emu::launch(nblocks, WARPS * 32, fmb::Tile3D<real, TZ>::WARP_BYTES * WARPS, [&] { fmb::solve3d_kernel<real, TZ, WARPS>(P); });
    if (getenv("FMB_EMU_POLISH3D") && atoi(getenv("FMB_EMU_POLISH3D")) && sizeof(real) == 8 && !ctl.abort && ctl.pending == 0) {
        P.arm_all = 1;
        emu::launch(2, 64, 0, [&] { fmb::init_resume3d_kernel<real>(P, (int)ring.size()); });
        emu::launch((unsigned)((ntiles + 63) / 64), 64, 0, [&] { fmb::activate_all3d_kernel<real>(P); });
        if (sweep) emu::launch(nblocks, 256, smem_sweep, [&] { fmb::solve3d_sweep_kernel<real, true>(P); });
        else emu::launch(nblocks, WARPS * 32, fmb::Tile3D<real, TZ>::WARP_BYTES * WARPS, [&] { fmb::solve3d_kernel<real, TZ, WARPS, true>(P); });
    }
    if (stats) { stats[0] = ctl.visits; stats[1] = ctl.steps; stats[2] = ctl.evals; stats[3] = ctl.pushes; stats[4] = ctl.cells_written; }
    return ctl.abort ? ctl.abort : (ctl.pending != 0 ? -1 : 0);
}
}  // namespace

// csrc/tiekeys.cuh: exact LIFO pop order in one ordered sweep (2D and 3D)
template <int D>
static int emu_tie_order(fmb::Grid<D> g, const double *T, const double *cost, const int *members, const int *gstart,
                         const int *gsize, int seed_idx, int transposed, int *rank, int *tau) {
    const size_t n = (size_t)g.size();
    std::vector<long long> key(n);
    std::vector<int> scratch(2 * n + 2, 0);
    emu::launch(4, 64, 0, [&] {
        fmb::tie_sweep_kernel<D>(g, T, cost, members, gstart, gsize, seed_idx, transposed, rank, tau, key.data(), scratch.data(),
                                 scratch.data() + n, scratch.data() + 2 * n, scratch.data() + 2 * n + 1);
    });
    return scratch[2 * n + 1];
}

// dense replay (as fmb_truncate*_f64), or -- FMB_REPLAY_SPARSE=1 -- the sparse one as truncate_dk runs it: cone expansion (two
// grid-wide rounds, then the one-block tail), tickets from the cell side, sorted, list sweep.  Returns the waits at the
// limit, or -1 when the sparse form reported failure.
template <int D>
int emu_truncate(fmb::Grid<D> g, const double *F, const double *cost, const int *rank, int k, double *out) {
    constexpr int NN = fmb::Grid<D>::NN;
    const size_t n = (size_t)g.size();
    std::vector<int> list(n); int counters[64] = {0};
    std::vector<double> memo(n * NN);
    emu::launch(4, 64, 0, [&] { fmb::truncate_mark_kernel<double, D>(g, F, rank, k, out, list.data()); });
    if (!(getenv("FMB_REPLAY_SPARSE") && atoi(getenv("FMB_REPLAY_SPARSE")))) {
        memset(memo.data(), 0xff, memo.size() * sizeof(double));   // as fmb_truncate2d_f64 does
        emu::launch(4, 64, 0, [&] { fmb::truncate_sweep_kernel<double, D>(g, F, cost, rank, list.data(), k, out, memo.data(), counters, counters + 1); });
        return counters[1];
    }
    std::vector<int> need(n), fa(n), fb(n), tickets(n * NN, fmb::CONE_TICKET_PAD);
    const int cap = (int)n, grid_rounds = 2;
    const int *kd = &k;
    emu::launch(4, 64, 0, [&] { fmb::cone_seed_kernel<double, D>(g, cost, rank, kd, need.data(), fa.data(), counters, cap); });
    for (int r = 0; r < grid_rounds; ++r)
        emu::launch(4, 64, 0, [&] { fmb::cone_expand_kernel<double, D>(g, cost, rank, kd, need.data(), (r & 1) ? fb.data() : fa.data(), (r & 1) ? fa.data() : fb.data(), counters, r, cap); });
    emu::launch(1, 128, fmb::CONE_TAIL_SMEM, [&] { fmb::cone_tail_kernel<double, D>(g, cost, rank, kd, need.data(), fa.data(), fb.data(), counters, grid_rounds, cap, 1 << 16); });
    emu::launch(4, 64, 0, [&] { fmb::cone_emit_kernel<double, D>(g, rank, kd, need.data(), tickets.data(), memo.data(), counters, (int)tickets.size()); });
    std::sort(tickets.begin(), tickets.end());
    emu::launch(4, 64, 0, [&] { fmb::truncate_sweep_list_kernel<double, D>(g, F, cost, rank, list.data(), out, memo.data(), tickets.data(), counters, kd, (int)tickets.size()); });
    return counters[5] ? -1 : counters[1];
}

extern "C" {

int emu_solve2d_f64(const double *cost, long long cost_qstride, double *T, int rows, int cols, int nq, const int *seeds,
                    int tw, int nblocks, unsigned long long *stats) {
    if (tw == 16) return run2d<double, 16>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, stats);
    return run2d<double, 32>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, stats);
}
int emu_solve2d_f32(const float *cost, long long cost_qstride, float *T, int rows, int cols, int nq, const int *seeds,
                    int tw, int nblocks, unsigned long long *stats) {
    if (tw == 16) return run2d<float, 16>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, stats);
    return run2d<float, 32>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, stats);
}
int emu_solve2d_cta_f64(const double *cost, long long cost_qstride, double *T, int rows, int cols, int nq, const int *seeds,
                        int R, int nblocks, int best_first, int windowed, int window, unsigned long long *stats) {
    if (R == 0) return run2d_cta<double, 0>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    if (R == 1) return run2d_cta<double, 1>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    if (R == 4) return run2d_cta<double, 4>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    return run2d_cta<double, 2>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
}
int emu_solve2d_cta_f32(const float *cost, long long cost_qstride, float *T, int rows, int cols, int nq, const int *seeds,
                        int R, int nblocks, int best_first, int windowed, int window, unsigned long long *stats) {
    if (R == 0) return run2d_cta<float, 0>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    if (R == 1) return run2d_cta<float, 1>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    if (R == 4) return run2d_cta<float, 4>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
    return run2d_cta<float, 2>(cost, cost_qstride, T, rows, cols, nq, seeds, nblocks, best_first, windowed, window, stats);
}
int emu_solve3d_f64(const double *cost, long long cost_qstride, double *T, int ny, int nx, int nz, int nq,
                    const int *seeds, int tz, int nblocks, unsigned long long *stats) {
    if (tz == 0) return run3d<double, 16>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats, true);      // sweep engine
    if (tz == 16) return run3d<double, 16>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats);
    return run3d<double, 32>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats);
}
int emu_solve3d_f32(const float *cost, long long cost_qstride, float *T, int ny, int nx, int nz, int nq,
                    const int *seeds, int tz, int nblocks, unsigned long long *stats) {
    if (tz == 0) return run3d<float, 16>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats, true);
    if (tz == 16) return run3d<float, 16>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats);
    return run3d<float, 32>(cost, cost_qstride, T, ny, nx, nz, nq, seeds, nblocks, stats);
}

void emu_trace2d_f64(const double *T, int rows, int cols, int npaths, const int *field_of_path, const double *init,
                     const double *end, double tau, int max_steps, double *out, long long cap, int *count, int *status) {
    fmb::TraceArgs2D<double> A;
    A.T = T; A.T_pitch = cols; A.T_qstride = (long long)rows * cols; A.rows = rows; A.cols = cols; A.npaths = npaths;
    A.field_of_path = field_of_path; A.init = init; A.end = end; A.tau = tau; A.max_steps = max_steps;
    A.out = out; A.cap = cap; A.count = count; A.status = status;
    emu::launch((npaths + 3) / 4, 128, 4 * fmb::TRACE2D_SMEM_PER_WARP, [&] { fmb::trace2d_kernel<double, 4>(A); });
}
void emu_trace3d_f64(const double *T, int ny, int nx, int nz, int npaths, const int *field_of_path, const double *init,
                     const double *end, double tau, int max_steps, double *out, long long cap, int *count, int *status) {
    fmb::TraceArgs3D<double> A;
    A.T = T; A.T_qstride = (long long)ny * nx * nz; A.ny = ny; A.nx = nx; A.nz = nz; A.npaths = npaths;
    A.field_of_path = field_of_path; A.init = init; A.end = end; A.tau = tau; A.max_steps = max_steps;
    A.out = out; A.cap = cap; A.count = count; A.status = status;
    emu::launch((npaths + 3) / 4, 128, 0, [&] { fmb::trace3d_kernel<double, 4>(A); });
}

int emu_truncate2d_f64(const double *F, const double *cost, const int *rank, int rows, int cols, int k, double *out) {
    fmb::Grid<2> g; g.rows = rows; g.cols = cols;
    return emu_truncate<2>(g, F, cost, rank, k, out);
}
int emu_truncate3d_f64(const double *F, const double *cost, const int *rank, int ny, int nx, int nz, int k, double *out) {
    fmb::Grid<3> g; g.ny = ny; g.nx = nx; g.nz = nz;
    return emu_truncate<3>(g, F, cost, rank, k, out);
}

void emu_pow2(const double *x, double *out, long long n) { for (long long i = 0; i < n; ++i) out[i] = fmb::pow2_glibc(x[i]); }
// the two forms of the 3D update on n (t0, t1, t2, c) tuples: out_sel / out_ref = branch-free / branching, slow = flag of the former
void emu_update3d(const double *t, long long n, int exact, double *out_sel, double *out_ref, int *slow) {
    for (long long i = 0; i < n; ++i) {
        bool sl = false;
        const double *a = t + 4 * i;
        if (exact == 1) { out_sel[i] = fmb::solve3d_update_sel<true>(a[0], a[1], a[2], a[3], sl); out_ref[i] = fmb::solve3d_update_exact(a[0], a[1], a[2], a[3]); }
        else { out_sel[i] = fmb::solve3d_update_sel<false>(a[0], a[1], a[2], a[3], sl); out_ref[i] = fmb::solve3d_update<double>(a[0], a[1], a[2], a[3]); }
        slow[i] = sl;
    }
}
void emu_div3(const double *x, double *out, long long n) { for (long long i = 0; i < n; ++i) out[i] = fmb::num<double>::div3(x[i]); }

// resume a 2D solve from the current contents of T (domain decomposition tests)
int emu_resolve2d_f64(const double *cost, double *T, int rows, int cols, const int *seed, int activate, int halo_rows,
                      int nblocks) {
    constexpr int TW = 32;
    fmb::Problem2D<double> P;
    P.cost = cost; P.cost_pitch = cols; P.cost_qstride = 0; P.T = T; P.T_pitch = cols; P.T_qstride = 0;
    P.rows = rows; P.cols = cols; P.nq = 1;
    P.ntx = (cols + TW - 1) / TW; P.nty = (rows + fmb::TILE_H - 1) / fmb::TILE_H;
    P.seeds = seed;
    const long long ntiles = (long long)P.ntx * P.nty;
    std::vector<int> state(ntiles), ring(pow2_at_least(ntiles));
    std::vector<unsigned long long> prio(ntiles);
    fmb::QueueCtl ctl = {};
    P.tile_state = state.data(); P.q.ctl = &ctl; P.q.ring = ring.data(); P.q.ring_mask = (unsigned)ring.size() - 1;
    P.q.watchdog_cycles = 1LL << 40; P.step_cap = 1 << 20; P.tile_prio = prio.data(); P.best_first = 0; P.arm_rows = halo_rows; P.windowed = 0; P.win_window = 0; P.win_div = 1; P.win_running = 0; P.check_passes = 1; P.precheck = 0; P.pipeline = getenv("FMB_PIPELINE") ? atoi(getenv("FMB_PIPELINE")) : 1;
    P.lev_count = nullptr; P.tile_level = nullptr; P.win_hint = nullptr; P.win_inv_delta = nullptr;
    emu::launch(2, 64, 0, [&] { fmb::init_resume2d_kernel<double>(P, (int)ring.size()); });
    if (activate & 7) emu::launch((unsigned)((ntiles + 63) / 64), 64, 0, [&] { fmb::activate_rows2d_kernel<double>(P, activate); });
    emu::launch(1, 32, 0, [&] { fmb::init_seed2d_kernel<double, TW>(P); });
    emu::launch(nblocks, WARPS * 32, fmb::Tile2D<double, TW>::WARP_BYTES * WARPS, [&] { fmb::solve2d_kernel<double, TW, WARPS, false>(P); });
    return ctl.abort ? ctl.abort : (ctl.pending != 0 ? -1 : 0);
}

// the whole cost-map pipeline of fm_capi_costmap.inc, same kernel sequence (2 blocks x 64 threads)
int emu_costmap2d_f64(const double *dem, const double *grid, int n, double resolution, double slope_max, int r_open,
                      int r_close, int r_expand, double *cost, unsigned char *raw, unsigned char *obst, double *pre,
                      int *n_positive) {
    const size_t nn = (size_t)n * n;
    std::vector<unsigned char> A(nn), B(nn);
    std::vector<int> g(nn), lab(nn);
    std::vector<double> tmp(nn), pre_own(nn);
    fmb::CostmapCtl ctl;
    const unsigned G = 2, T = 64;
    auto fill = [&](unsigned char *im) {
        emu::launch(G, T, 0, [&] { fmb::cm_runs_kernel(im, n, lab.data()); });
        emu::launch(G, T, 0, [&] { fmb::cm_link_kernel(im, n, lab.data()); });
        emu::launch(G, T, 0, [&] { fmb::cm_fill_apply_kernel(im, n, lab.data(), im); });
    };
    std::vector<int> seg_first((size_t)((n + fmb::CM_SEG - 1) / fmb::CM_SEG) * n), seg_last(seg_first.size());
    auto vscan = [&](const unsigned char *src, int fv) {
        emu::launch(G, T, 0, [&] { fmb::cm_vseg_kernel(src, fv, n, seg_first.data(), seg_last.data()); });
        emu::launch(G, T, 0, [&] { fmb::cm_vscan_full_kernel(src, fv, n, seg_first.data(), seg_last.data(), g.data()); });
    };
    auto morph = [&](const unsigned char *src, unsigned char *dst, int r, bool dilate) {
        vscan(src, dilate ? 1 : 0);
        emu::launch(G, T, 0, [&] { fmb::cm_hscan_threshold_kernel(g.data(), n, r, dilate ? 1 : 0, dst); });
    };
    emu::launch(1, 32, 0, [&] { fmb::cm_init_ctl_kernel(&ctl); });
    emu::launch(G, T, 0, [&] { fmb::cm_slope_kernel(dem, grid, n, slope_max, A.data()); });
    fill(A.data());
    if (raw) memcpy(raw, A.data(), nn);
    morph(A.data(), B.data(), r_open, false);
    morph(B.data(), A.data(), r_open, true);
    morph(A.data(), B.data(), r_close, true);
    fill(B.data());
    morph(B.data(), A.data(), r_close, false);
    emu::launch(1, 64, 0, [&] { fmb::cm_set_border_kernel(A.data(), n, 1); });
    if (obst) memcpy(obst, A.data(), nn);
    morph(A.data(), B.data(), r_expand, true);
    vscan(A.data(), 1);
    emu::launch(G, T, 0, [&] { fmb::cm_hscan_exact_kernel(g.data(), n, lab.data(), &ctl); });
    emu::launch(G, T, 0, [&] { fmb::cm_band_min_kernel(B.data(), lab.data(), n, resolution, &ctl); });
    double *p = pre ? pre : pre_own.data();
    emu::launch(G, T, 0, [&] { fmb::cm_compose_kernel(A.data(), B.data(), lab.data(), n, resolution, &ctl, p); });
    emu::launch(G, T, 0, [&] { fmb::cm_blur_rows_kernel(p, n, tmp.data()); });
    emu::launch(G, T, 0, [&] { fmb::cm_blur_cols_kernel(tmp.data(), n, cost); });
    if (n_positive) *n_positive = ctl.n_positive;
    return 0;
}

// csrc/costvolume.cuh with host pointers in the descriptor (same kernel sequence as fmb_costvolume_f64)
int emu_costvolume_f64(const fmb_costvolume_desc *d, double *cmap, double *tunnel, double *terrain) {
    const size_t cells = (size_t)d->sX * d->sY * d->sZ;
    std::vector<int> first(cells);
    std::vector<unsigned char> blocked(cells);
    fmb::CostVolumeArgs A;
    A.Zs = d->d_Zs; A.zm = d->zs_rows; A.zn = d->zs_cols;
    A.resX = d->resX; A.resY = d->resY; A.resZ = d->resZ; A.xm = d->xm; A.ym = d->ym;
    A.sX = d->sX; A.sY = d->sY; A.sZ = d->sZ;
    A.frames = d->d_frames; A.npose = d->npose;
    A.li = d->d_li; A.lk = d->d_lk; A.nX = d->nX; A.nZ = d->nZ;
    A.norm = d->d_norm; A.val = d->d_val; A.rlim = d->rlim;
    A.lr = d->d_lr; A.hval = d->d_hval; A.nK = d->nK;
    A.ct = d->d_angles; A.st = d->d_angles + 100; A.cs = d->d_angles + 200; A.ss = d->d_angles + 290;
    A.shell = d->shell;
    for (int k = 0; k < 3; ++k) { A.fin[k] = d->fin[k]; A.ini[k] = d->ini[k]; }
    A.first = first.data(); A.blocked = blocked.data();
    A.cmap = cmap; A.tunnel = tunnel; A.terrain = terrain;
    emu::launch(2, 64, 0, [&] { fmb::cv_init_kernel(A); });
    emu::launch(2, 64, 0, [&] { fmb::cv_scatter_kernel(A); });
    emu::launch(2, 64, 0, [&] { fmb::cv_compose_kernel(A); });
    return 0;
}

int emu_tie_order2d(const double *T, const double *cost, const int *members, const int *gstart, const int *gsize, int rows,
                    int cols, int seed_idx, int transposed, int *rank, int *tau) {
    fmb::Grid<2> g; g.rows = rows; g.cols = cols;
    return emu_tie_order<2>(g, T, cost, members, gstart, gsize, seed_idx, transposed, rank, tau);
}
int emu_tie_order3d(const double *T, const double *cost, const int *members, const int *gstart, const int *gsize, int ny,
                    int nx, int nz, int seed_idx, int *rank, int *tau) {
    fmb::Grid<3> g; g.ny = ny; g.nx = nx; g.nz = nz;
    return emu_tie_order<3>(g, T, cost, members, gstart, gsize, seed_idx, 0, rank, tau);
}

// csrc/tiekeys.cuh: join of the two fronts (same kernel sequence as fmb_bi_join)
int emu_bi_join(const int *rankG, const int *rankS, long long total, int *out2) {
    alignas(8) int out[4] = {0x7fffffff, 0x7fffffff, -1, -1};
    emu::launch(2, 64, 0, [&] { fmb::bi_join_k_kernel(rankG, rankS, total, out); });
    emu::launch(2, 64, 0, [&] { fmb::bi_join_cell_kernel(rankG, rankS, total, out, (unsigned long long *)(out + 2)); });
    emu::launch(1, 32, 0, [&] { fmb::bi_join_finish_kernel(out, (const unsigned long long *)(out + 2)); });
    out2[0] = out[0]; out2[1] = out[1];
    return 0;
}

}  // extern "C"
