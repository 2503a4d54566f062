// cuda_emu.h -- minimal host-side SIMT emulation used ONLY by tests/tools.
//
// It lets the device code in planning_motion_planning_b200/csrc/*.cuh be compiled
// by g++ (-DFMB_HOST_EMU) and executed on the CPU so that the kernel LOGIC (tile
// staging, masks, queue protocol, tracer quirks) can be checked against the oracle in
// the GPU-less build container before GPU minutes are spent.  It is test tooling:
// the shipped library never includes this header and there is no CPU path in the
// product.
//
// Model: every lane is a ucontext coroutine; a warp's 32 lanes meet at collectives
// (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp), which also asserts that all
// lanes of the warp reached the SAME collective (catches divergence bugs).  Warps are
// scheduled round-robin on one OS thread; __nanosleep yields to the other warps.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n)
#define __restrict__

namespace emu {

struct dim3_ { unsigned x = 1, y = 1, z = 1; };

enum Coll { C_NONE = 0, C_SHFL, C_SHFL_UP, C_SHFL_DOWN, C_SHFL_XOR, C_BALLOT, C_ANY, C_SYNC, C_BLOCK };   // C_BLOCK: __syncthreads[_or]

struct Warp;
struct Lane {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false, waiting = false;
    int coll = C_NONE, arg = 0;
    uint64_t payload = 0, result = 0;
    dim3_ tid, bid;
    Warp *warp = nullptr;
    int lane_id = 0;
};
struct Warp {
    Lane lanes[32];
    unsigned char *smem = nullptr;     // block shared memory (shared by the warps of a block)
    int block = 0;                     // block index (block-wide barriers)
};

struct Machine {
    std::vector<Warp *> warps;
    dim3_ grid, block;
    ucontext_t sched;
    Lane *cur = nullptr;
    bool yielded_sleep = false;
    uint64_t clock = 0;
    std::function<void()> body;
};
inline Machine &M() { static Machine m; return m; }

inline void yield_to_sched() { Lane *l = M().cur; swapcontext(&l->ctx, &M().sched); }

inline uint64_t collective(int kind, uint64_t payload, int arg) {
    Lane *l = M().cur;
    l->coll = kind; l->payload = payload; l->arg = arg; l->waiting = true;
    yield_to_sched();
    return l->result;
}

inline void lane_entry() {
    M().body();
    M().cur->done = true;
    yield_to_sched();
}

inline bool block_resolve(Warp *w);
inline void resolve(Warp *w) {
    int kind = C_NONE;
    for (auto &l : w->lanes) {
        if (l.done) continue;
        if (kind == C_NONE) kind = l.coll;
        if (l.coll != kind) { fprintf(stderr, "EMU: divergent collectives in a warp (%d vs %d)\n", kind, l.coll); abort(); }
    }
    for (auto &l : w->lanes) if (l.done) { fprintf(stderr, "EMU: collective with exited lanes\n"); abort(); }
    if (kind == C_BLOCK) { block_resolve(w); return; }
    unsigned ballot = 0;
    for (int i = 0; i < 32; ++i) if (w->lanes[i].payload) ballot |= 1u << i;
    for (int i = 0; i < 32; ++i) {
        Lane &l = w->lanes[i];
        switch (kind) {
            case C_SHFL: l.result = w->lanes[l.arg & 31].payload; break;
            case C_SHFL_UP: l.result = (i - l.arg >= 0) ? w->lanes[i - l.arg].payload : l.payload; break;
            case C_SHFL_DOWN: l.result = (i + l.arg < 32) ? w->lanes[i + l.arg].payload : l.payload; break;
            case C_SHFL_XOR: l.result = w->lanes[(i ^ l.arg) & 31].payload; break;
            case C_BALLOT: l.result = ballot; break;
            case C_ANY: l.result = ballot != 0; break;
            default: l.result = 0; break;
        }
    }
    for (auto &l : w->lanes) l.waiting = false;
}

// __syncthreads / __syncthreads_or: released when every lane of every warp of the block waits on it
// (whole warps that already exited are ignored, like on the device)
inline bool block_resolve(Warp *w) {
    Machine &m = M();
    uint64_t any = 0;
    for (Warp *o : m.warps) {
        if (o->block != w->block) continue;
        for (auto &l : o->lanes) {
            if (l.done) continue;
            if (!l.waiting || l.coll != C_BLOCK) return false;
            any |= l.payload;
        }
    }
    for (Warp *o : m.warps) {
        if (o->block != w->block) continue;
        for (auto &l : o->lanes) if (!l.done) { l.result = any; l.waiting = false; }
    }
    return true;
}

// run `body` as a kernel of grid x block threads (block.x multiple of 32)
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> body) {
    Machine &m = M();
    m.body = body; m.grid.x = grid; m.block.x = block;
    const unsigned wpb = block / 32;
    std::vector<unsigned char *> smems;
    for (unsigned b = 0; b < grid; ++b) {
        unsigned char *sm = (unsigned char *)aligned_alloc(64, ((smem_bytes + 63) / 64) * 64 + 64);
        smems.push_back(sm);
        for (unsigned wi = 0; wi < wpb; ++wi) {
            Warp *w = new Warp(); w->smem = sm; w->block = (int)b;
            for (int i = 0; i < 32; ++i) {
                Lane &l = w->lanes[i];
                l.warp = w; l.lane_id = i; l.tid.x = wi * 32 + i; l.bid.x = b;
                l.stack = (char *)malloc(256 * 1024);
                getcontext(&l.ctx);
                l.ctx.uc_stack.ss_sp = l.stack; l.ctx.uc_stack.ss_size = 256 * 1024; l.ctx.uc_link = &m.sched;
                makecontext(&l.ctx, (void (*)())lane_entry, 0);
            }
            m.warps.push_back(w);
        }
    }
    bool alive = true;
    while (alive) {
        alive = false;
        for (Warp *w : m.warps) {
            int live = 0, waiting = 0;
            for (auto &l : w->lanes) {
                if (l.done) continue;
                if (!l.waiting) {               // run this lane until it yields (collective, sleep or exit)
                    m.cur = &l;
                    swapcontext(&m.sched, &l.ctx);
                    m.clock += 50;
                }
                if (l.done) continue;
                ++live;
                if (l.waiting) ++waiting;
            }
            if (live) alive = true;
            if (live && waiting == live) resolve(w);
        }
    }
    for (Warp *w : m.warps) { for (auto &l : w->lanes) free(l.stack); delete w; }
    m.warps.clear();
    for (auto p : smems) free(p);
}

}  // namespace emu

#define threadIdx (emu::M().cur->tid)
#define blockIdx (emu::M().cur->bid)
#define blockDim (emu::M().block)
#define gridDim (emu::M().grid)
#define FMB_DYN_SMEM(name) unsigned char *name = emu::M().cur->warp->smem

template <typename T> inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_unbits<T>(emu::collective(emu::C_SHFL, emu_bits(v), src)); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu_unbits<T>(emu::collective(emu::C_SHFL_UP, emu_bits(v), d)); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu_unbits<T>(emu::collective(emu::C_SHFL_DOWN, emu_bits(v), d)); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_unbits<T>(emu::collective(emu::C_SHFL_XOR, emu_bits(v), m)); }
inline unsigned __ballot_sync(unsigned, bool p) { return (unsigned)emu::collective(emu::C_BALLOT, p ? 1 : 0, 0); }
inline bool __any_sync(unsigned, bool p) { return emu::collective(emu::C_ANY, p ? 1 : 0, 0) != 0; }
inline void __syncwarp() { emu::collective(emu::C_SYNC, 0, 0); }
inline void __syncthreads() { emu::collective(emu::C_BLOCK, 0, 0); }
inline int __syncthreads_or(int p) { return emu::collective(emu::C_BLOCK, p ? 1 : 0, 0) != 0; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { emu::M().yielded_sleep = true; emu::M().clock += 100; emu::yield_to_sched(); }
inline long long clock64() { return (long long)(emu::M().clock += 1); }

inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicSub(int *p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline long long __double_as_longlong(double v) { long long b; memcpy(&b, &v, 8); return b; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline int atomicCAS(int *p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

template <typename T> inline T __ldcg(const T *p) { return *(const volatile T *)p; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline void __stcg(T *p, T v) { *(volatile T *)p = v; }

inline void cp_async16_cg_emu(void *d, const void *s) { memcpy(d, s, 16); }
namespace fmb { inline void cp_async16_cg(void *d, const void *s) { memcpy(d, s, 16); } inline void cp_async_wait_all() {} }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __longlong_as_double(long long v) { return emu_unbits<double>((uint64_t)v); }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }

struct int2 { int x, y; };
inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
inline int __double2hiint(double v) { return (int)(emu_bits(v) >> 32); }
inline int __double2loint(double v) { return (int)(emu_bits(v) & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { return emu_unbits<double>(((uint64_t)(unsigned)hi << 32) | (unsigned)lo); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
using std::max;
using std::min;
