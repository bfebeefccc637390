/*
 * DEV/TEST TOOLING ONLY — not a product path, not a fallback.
 *
 * A tiny single-host-thread CUDA emulator: enough of the runtime API and the device
 * built-ins to compile the flipviscosity3d_b200/csrc sources with g++ (-DFLIP_CPU_EMU) into
 * tests/cpu_emu/libflip_emu.so, so that kernel LOGIC can be checked against the
 * oracle in the no-GPU container (pytest -m "not gpu").  The shipped package never
 * loads this library: flipviscosity3d_b200 loads only lib/libflip_b200.so (nvcc,
 * sm_100a) and raises if it is missing.
 *
 * Blocks run one after another; threads of a block are ucontext fibers when the
 * kernel is launched with FLIP_LAUNCH_SYNC (it uses __syncthreads / warp shuffles),
 * plain loop iterations otherwise.
 */
#ifndef FLIP_CUDA_EMU_H
#define FLIP_CUDA_EMU_H

#include <ucontext.h>
#include <setjmp.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }

typedef int cudaError_t;
typedef void *cudaStream_t;
struct EmuEvent { double t; };
typedef EmuEvent *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

namespace emu {

struct Fiber {
    ucontext_t ctx;
    jmp_buf jb;      // fast resume point (no signal-mask syscall, unlike swapcontext)
    char *stack;
    bool done;
    bool started;
    unsigned tid;
};

struct State {
    dim3 tIdx, bIdx, bDim, gDim;
    bool fiberMode = false;
    ucontext_t sched;
    jmp_buf schedJb;
    std::vector<Fiber> fibers;
    Fiber *cur = nullptr;
    std::function<void()> body;
    // block barrier
    unsigned live = 0, barArrived = 0; unsigned long long barGen = 0;
    // warp barriers / exchange
    unsigned warpLive[64], warpArrived[64]; unsigned long long warpGen[64];
    unsigned long long xchg[64][32];
    unsigned ballot[64];
};

inline State &S() { static thread_local State s; return s; }
// ranks of a multi-rank test are host threads: their kernels hand-shake through __syncthreads-using helpers (xch.h)
inline bool &force_sync() { static thread_local bool f = false; return f; }

inline void yield() {
    State &s = S();
    if (!_setjmp(s.cur->jb)) _longjmp(s.schedJb, 1);
}

inline void checkBlockRelease() {
    State &s = S();
    if (s.live > 0 && s.barArrived >= s.live) { s.barArrived = 0; s.barGen++; }
}
inline void checkWarpRelease(unsigned w) {
    State &s = S();
    if (s.warpLive[w] > 0 && s.warpArrived[w] >= s.warpLive[w]) { s.warpArrived[w] = 0; s.warpGen[w]++; }
}

inline void blockBarrier() {
    State &s = S();
    if (!s.fiberMode) { fprintf(stderr, "cuda_emu: __syncthreads in a kernel launched without FLIP_LAUNCH_SYNC\n"); abort(); }
    unsigned long long g = s.barGen;
    s.barArrived++;
    checkBlockRelease();
    while (s.barGen == g) yield();
}
inline void warpBarrier() {
    State &s = S();
    if (!s.fiberMode) { fprintf(stderr, "cuda_emu: warp sync in a kernel launched without FLIP_LAUNCH_SYNC\n"); abort(); }
    unsigned w = s.cur->tid / 32;
    unsigned long long g = s.warpGen[w];
    s.warpArrived[w]++;
    checkWarpRelease(w);
    while (s.warpGen[w] == g) yield();
}

inline void fiberEntry() {
    State &s = S();
    s.body();
    Fiber *f = s.cur;
    f->done = true;
    s.live--;
    s.warpLive[f->tid / 32]--;
    checkBlockRelease();
    checkWarpRelease(f->tid / 32);
    _longjmp(s.schedJb, 1);
}

template <class F>
void launch(unsigned grid, unsigned block, bool sync, F f) {
    State &s = S();
    s.gDim = dim3(grid); s.bDim = dim3(block);
    if (force_sync()) sync = true;
    if (!sync) {
        s.fiberMode = false;
        for (unsigned b = 0; b < grid; b++) {
            s.bIdx = dim3(b);
            for (unsigned t = 0; t < block; t++) { s.tIdx = dim3(t); f(); }
        }
        return;
    }
    s.fiberMode = true;
    s.body = f;
    const size_t STK = 64 * 1024;
    if (s.fibers.size() < block) {
        size_t old = s.fibers.size();
        s.fibers.resize(block);
        for (size_t i = old; i < block; i++) s.fibers[i].stack = (char *)malloc(STK);
    }
    for (unsigned b = 0; b < grid; b++) {
        s.bIdx = dim3(b);
        s.live = block; s.barArrived = 0;
        for (unsigned w = 0; w < 64; w++) { s.warpLive[w] = 0; s.warpArrived[w] = 0; }
        for (unsigned t = 0; t < block; t++) {
            Fiber &fb = s.fibers[t];
            fb.done = false; fb.started = false; fb.tid = t;
            s.warpLive[t / 32]++;
            getcontext(&fb.ctx);
            fb.ctx.uc_stack.ss_sp = fb.stack; fb.ctx.uc_stack.ss_size = STK; fb.ctx.uc_link = nullptr;
            makecontext(&fb.ctx, (void (*)())fiberEntry, 0);
        }
        unsigned remaining = block;
        while (remaining) {
            remaining = 0;
            for (unsigned t = 0; t < block; t++) {
                Fiber &fb = s.fibers[t];
                if (fb.done) continue;
                s.cur = &fb; s.tIdx = dim3(t);
                if (!_setjmp(s.schedJb)) {
                    if (!fb.started) { fb.started = true; setcontext(&fb.ctx); }
                    else _longjmp(fb.jb, 1);
                }
                if (!fb.done) remaining++;
            }
        }
    }
    s.fiberMode = false;
}

template <class T> inline unsigned long long toBits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T fromBits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T shflIdx(T v, unsigned src) {
    State &s = S();
    unsigned w = s.cur->tid / 32, l = s.cur->tid % 32;
    s.xchg[w][l] = toBits(v);
    warpBarrier();
    T r = fromBits<T>(s.xchg[w][src % 32]);
    warpBarrier();
    return r;
}

}  // namespace emu

#define threadIdx (emu::S().tIdx)
#define blockIdx (emu::S().bIdx)
#define blockDim (emu::S().bDim)
#define gridDim (emu::S().gDim)

using std::min;
using std::max;

inline void __syncthreads() { emu::blockBarrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warpBarrier(); }
inline void __threadfence() {}
inline void __threadfence_system() { __sync_synchronize(); }
inline long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shflIdx(v, (unsigned)src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shflIdx(v, (emu::S().cur->tid % 32) ^ (unsigned)m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    unsigned l = emu::S().cur->tid % 32; return emu::shflIdx(v, l + d < 32 ? l + d : l);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    unsigned l = emu::S().cur->tid % 32; return emu::shflIdx(v, l >= d ? l - d : l);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::State &s = emu::S();
    unsigned w = s.cur->tid / 32, l = s.cur->tid % 32;
    s.xchg[w][l] = pred ? 1 : 0;
    emu::warpBarrier();
    unsigned r = 0;
    unsigned n = std::min(32u, s.bDim.x - w * 32);
    for (unsigned i = 0; i < n; i++) if (s.xchg[w][i]) r |= 1u << i;
    emu::warpBarrier();
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T> inline T __ldg(const T *p) { return *p; }

template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
inline int atomicCAS(int *p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }

/* ---- runtime API subset ---- */
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = n ? aligned_alloc(256, (n + 255) / 256 * 256) : nullptr; return (*p || !n) ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) {
    e->t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new EmuEvent(); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((b->t - a->t) * 1e3); return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
// "IPC" inside one process: the handle carries the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 1; strcpy(p->name, "cpu-emu"); return cudaSuccess; }

#endif
