// Cross-GPU exchange fused into the compute kernels: peer-memory stores over NVLink + per-kernel flag hand-shakes.
//
// Model.  Every rank holds the whole simulation state in its own HBM (symmetric heap, heap.h) and runs the SAME sequence
// of kernels; a "sharded" kernel only computes its rank's k-slab of the work and delivers what the others need by
// storing straight into their copies (ghost planes to the two k-neighbours, reduction partials and gathered slabs to
// everybody).  Ordering is a counter per rank, carried in device memory so that it survives CUDA-graph replay:
//
//   xch_enter   thread 0 of every CTA waits until every other rank has COMPLETED at least as many sharded kernels as this
//               rank has (their counters are stored into this rank's Link by their last CTAs), i.e. until everything this
//               kernel may read was delivered and nobody still reads what this kernel may overwrite remotely;
//   xch_leave   the last CTA to finish bumps this rank's counter and stores it into every peer's Link.
// Pushes (k_push_planes / k_push_rows) only leave: see the note above k_push_planes.
//
// One NVLink store per peer and kernel; no host involvement, no NCCL on the substep path, no separate all-reduce kernels:
// reduction partials are pushed by the producing kernel (cg.h part_store) and re-reduced in a fixed rank-major order by
// every CTA of the consumer, so all ranks see bit-identical scalars and stay in lock step.  Every wait has a timeout that
// sets Link::status; once set, every later sharded kernel returns immediately and the host reports the failure
// (xch_check) before anything is written back.
//
// The reference is single-threaded and has no exchange of any kind.
#pragma once
#include "grid.h"
#include "heap.h"

struct Link {
    unsigned long long arrive[FLIP_MAX_RANKS];   // arrive[src] = sharded kernels rank src has completed (stored by src)
    unsigned long long done;                     // this rank's own count
    long long timeout_cycles;
    unsigned int cta_done;                       // CTAs of the running kernel that have finished
    int status;                                  // != 0: a wait timed out
};

struct Xch {
    Link *local;
    Link *const *peers;   // device table [nranks]; peers[rank] == local
    int rank, nranks;     // nranks == 1: single GPU or replicas only - every call below is a no-op
    int nbr;              // 1: kernels that only consume ghost planes wait for their two k-neighbours instead of all ranks
};

#ifdef FLIP_CPU_EMU
#include <sched.h>
#define XCH_SPIN_PAUSE() sched_yield()
FLIP_D unsigned long long xch_ld_acquire(const volatile unsigned long long *p) { unsigned long long v = *p; __sync_synchronize(); return v; }
FLIP_D void xch_st_relaxed(volatile unsigned long long *p, unsigned long long v) { *p = v; }
FLIP_D void xch_fence() { __sync_synchronize(); }
#else
#define XCH_SPIN_PAUSE()
// System-scope primitives in PTX: __threadfence_system() is MEMBAR.SC.SYS + an L1 invalidate in every thread that calls
// it; the hand-shake only needs an acquire load of the flag (LDG.STRONG.SYS + CCTL.IVALL, no barrier) and ONE release
// fence (MEMBAR.ALL.SYS) per CTA that stored into peer memory.
FLIP_D unsigned long long xch_ld_acquire(const volatile unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
FLIP_D void xch_st_relaxed(volatile unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
FLIP_D void xch_fence() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
#endif

// all threads of the CTA; false = the exchange is broken (time-out somewhere): return without doing anything.
// nbr_only: wait for ranks rank-1 and rank+1 only.  Enough for a kernel that reads nothing remote but ghost planes: those
// are stored by the neighbours, and the ghost planes this rank's next push overwrites are read by the neighbours.  Kernels
// that read partials or gathered slabs (stored by everybody) wait for everybody; there are four of those per CG iteration,
// so no rank ever runs more than a V-cycle ahead of any other.
FLIP_D bool xch_enter(const Xch &X, bool nbr_only = false) {
    if (X.nranks == 1) return true;
    __shared__ int xch_ok_s;
    if (threadIdx.x == 0) {
        int ok = 1;
        volatile Link *L = X.local;
        if (L->status) ok = 0;
        else {
            const unsigned long long want = L->done;
            const long long t0 = clock64(), lim = L->timeout_cycles;
            for (int src = 0; src < X.nranks && ok; src++) {
                if (src == X.rank || (nbr_only && src != X.rank - 1 && src != X.rank + 1)) continue;
                while (xch_ld_acquire(&L->arrive[src]) < want) {   // acquire: what the peer stored before its flag is visible
                    if (clock64() - t0 > lim) { L->status = 1; ok = 0; break; }
                    XCH_SPIN_PAUSE();
                }
            }
        }
        xch_ok_s = ok;
    }
    __syncthreads();
    return xch_ok_s != 0;
}

// all threads of the CTA, at the very end.  remote_stores: this kernel stored into peer memory.  The block barrier orders
// every thread's stores before thread 0's system-scope fence (fences are cumulative), so one fence per CTA publishes them
// all - a fence in every thread costs microseconds per kernel.
FLIP_D void xch_leave(const Xch &X, bool remote_stores) {
    if (X.nranks == 1) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (remote_stores) xch_fence();       // release: this CTA's peer stores are performed before its arrival is counted
        unsigned int prev = atomicAdd(&X.local->cta_done, 1u);
        if (prev == gridDim.x - 1) {
            if (remote_stores) xch_fence();   // acquire the other CTAs' arrivals, release the flag stores below
            X.local->cta_done = 0;
            unsigned long long n = X.local->done + 1;
            X.local->done = n;
            for (int p = 0; p < X.nranks; p++)
                if (p != X.rank) xch_st_relaxed(&((volatile unsigned long long *)X.peers[p]->arrive)[X.rank], n);
        }
    }
}

// ---- slab cuts -----------------------------------------------------------------------------------------------
// cut[l][r] .. cut[l][r+1] = the k-planes of multigrid level l that rank r owns (level 0 = the simulation grid).
// Level-0 cuts are multiples of 4 planes at least 4 planes apart, so that the cuts of levels 1 and 2 are whole
// planes and a 2-plane ghost layer of level 1 never reaches past the neighbouring slab.
#define XCH_LEVELS 3
struct Cuts { int c[XCH_LEVELS][FLIP_MAX_RANKS + 1]; };

// liquid cells (phi < 0) per k-plane; one CTA per plane
static __global__ void __launch_bounds__(256) k_plane_liquid(Grid g, const float *__restrict__ phi, int *__restrict__ count) {
    __shared__ int ws[8];
    const int k = blockIdx.x;
    int n = 0;
    for (int t = threadIdx.x; t < g.ni * g.nj; t += 256) n += phi[gidx(g, t % g.ni, t / g.ni, k)] < 0.0f ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += ws[w]; count[k] = t; }
}

// one thread: balanced cuts from the per-plane counts.  Planes 0 .. nk (nk + 1 of them: W faces live on plane nk).
static __global__ void k_make_cuts(Grid g, const int *__restrict__ count, int nranks, Cuts *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int np = g.nk + 1;
    long long total = 0;
    for (int k = 0; k < g.nk; k++) total += count[k];
    int c[FLIP_MAX_RANKS + 1];
    c[0] = 0;
    long long run = 0;
    int k = 0;
    for (int r = 1; r < nranks; r++) {
        long long target = total * r / nranks;
        while (k < g.nk && run + count[k] <= target) { run += count[k]; k++; }
        int cut = (k + 2) & ~3;                               // nearest multiple of 4
        int lo = c[r - 1] + 4, hi = ((np - 4 * (nranks - r)) & ~3);
        if (cut < lo) cut = lo;
        if (cut > hi) cut = hi;
        c[r] = cut;
    }
    for (int r = 0; r < nranks; r++)
        for (int l = 0; l < XCH_LEVELS; l++) out->c[l][r] = c[r] >> l;
    int n = g.nk;
    for (int l = 0; l < XCH_LEVELS; l++) { out->c[l][nranks] = n + 1; n = (n + 1) / 2; }
    for (int r = nranks + 1; r <= FLIP_MAX_RANKS; r++)
        for (int l = 0; l < XCH_LEVELS; l++) out->c[l][r] = out->c[l][nranks];
}

// ---- plane pushes --------------------------------------------------------------------------------------------
// Copies whole k-planes of a dense padded field (k is the slowest index: a plane range is one contiguous run per
// component) from this rank's slab into the same addresses of other ranks' copies:
//   halo > 0    the first `halo` planes of the slab -> the lower neighbour, the last `halo` planes -> the upper one
//               (their ghost layers);
//   halo == 0   the whole slab -> every other rank (slab all-gather).
struct PushDesc {
    const char *src;                 // local field
    char *dst[FLIP_MAX_RANKS];       // the same field on every rank (dst[rank] unused)
    size_t plane_bytes;              // ax * ay * sizeof(element), a multiple of 16
    size_t comp_stride_bytes;        // total * sizeof(element)
    int ncomp;
    int level;                       // which row of Cuts
    int halo;
    int pz;                          // ghost planes below plane 0 in the padded layout (FLIP_PZ)
};

// A push does not wait (no xch_enter): it follows, in stream order, the kernel that produced the planes, and no kernel
// reads the ghost planes of the field it produces, so by the time the producer could start (its own xch_enter) every
// rank had finished the last reader of the old ghost values.  The consumers' xch_enter waits for the pushes to land.
static __global__ void __launch_bounds__(256) k_push_planes(Xch X, const Cuts *__restrict__ cuts, PushDesc d) {
    if (((volatile Link *)X.local)->status) return;
    const int c0 = cuts->c[d.level][X.rank], c1 = cuts->c[d.level][X.rank + 1];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int t = 0; t < X.nranks; t++) {
        if (t == X.rank) continue;
        int p0, p1;
        if (d.halo > 0) {
            if (t == X.rank - 1) { p0 = c0; p1 = min(c0 + d.halo, c1); }
            else if (t == X.rank + 1) { p1 = c1; p0 = max(c1 - d.halo, c0); }
            else continue;
        } else { p0 = c0; p1 = c1; }
        if (p1 <= p0) continue;
        const size_t off = (size_t)(p0 + d.pz) * d.plane_bytes, n16 = (size_t)(p1 - p0) * d.plane_bytes / 16;
        for (int c = 0; c < d.ncomp; c++) {
            const int4 *__restrict__ s = (const int4 *)(d.src + c * d.comp_stride_bytes + off);
            int4 *__restrict__ q = (int4 *)(d.dst[t] + c * d.comp_stride_bytes + off);
            for (size_t e = tid; e < n16; e += nth) q[e] = s[e];
        }
    }
    xch_leave(X, true);
}

// Pushes a contiguous range of ROWS (explicit multigrid levels: coefficient rows, smoothing weights) of this rank to every
// other rank.  The range is read from device memory: rng[0] .. rng[1] in units of `row_bytes` (a multiple of 16).
static __global__ void __launch_bounds__(256) k_push_rows(Xch X, const int *__restrict__ rng, const char *src, PushDesc d, size_t row_bytes) {
    if (((volatile Link *)X.local)->status) return;
    const size_t off = (size_t)rng[0] * row_bytes, n16 = (size_t)(rng[1] - rng[0]) * row_bytes / 16;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const int4 *__restrict__ s = (const int4 *)(src + off);
    for (int t = 0; t < X.nranks; t++) {
        if (t == X.rank) continue;
        int4 *__restrict__ q = (int4 *)(d.dst[t] + off);
        for (size_t e = tid; e < n16; e += nth) q[e] = s[e];
    }
    xch_leave(X, true);
}
