// Multi-GPU support: one process per GPU; every rank holds the whole (bit-identical) simulation state in its own
// 180 GB and the WORK of a substep is cut into k-slabs (k is the slowest index, so a slab and its ghost planes are
// contiguous in memory).  Results travel by peer-memory stores over NVLink / NVSwitch straight into the other ranks'
// copies (symmetric heap, heap.h) with per-kernel flag hand-shakes (xch.h): no NCCL call, no host round trip and no
// packing anywhere on the substep path.  NCCL is only used here for the rendezvous (communicator creation doubles as a
// sanity check that all ranks can talk) — torch.distributed / the host application ships the unique id and the heap
// handles.
//
// What is cut into slabs (balanced by liquid cells per k-plane, re-cut every substep on the device):
//   both CG solves (stencil applies, vector updates, dot products), multigrid levels 0 and 1 of the viscosity
//   preconditioner (sweeps, residuals, transfers), the Galerkin products of levels 1 and 2.
// What every rank runs in full on its own copy: multigrid levels >= 2 (small, latency bound), the grid set-up sweeps and
// the particle stages (deterministic, so the replicas stay bit-identical).
//
// The reference has no multi-process path at all, so nothing here replaces reference code.
#include "sim.h"
#include <cstring>

static __global__ void k_xch_barrier(Xch X) {
    if (!xch_enter(X)) return;
    xch_leave(X, false);
}

struct HeapBlob {   // what one rank exports: its heap chunks
    int rank, nchunks;
    unsigned long long size[FLIP_HEAP_MAX_CHUNKS];
    cudaIpcMemHandle_t handle[FLIP_HEAP_MAX_CHUNKS];
};

int dist_p2p_blob_size() { return (int)sizeof(HeapBlob); }

void dist_p2p_export(Sim &s, void *out) {
    HeapBlob b;
    memset(&b, 0, sizeof(b));
    b.rank = s.rank;
    b.nchunks = (int)s.heap.chunks.size();
    for (int c = 0; c < b.nchunks; c++) {
        b.size[c] = s.heap.chunks[c].size;
        CUDA_CHECK(cudaIpcGetMemHandle(&b.handle[c], s.heap.chunks[c].base));
    }
    memcpy(out, &b, sizeof(b));
}

static void close_peers(Sim &s) {
    for (int r = 0; r < FLIP_MAX_RANKS; r++)
        for (int c = 0; c < FLIP_HEAP_MAX_CHUNKS; c++) {
            if (s.heap.peer_base[r][c] && r != s.rank) cudaIpcCloseMemHandle(s.heap.peer_base[r][c]);
            s.heap.peer_base[r][c] = nullptr;
        }
}

void dist_p2p_shutdown(Sim &s) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    close_peers(s);
    s.heap.frozen = false;
    s.sharded = false;
    s.xch_epoch++;   // captured graphs hold the hand-shakes of the old set-up
}

void dist_p2p_import(Sim &s, const void *all_blobs) {
    if (s.nranks < 2) return;
    if (s.nranks > FLIP_MAX_RANKS) throw FlipError("flip_dist_p2p_import: too many ranks");
    if (s.g.nk + 1 < 4 * s.nranks) throw FlipError("flip_dist_p2p_import: fewer than 4 k-planes per rank");
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    close_peers(s);
    s.sharded = false;
    const HeapBlob *blobs = (const HeapBlob *)all_blobs;
    const int nch = (int)s.heap.chunks.size();
    try {
        for (int r = 0; r < s.nranks; r++) {
            if (blobs[r].rank != r) throw FlipError("flip_dist_p2p_import: blobs are not in rank order");
            if (blobs[r].nchunks != nch) throw FlipError("flip_dist_p2p_import: the ranks' heaps differ (same scene and calls on every rank?)");
            for (int c = 0; c < nch; c++) {
                if (blobs[r].size[c] != s.heap.chunks[c].size) throw FlipError("flip_dist_p2p_import: the ranks' heap chunks differ in size");
                if (r == s.rank) { s.heap.peer_base[r][c] = s.heap.chunks[c].base; continue; }
                void *p = nullptr;
                CUDA_CHECK(cudaIpcOpenMemHandle(&p, blobs[r].handle[c], cudaIpcMemLazyEnablePeerAccess));
                s.heap.peer_base[r][c] = (char *)p;
            }
        }
    } catch (...) {
        close_peers(s);   // do not leave half-opened mappings behind
        throw;
    }
    s.heap.rank = s.rank; s.heap.nranks = s.nranks;
    s.heap.frozen = true;
    // hand-shake state: all counters equal (zero) on every rank, tables of the peers' Link / partial arrays
    Link L;
    memset(&L, 0, sizeof(L));
#ifdef FLIP_CPU_EMU
    L.timeout_cycles = (long long)(s.xch_timeout_s * 1e9);
#else
    int dev = 0, khz = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    L.timeout_cycles = (long long)(s.xch_timeout_s * 1e3 * (double)khz);
#endif
    CUDA_CHECK(cudaMemcpy(s.link, &L, sizeof(L), cudaMemcpyHostToDevice));
    Link *lp[FLIP_MAX_RANKS];
    double *pp[FLIP_MAX_RANKS];
    for (int r = 0; r < FLIP_MAX_RANKS; r++) {
        lp[r] = r < s.nranks ? s.heap.peer(r, s.link) : nullptr;
        pp[r] = r < s.nranks ? s.heap.peer(r, s.part) : nullptr;
    }
    CUDA_CHECK(cudaMemcpy(s.link_peers, lp, sizeof(lp), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(s.part_peers, pp, sizeof(pp), cudaMemcpyHostToDevice));
    s.sharded = true;
    s.xch_epoch++;
}

Xch xch_of(Sim &s) {
    Xch X;
    X.local = s.link; X.peers = s.link_peers;
    X.rank = s.sharded ? s.rank : 0;
    X.nranks = s.sharded ? s.nranks : 1;
    X.nbr = s.xch_nbr_wait;
    return X;
}
const Cuts *xch_cuts(Sim &s) { return s.sharded ? s.cuts : nullptr; }
int xch_rank(Sim &s) { return s.sharded ? s.rank : 0; }

void xch_update_cuts(Sim &s) {
    if (!s.sharded) return;
    FLIP_LAUNCH_SYNC(k_plane_liquid, s.g.nk, 256, s.stream, s.g, (const float *)s.phi_liq, s.plane_count);
    FLIP_LAUNCH(k_make_cuts, 1, 32, s.stream, s.g, (const int *)s.plane_count, s.nranks, s.cuts);
    s.kernel_launches += 2;
    KERNEL_CHECK();
}

static PushDesc push_desc(Sim &s, const Grid &g, void *field, size_t elem, int ncomp, int level, int halo) {
    PushDesc d;
    memset(&d, 0, sizeof(d));
    d.src = (const char *)field;
    for (int r = 0; r < s.nranks; r++) d.dst[r] = (char *)s.heap.peer(r, field);
    d.plane_bytes = (size_t)g.ax * g.ay * elem;
    d.comp_stride_bytes = (size_t)g.total * elem;
    d.ncomp = ncomp; d.level = level; d.halo = halo; d.pz = FLIP_PZ;
    return d;
}

// CTAs of a push: every CTA ends with one system-scope fence (~microseconds when hundreds of CTAs do it, measured), so
// a push uses few, fat CTAs: 32 KB per CTA, at most 64 of them for ghost planes and 4 per SM for bulk gathers.
static int push_grid(Sim &s, size_t bytes, bool bulk) {
    long long ctas = (long long)(bytes / (32 << 10)) + 1;
    long long cap = bulk ? (long long)s.num_sms * 4 : 64;
    return (int)(ctas < cap ? ctas : cap);
}

void xch_push_halo(Sim &s, const Grid &g, void *field, size_t elem, int ncomp, int level, int halo) {
    if (!s.sharded) return;
    PushDesc d = push_desc(s, g, field, elem, ncomp, level, halo);
    int G = push_grid(s, 2 * (size_t)halo * ncomp * d.plane_bytes, false);
    FLIP_LAUNCH_SYNC(k_push_planes, G, 256, s.stream, xch_of(s), (const Cuts *)s.cuts, d);
    s.kernel_launches++;
}

void xch_push_gather(Sim &s, const Grid &g, void *field, size_t elem, int ncomp, int level) {
    if (!s.sharded) return;
    PushDesc d = push_desc(s, g, field, elem, ncomp, level, 0);
    // a slab is at most the whole field; the usual one is 1/nranks of it
    int G = push_grid(s, (size_t)ncomp * d.comp_stride_bytes / s.nranks * (s.nranks - 1), true);
    FLIP_LAUNCH_SYNC(k_push_planes, G, 256, s.stream, xch_of(s), (const Cuts *)s.cuts, d);
    s.kernel_launches++;
}

void xch_push_rows(Sim &s, const int *rng_dev, void *base, size_t row_bytes) {
    if (!s.sharded) return;
    PushDesc d;
    memset(&d, 0, sizeof(d));
    d.src = (const char *)base;
    for (int r = 0; r < s.nranks; r++) d.dst[r] = (char *)s.heap.peer(r, base);
    FLIP_LAUNCH_SYNC(k_push_rows, s.num_sms * 4, 256, s.stream, xch_of(s), rng_dev, (const char *)base, d, row_bytes);
    s.kernel_launches++;
}

void xch_barrier(Sim &s) {
    if (!s.sharded) return;
    FLIP_LAUNCH_SYNC(k_xch_barrier, 1, 32, s.stream, xch_of(s));
    s.kernel_launches++;
}

void xch_status_fetch(Sim &s) {
    if (!s.sharded) return;
    CUDA_CHECK(cudaMemcpyAsync(s.xch_status_host, &s.link->status, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
}
bool xch_status_bad(Sim &s) { return s.sharded && *s.xch_status_host != 0; }

void xch_check(Sim &s) {
    if (!s.sharded) return;
    int status = 0;
    CUDA_CHECK(cudaMemcpy(&status, &s.link->status, sizeof(int), cudaMemcpyDeviceToHost));
    if (status != 0)
        throw FlipError("multi-GPU exchange timed out: a rank is missing or out of step (every rank must hold the same scene "
                        "and make the same calls); call flip_dist_p2p_export/_import again on all ranks to re-arm");
}

#ifdef FLIP_CPU_EMU
// the CPU emulator has no NCCL: ranks are host threads of one process (tests/test_sharded_emu.py)
void dist_get_unique_id(void *out128) { memset(out128, 0, 128); }
void dist_init(Sim &s, int rank, int nranks, const void *) {
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > FLIP_MAX_RANKS) throw FlipError("dist_init: bad rank / nranks");
    dist_shutdown(s);
    s.rank = rank; s.nranks = nranks;
    s.heap.rank = rank; s.heap.nranks = nranks;
}
void dist_shutdown(Sim &s) {
    dist_p2p_shutdown(s);
    s.rank = 0; s.nranks = 1;
    s.heap.rank = 0; s.heap.nranks = 1;
}
#else
#include <nccl.h>   // types only: the library is bound at run time (see NcclApi)
#include <dlfcn.h>
#include <cstring>

// NCCL is resolved with dlopen at flip_dist_init time instead of at link time.  A process that
// also uses PyTorch has torch's bundled libnccl.so.2 (2.28) mapped; linking this library against
// the system one (2.27) made whichever loaded first win for both, and torch then failed to find its
// newer symbols.  dlopen("libnccl.so.2") returns the copy that is already mapped, or the system
// one in a plain C++ host.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    if (api.handle) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw FlipError(std::string("cannot load libnccl.so.2: ") + dlerror());
#define FLIP_NCCL_SYM(field, name)                                             \
    api.field = (decltype(api.field))dlsym(h, name);                           \
    if (!api.field) throw FlipError(std::string("libnccl lacks ") + name);
    FLIP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FLIP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FLIP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FLIP_NCCL_SYM(AllReduce, "ncclAllReduce")
    FLIP_NCCL_SYM(Broadcast, "ncclBroadcast")
    FLIP_NCCL_SYM(Send, "ncclSend")
    FLIP_NCCL_SYM(Recv, "ncclRecv")
    FLIP_NCCL_SYM(GroupStart, "ncclGroupStart")
    FLIP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FLIP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FLIP_NCCL_SYM
    api.handle = h;
    return api;
}
#define ncclGetUniqueId nccl_api().GetUniqueId
#define ncclCommInitRank nccl_api().CommInitRank
#define ncclCommDestroy nccl_api().CommDestroy
#define ncclAllReduce nccl_api().AllReduce
#define ncclBroadcast nccl_api().Broadcast
#define ncclSend nccl_api().Send
#define ncclRecv nccl_api().Recv
#define ncclGroupStart nccl_api().GroupStart
#define ncclGroupEnd nccl_api().GroupEnd
#define ncclGetErrorString nccl_api().GetErrorString

#define NCCL_CHECK(expr)                                                                          \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess)                                                                    \
            throw FlipError(std::string(#expr) + " failed: " + ncclGetErrorString(_r));           \
    } while (0)

void dist_get_unique_id(void *out128) {
    ncclUniqueId id;
    NCCL_CHECK(ncclGetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
}

void dist_init(Sim &s, int rank, int nranks, const void *unique_id) {
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > FLIP_MAX_RANKS) throw FlipError("dist_init: bad rank / nranks");
    dist_shutdown(s);
    s.rank = rank; s.nranks = nranks;
    s.heap.rank = rank; s.heap.nranks = nranks;
    if (nranks > 1 && unique_id) {
        ncclUniqueId id;
        memcpy(&id, unique_id, 128);
        ncclComm_t comm;
        NCCL_CHECK(ncclCommInitRank(&comm, nranks, id, rank));
        s.nccl = (void *)comm;
        // one collective on the communicator: every rank is really there before anybody maps peer memory
        int *one = nullptr;
        CUDA_CHECK(cudaMalloc((void **)&one, sizeof(int)));
        CUDA_CHECK(cudaMemset(one, 0, sizeof(int)));
        NCCL_CHECK(ncclAllReduce(one, one, 1, ncclInt, ncclSum, comm, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        cudaFree(one);
    }
}

void dist_shutdown(Sim &s) {
    dist_p2p_shutdown(s);
    if (s.nccl) { ncclCommDestroy((ncclComm_t)s.nccl); s.nccl = 0; }
    s.rank = 0; s.nranks = 1;
    s.heap.rank = 0; s.heap.nranks = 1;
}
#endif

// ---- exchange micro-benchmark (dev tool, tests/gpu_dev_xch.py): what one synchronising kernel costs ----------------
#ifndef FLIP_CPU_EMU
static __global__ void __launch_bounds__(512) k_xb_plain(int dummy) {
    if (dummy == 12345 && threadIdx.x == 0) printf("never\n");
}
static __global__ void __launch_bounds__(512) k_xb_enter_leave(Xch X, int fence) {
    if (!xch_enter(X)) return;
    xch_leave(X, fence != 0);
}
static __global__ void __launch_bounds__(512) k_xb_leave_only(Xch X, int fence) { xch_leave(X, fence != 0); }
// every thread stores one float into the upper/lower neighbour's copy of `buf` (coalesced or strided), then leave
static __global__ void __launch_bounds__(512) k_xb_scatter(Xch X, float *peer_buf, int stride) {
    if (!xch_enter(X)) return;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    peer_buf[t * (size_t)stride] = (float)t;
    xch_leave(X, true);
}
extern "C" int flip_debug_xch_bench(void *hsim, int mode, int reps, int ctas, float *us_per_kernel) {
    Sim &s = *(Sim *)hsim;
    try {
        Xch X = xch_of(s);
        float *buf = s.vnode;
        float *peer = s.sharded ? s.heap.peer((s.rank + 1) % s.nranks, buf) : buf;
        PushDesc d;
        Grid g = s.g;
        auto launch = [&]() {
            switch (mode) {
                case 0: k_xb_plain<<<ctas, 512, 0, s.stream>>>(0); break;
                case 1: k_xb_enter_leave<<<ctas, 512, 0, s.stream>>>(X, 0); break;
                case 2: k_xb_enter_leave<<<ctas, 512, 0, s.stream>>>(X, 1); break;
                case 3: k_xb_leave_only<<<ctas, 512, 0, s.stream>>>(X, 0); break;
                case 4: k_xb_leave_only<<<ctas, 512, 0, s.stream>>>(X, 1); break;
                case 5: xch_push_halo(s, g, s.cg_s, sizeof(double), 3, 0, 1); break;      // fp64 search direction, 1 plane each way
                case 6: k_xb_scatter<<<ctas, 512, 0, s.stream>>>(X, peer, 1); break;       // coalesced 4-byte remote stores
                case 7: k_xb_scatter<<<ctas, 512, 0, s.stream>>>(X, peer, 8); break;       // one store per 32-byte sector
                case 8: xch_barrier(s); break;
                case 9: xch_push_halo(s, g, s.vnode, sizeof(float), 3, 0, 1); break;       // fp32 level-0 iterate, 1 plane each way
            }
        };
        xch_update_cuts(s);
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        CUDA_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        for (int r = 0; r < reps; r++) launch();
        CUDA_CHECK(cudaStreamEndCapture(s.stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_CHECK(cudaGraphLaunch(exec, s.stream));   // warm-up
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        CUDA_CHECK(cudaEventRecord(e0, s.stream));
        CUDA_CHECK(cudaGraphLaunch(exec, s.stream));
        CUDA_CHECK(cudaEventRecord(e1, s.stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        *us_per_kernel = ms * 1e3f / reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        xch_check(s);
    } catch (const std::exception &e) { s.last_error = e.what(); return -2; }
    return 0;
}
#endif
