// Multi-GPU support: one process per GPU, 1-D slab decomposition along k (the slowest index, so a
// slab and its ghost planes are contiguous in memory), NCCL over NVLink for the exchanges.
//
// What is decomposed: the two CG solves, which are > 90 % of a substep — each rank applies the
// stencil and the vector updates on its own slab of 8x8x8 blocks, exchanges ONE ghost plane of the
// search direction with each k-neighbour per stencil apply (three planes, U/V/W, for the viscosity
// system) and all-reduces the CG scalars.  What is replicated: the particle and grid stages
// (binning, SDF, P2G, extrapolation, G2P; ~2 % of a 256^3 substep) run identically on every rank —
// they are deterministic, so the replicas stay bit-identical — and each solve ends with an
// all-gather of the solution slabs.  SURVEY.md §8(e) describes the fully partitioned variant
// (particle migration, halo-min/halo-add); that is the next step, see DESIGN.md.
//
// The reference has no multi-process path at all, so nothing here replaces
// reference code.
#include "sim.h"

#ifdef FLIP_CPU_EMU
// the CPU emulator is single-process
void dist_setup_slab(Sim &s) { s.bz0 = 0; s.bz1 = s.g.nbz; }
void dist_init(Sim &, int, int nranks, const void *) { if (nranks != 1) throw FlipError("cpu-emu build is single process"); }
void dist_shutdown(Sim &) {}
void dist_get_unique_id(void *out128) { memset(out128, 0, 128); }
void dist_reduce_partials(Sim &, double *, int, bool) {}
void dist_halo_exchange(Sim &, double *, int) {}
void dist_allgather_slabs(Sim &, double *, int) {}
void dist_allreduce_int(Sim &, int *) {}
void dist_reduce_pair(Sim &, double *, double *, int, double *, int) {}
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

void dist_setup_slab(Sim &s) {
    int nbz = s.g.nbz;
    s.bz0 = (int)((long long)nbz * s.rank / s.nranks);
    s.bz1 = (int)((long long)nbz * (s.rank + 1) / s.nranks);
}

void dist_get_unique_id(void *out128) {
    ncclUniqueId id;
    NCCL_CHECK(ncclGetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
}

void dist_init(Sim &s, int rank, int nranks, const void *unique_id) {
    if (nranks < 1 || rank < 0 || rank >= nranks) throw FlipError("dist_init: bad rank / nranks");
    if (nranks > s.g.nbz) throw FlipError("dist_init: more ranks than 8-cell block layers along k");
    dist_shutdown(s);
    s.rank = rank; s.nranks = nranks;
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, unique_id, 128);
        ncclComm_t comm;
        NCCL_CHECK(ncclCommInitRank(&comm, nranks, id, rank));
        s.nccl = (void *)comm;
    }
    dist_setup_slab(s);
}

void dist_shutdown(Sim &s) {
    dist_p2p_shutdown(s);
    if (s.nccl) { ncclCommDestroy((ncclComm_t)s.nccl); s.nccl = 0; }
    s.rank = 0; s.nranks = 1;
    dist_setup_slab(s);
}

// part[0] = reduce(part[0..n)), part[1..n) = 0: consumers that re-reduce the n partials then see the
// global value once the all-reduce has run on part[0]
__global__ void __launch_bounds__(512) k_collapse_partials(double *__restrict__ part, int n, int is_max) {
    __shared__ double sm[16];
    double v = 0.0;
    for (int q = threadIdx.x; q < n; q += 512) v = is_max ? fmax(v, part[q]) : v + part[q];
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, u) : v + u;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < 16; w++) r = is_max ? fmax(r, sm[w]) : r + sm[w];
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += 512) part[q] = (q == 0) ? r : 0.0;
}

void dist_reduce_partials(Sim &s, double *part, int n, bool is_max) {
    if (s.nranks == 1) return;
    if (dist_p2p_active(s)) { dist_p2p_reduce(s, part, n, is_max); return; }
    k_collapse_partials<<<1, 512, 0, s.stream>>>(part, n, is_max ? 1 : 0);
    s.kernel_launches++;
    NCCL_CHECK(ncclAllReduce(part, part, 1, ncclDouble, is_max ? ncclMax : ncclSum, (ncclComm_t)s.nccl, s.stream));
}

void dist_reduce_pair(Sim &s, double *part_sum, double *part_max, int n, double *field, int ncomp) {
    if (s.nranks == 1) return;
    if (dist_p2p_active(s)) { dist_p2p_step(s, part_sum, part_max, n, field, ncomp); return; }
    dist_reduce_partials(s, part_sum, n, false);
    dist_reduce_partials(s, part_max, n, true);
    if (field) dist_halo_exchange(s, field, ncomp);
}

void dist_allreduce_int(Sim &s, int *dev_value) {
    if (s.nranks == 1) return;
    NCCL_CHECK(ncclAllReduce(dev_value, dev_value, 1, ncclInt, ncclSum, (ncclComm_t)s.nccl, s.stream));
}

static inline size_t plane_elems(const Grid &g) { return (size_t)g.ax * g.ay; }
static inline size_t plane_offset(const Grid &g, int k) { return (size_t)(k + FLIP_PZ) * plane_elems(g); }

// Owned cell planes of this rank: k in [8*bz0, min(8*bz1, nk+1)).  The plane just below / above the
// slab is the ghost layer the 7-point and the coupled-face stencils read.
void dist_halo_exchange(Sim &s, double *field, int ncomp) {
    if (s.nranks == 1) return;
    if (dist_p2p_active(s)) { dist_p2p_halo(s, field, ncomp); return; }
    const Grid &g = s.g;
    ncclComm_t comm = (ncclComm_t)s.nccl;
    int k0 = FLIP_B * s.bz0, k1 = FLIP_B * s.bz1;   // k1 may exceed nk+1 on the last rank (no upper neighbour then)
    size_t pe = plane_elems(g);
    NCCL_CHECK(ncclGroupStart());
    for (int c = 0; c < ncomp; c++) {
        double *f = field + (size_t)c * g.total;
        if (s.rank > 0) {
            NCCL_CHECK(ncclSend(f + plane_offset(g, k0), pe, ncclDouble, s.rank - 1, comm, s.stream));
            NCCL_CHECK(ncclRecv(f + plane_offset(g, k0 - 1), pe, ncclDouble, s.rank - 1, comm, s.stream));
        }
        if (s.rank < s.nranks - 1) {
            NCCL_CHECK(ncclSend(f + plane_offset(g, k1 - 1), pe, ncclDouble, s.rank + 1, comm, s.stream));
            NCCL_CHECK(ncclRecv(f + plane_offset(g, k1), pe, ncclDouble, s.rank + 1, comm, s.stream));
        }
    }
    NCCL_CHECK(ncclGroupEnd());
}

void dist_allgather_slabs(Sim &s, double *field, int ncomp) {
    if (s.nranks == 1) return;
    const Grid &g = s.g;
    ncclComm_t comm = (ncclComm_t)s.nccl;
    size_t pe = plane_elems(g);
    NCCL_CHECK(ncclGroupStart());
    for (int r = 0; r < s.nranks; r++) {
        int b0 = (int)((long long)g.nbz * r / s.nranks), b1 = (int)((long long)g.nbz * (r + 1) / s.nranks);
        int k0 = FLIP_B * b0, k1 = FLIP_B * b1;
        if (k1 > g.nk + 1) k1 = g.nk + 1;
        if (k1 <= k0) continue;
        for (int c = 0; c < ncomp; c++) {
            double *f = field + (size_t)c * g.total + plane_offset(g, k0);
            NCCL_CHECK(ncclBroadcast(f, f, (size_t)(k1 - k0) * pe, ncclDouble, r, comm, s.stream));
        }
    }
    NCCL_CHECK(ncclGroupEnd());
}
#endif
