// Peer-memory collectives for the slab-decomposed CG: the two exchanges that happen EVERY CG
// iteration — the scalar reductions and the one-plane halo of the search direction — done by small
// kernels that store straight into the other GPUs' memory over NVLink (CUDA IPC mappings) and
// synchronise with sequence-numbered flags, instead of NCCL calls.
//
// Why: at 2 M unknowns a CG iteration is ~70 us of work on one B200; each NCCL scalar all-reduce or
// send/recv pair costs 20-30 us, and an iteration needs three reductions and a halo exchange, so the
// NCCL version of the multi-GPU solve is slower than one GPU.  A remote store plus a
// flag costs one NVLink round trip (a few us), the kernels need no host involvement and are captured
// into the same CUDA graph as the iteration.
//
//   k_p2p_allreduce   1 CTA: reduce this rank's G partials in fixed order, store the value into
//                     slot [seq&3][rank] of EVERY rank's mailbox, fence, raise the flags; then wait for
//                     all ranks' flags in the local mailbox and combine in rank order (deterministic,
//                     identical on all ranks); the result replaces the partials (part[0], rest 0).
//   k_p2p_halo        copies the first / last owned plane of the field into the k-neighbours' ghost
//                     planes through peer pointers; the last CTA to finish raises the neighbours'
//                     flags and waits for theirs, so the stencil kernel that follows in the stream sees
//                     complete ghost planes.
// Every spin has a clock64 timeout that sets a status word instead of hanging the GPU.
// NCCL stays in use for set-up and for the one slab all-gather at the end of a solve.
#include "sim.h"

#ifndef FLIP_CPU_EMU
#include <cstring>

#define P2P_MAXR 16
#define P2P_SLOTS 4

struct Mailbox {
    double val[P2P_SLOTS][P2P_MAXR];
    unsigned long long flag[P2P_SLOTS][P2P_MAXR];
    unsigned long long hflag[2][P2P_SLOTS];   // [0] written by the lower neighbour, [1] by the upper one
    unsigned long long seq_reduce;            // local call counters (advance in lock step on all ranks)
    unsigned long long seq_halo;
    unsigned int halo_done;                   // CTAs of the current halo kernel that finished their copies
    int status;                               // != 0: a wait timed out
};

struct P2PState {
    Mailbox *local = nullptr;
    Mailbox *peer[P2P_MAXR] = {nullptr};      // peer[r] = rank r's mailbox mapped here (peer[rank] = local)
    double *field_peer[2][P2P_MAXR] = {{nullptr}};   // [0] cg_s, [1] cg_z of every rank
    Mailbox **peer_dev = nullptr;             // device copy of peer[]
    bool active = false;
};

struct P2PBlob {   // what one rank exports
    cudaIpcMemHandle_t mailbox, f0, f1;
    int rank;
    int pad;
};

static P2PState *p2p_of(Sim &s) { return (P2PState *)s.p2p; }

int dist_p2p_blob_size() { return (int)sizeof(P2PBlob); }

void dist_p2p_export(Sim &s, void *out) {
    if (!s.p2p) {
        P2PState *st = new P2PState();
        CUDA_CHECK(cudaMalloc((void **)&st->local, sizeof(Mailbox)));
        CUDA_CHECK(cudaMemset(st->local, 0, sizeof(Mailbox)));
        s.p2p = st;
    }
    P2PState *st = p2p_of(s);
    P2PBlob b;
    memset(&b, 0, sizeof(b));
    CUDA_CHECK(cudaIpcGetMemHandle(&b.mailbox, st->local));
    CUDA_CHECK(cudaIpcGetMemHandle(&b.f0, s.cg_s));
    CUDA_CHECK(cudaIpcGetMemHandle(&b.f1, s.cg_z));
    b.rank = s.rank;
    memcpy(out, &b, sizeof(b));
}

void dist_p2p_import(Sim &s, const void *all_blobs) {
    if (s.nranks < 2) return;
    if (s.nranks > P2P_MAXR) throw FlipError("p2p: too many ranks");
    P2PState *st = p2p_of(s);
    if (!st) throw FlipError("p2p: export must be called before import");
    const P2PBlob *blobs = (const P2PBlob *)all_blobs;
    for (int r = 0; r < s.nranks; r++) {
        if (blobs[r].rank != r) throw FlipError("p2p: blobs are not in rank order");
        if (r == s.rank) {
            st->peer[r] = st->local;
            st->field_peer[0][r] = s.cg_s;
            st->field_peer[1][r] = s.cg_z;
            continue;
        }
        void *p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, blobs[r].mailbox, cudaIpcMemLazyEnablePeerAccess));
        st->peer[r] = (Mailbox *)p;
        // only the k-neighbours' fields are ever written
        if (r == s.rank - 1 || r == s.rank + 1) {
            CUDA_CHECK(cudaIpcOpenMemHandle(&p, blobs[r].f0, cudaIpcMemLazyEnablePeerAccess));
            st->field_peer[0][r] = (double *)p;
            CUDA_CHECK(cudaIpcOpenMemHandle(&p, blobs[r].f1, cudaIpcMemLazyEnablePeerAccess));
            st->field_peer[1][r] = (double *)p;
        }
    }
    CUDA_CHECK(cudaMalloc((void **)&st->peer_dev, sizeof(Mailbox *) * P2P_MAXR));
    CUDA_CHECK(cudaMemcpy(st->peer_dev, st->peer, sizeof(Mailbox *) * P2P_MAXR, cudaMemcpyHostToDevice));
    st->active = true;
}

void dist_p2p_shutdown(Sim &s) {
    P2PState *st = p2p_of(s);
    if (!st) return;
    for (int r = 0; r < P2P_MAXR; r++) {
        if (r == s.rank) continue;
        if (st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
        for (int f = 0; f < 2; f++) if (st->field_peer[f][r]) cudaIpcCloseMemHandle(st->field_peer[f][r]);
    }
    if (st->peer_dev) cudaFree(st->peer_dev);
    if (st->local) cudaFree(st->local);
    delete st;
    s.p2p = nullptr;
}

bool dist_p2p_active(Sim &s) { return s.p2p && p2p_of(s)->active && s.nranks > 1; }

int dist_p2p_status(Sim &s) {
    if (!dist_p2p_active(s)) return 0;
    int v = 0;
    cudaMemcpy(&v, &p2p_of(s)->local->status, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

#define P2P_TIMEOUT_CYCLES 6000000000LL   // ~3 s at 2 GHz

__device__ __forceinline__ bool p2p_wait(volatile unsigned long long *flag, unsigned long long want, int *status) {
    long long t0 = clock64();
    while (*flag < want) {
        if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { *status = 1; return false; }
    }
    return true;
}

__global__ void __launch_bounds__(512) k_p2p_allreduce(double *__restrict__ part, int n, int is_max, Mailbox *local,
                                                       Mailbox **peers, int rank, int nranks) {
    __shared__ double sm[16];
    double v = 0.0;
    for (int q = threadIdx.x; q < n; q += 512) v = is_max ? fmax(v, part[q]) : v + part[q];
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, u) : v + u;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = sm[0];
        for (int w = 1; w < 16; w++) r = is_max ? fmax(r, sm[w]) : r + sm[w];
        unsigned long long seq = local->seq_reduce + 1;
        local->seq_reduce = seq;
        int slot = (int)(seq & (P2P_SLOTS - 1));
        for (int p = 0; p < nranks; p++) ((volatile double *)peers[p]->val[slot])[rank] = r;
        __threadfence_system();
        for (int p = 0; p < nranks; p++) ((volatile unsigned long long *)peers[p]->flag[slot])[rank] = seq;
        double total = 0.0;
        for (int src = 0; src < nranks; src++) {
            p2p_wait(&((volatile unsigned long long *)local->flag[slot])[src], seq, &local->status);
            __threadfence_system();
            double x = ((volatile double *)local->val[slot])[src];
            total = is_max ? fmax(total, x) : total + x;
        }
        sm[0] = total;
    }
    __syncthreads();
    double total = sm[0];
    for (int q = threadIdx.x; q < n; q += 512) part[q] = (q == 0) ? total : 0.0;
}

// grid: any number of CTAs; copies `ncomp` planes of `pe` doubles each way
__global__ void __launch_bounds__(256) k_p2p_halo(const double *__restrict__ field, double *lo_peer, double *hi_peer, int ncomp,
                                                  size_t total, size_t pe, size_t off_first, size_t off_last, Mailbox *local,
                                                  Mailbox *lo_box, Mailbox *hi_box) {
    size_t n = (size_t)ncomp * pe;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        size_t c = t / pe, e = t % pe;
        if (lo_peer) lo_peer[c * total + off_first + e] = field[c * total + off_first + e];   // my first plane = their ghost above
        if (hi_peer) hi_peer[c * total + off_last + e] = field[c * total + off_last + e];     // my last plane  = their ghost below
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(&local->halo_done, 1u);
        if (done == gridDim.x - 1) {
            local->halo_done = 0;
            unsigned long long seq = local->seq_halo + 1;
            local->seq_halo = seq;
            int slot = (int)(seq & (P2P_SLOTS - 1));
            __threadfence_system();
            // I am the UPPER neighbour of lo_box's owner and the LOWER neighbour of hi_box's owner
            if (lo_box) ((volatile unsigned long long *)lo_box->hflag[1])[slot] = seq;
            if (hi_box) ((volatile unsigned long long *)hi_box->hflag[0])[slot] = seq;
            if (lo_box) p2p_wait(&((volatile unsigned long long *)local->hflag[0])[slot], seq, &local->status);
            if (hi_box) p2p_wait(&((volatile unsigned long long *)local->hflag[1])[slot], seq, &local->status);
            __threadfence_system();
        }
    }
}


// Fused exchange after a kernel that produced a sum partial, a max partial and new values of a
// halo-exchanged field: halo copy by all CTAs; the last CTA reduces both partial arrays, posts them,
// raises the halo flags and waits for everything.  One kernel instead of three.
__global__ void __launch_bounds__(256) k_p2p_step(double *__restrict__ part_sum, double *__restrict__ part_max, int n,
                                                  const double *__restrict__ field, double *lo_peer, double *hi_peer, int ncomp,
                                                  size_t total, size_t pe, size_t off_first, size_t off_last, Mailbox *local,
                                                  Mailbox **peers, Mailbox *lo_box, Mailbox *hi_box, int rank, int nranks) {
    __shared__ double sm[2][8];
    __shared__ int is_last;
    if (field) {
        size_t cnt = (size_t)ncomp * pe;
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += (size_t)gridDim.x * blockDim.x) {
            size_t c = t / pe, e = t % pe;
            if (lo_peer) lo_peer[c * total + off_first + e] = field[c * total + off_first + e];
            if (hi_peer) hi_peer[c * total + off_last + e] = field[c * total + off_last + e];
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(&local->halo_done, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    // ---- last CTA ----
    double vs = 0.0, vm = 0.0;
    for (int q = threadIdx.x; q < n; q += 256) { vs += part_sum[q]; vm = fmax(vm, part_max[q]); }
    for (int o = 16; o > 0; o >>= 1) {
        vs += __shfl_xor_sync(0xffffffffu, vs, o);
        vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, o));
    }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = vs; sm[1][threadIdx.x >> 5] = vm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        local->halo_done = 0;
        double rs = sm[0][0], rm = sm[1][0];
        for (int w = 1; w < 8; w++) { rs += sm[0][w]; rm = fmax(rm, sm[1][w]); }
        // two consecutive reduction sequence numbers: [sum, max]
        unsigned long long seq = local->seq_reduce + 2;
        local->seq_reduce = seq;
        int s0 = (int)((seq - 1) & (P2P_SLOTS - 1)), s1 = (int)(seq & (P2P_SLOTS - 1));
        for (int p = 0; p < nranks; p++) {
            ((volatile double *)peers[p]->val[s0])[rank] = rs;
            ((volatile double *)peers[p]->val[s1])[rank] = rm;
        }
        unsigned long long hseq = 0;
        int hslot = 0;
        if (field) { hseq = local->seq_halo + 1; local->seq_halo = hseq; hslot = (int)(hseq & (P2P_SLOTS - 1)); }
        __threadfence_system();
        for (int p = 0; p < nranks; p++) {
            ((volatile unsigned long long *)peers[p]->flag[s0])[rank] = seq - 1;
            ((volatile unsigned long long *)peers[p]->flag[s1])[rank] = seq;
        }
        if (field) {
            if (lo_box) ((volatile unsigned long long *)lo_box->hflag[1])[hslot] = hseq;
            if (hi_box) ((volatile unsigned long long *)hi_box->hflag[0])[hslot] = hseq;
        }
        double ts = 0.0, tm = 0.0;
        for (int src = 0; src < nranks; src++) {
            p2p_wait(&((volatile unsigned long long *)local->flag[s1])[src], seq, &local->status);   // s1 is posted after s0
            __threadfence_system();
            ts += ((volatile double *)local->val[s0])[src];
            tm = fmax(tm, ((volatile double *)local->val[s1])[src]);
        }
        if (field) {
            if (lo_box) p2p_wait(&((volatile unsigned long long *)local->hflag[0])[hslot], hseq, &local->status);
            if (hi_box) p2p_wait(&((volatile unsigned long long *)local->hflag[1])[hslot], hseq, &local->status);
            __threadfence_system();
        }
        sm[0][0] = ts; sm[1][0] = tm;
    }
    __syncthreads();
    double ts = sm[0][0], tm = sm[1][0];
    for (int q = threadIdx.x; q < n; q += 256) { part_sum[q] = (q == 0) ? ts : 0.0; part_max[q] = (q == 0) ? tm : 0.0; }
}

void dist_p2p_step(Sim &s, double *part_sum, double *part_max, int n, double *field, int ncomp) {
    P2PState *st = p2p_of(s);
    const Grid &g = s.g;
    size_t pe = (size_t)g.ax * g.ay;
    int k0 = FLIP_B * s.bz0, k1 = FLIP_B * s.bz1;
    size_t off_first = (size_t)(k0 + FLIP_PZ) * pe, off_last = (size_t)(k1 - 1 + FLIP_PZ) * pe;
    double *lo = nullptr, *hi = nullptr;
    if (field) {
        int which = field == s.cg_s ? 0 : (field == s.cg_z ? 1 : -1);
        if (which < 0) throw FlipError("p2p step: field is not one of the exported arrays");
        lo = s.rank > 0 ? st->field_peer[which][s.rank - 1] : nullptr;
        hi = s.rank < s.nranks - 1 ? st->field_peer[which][s.rank + 1] : nullptr;
    }
    Mailbox *lob = s.rank > 0 ? st->peer[s.rank - 1] : nullptr;
    Mailbox *hib = s.rank < s.nranks - 1 ? st->peer[s.rank + 1] : nullptr;
    int grid = field ? s.num_sms : 1;
    k_p2p_step<<<grid, 256, 0, s.stream>>>(part_sum, part_max, n, (const double *)field, lo, hi, ncomp, (size_t)g.total, pe, off_first,
                                           off_last, st->local, st->peer_dev, lob, hib, s.rank, s.nranks);
    s.kernel_launches++;
}

void dist_p2p_reduce(Sim &s, double *part, int n, bool is_max) {
    P2PState *st = p2p_of(s);
    k_p2p_allreduce<<<1, 512, 0, s.stream>>>(part, n, is_max ? 1 : 0, st->local, st->peer_dev, s.rank, s.nranks);
    s.kernel_launches++;
}

void dist_p2p_halo(Sim &s, double *field, int ncomp) {
    P2PState *st = p2p_of(s);
    const Grid &g = s.g;
    int which = field == s.cg_s ? 0 : (field == s.cg_z ? 1 : -1);
    if (which < 0) throw FlipError("p2p halo: field is not one of the exported arrays");
    size_t pe = (size_t)g.ax * g.ay;
    int k0 = FLIP_B * s.bz0, k1 = FLIP_B * s.bz1;
    size_t off_first = (size_t)(k0 + FLIP_PZ) * pe, off_last = (size_t)(k1 - 1 + FLIP_PZ) * pe;
    double *lo = s.rank > 0 ? st->field_peer[which][s.rank - 1] : nullptr;
    double *hi = s.rank < s.nranks - 1 ? st->field_peer[which][s.rank + 1] : nullptr;
    Mailbox *lob = s.rank > 0 ? st->peer[s.rank - 1] : nullptr;
    Mailbox *hib = s.rank < s.nranks - 1 ? st->peer[s.rank + 1] : nullptr;
    int grid = s.num_sms;
    k_p2p_halo<<<grid, 256, 0, s.stream>>>((const double *)field, lo, hi, ncomp, (size_t)g.total, pe, off_first, off_last, st->local,
                                           lob, hib);
    s.kernel_launches++;
}
#else
int dist_p2p_blob_size() { return 0; }
void dist_p2p_export(Sim &, void *) { throw FlipError("cpu-emu build has no peer memory"); }
void dist_p2p_import(Sim &, const void *) {}
void dist_p2p_shutdown(Sim &) {}
bool dist_p2p_active(Sim &) { return false; }
int dist_p2p_status(Sim &) { return 0; }
void dist_p2p_reduce(Sim &, double *, int, bool) {}
void dist_p2p_halo(Sim &, double *, int) {}
void dist_p2p_step(Sim &, double *, double *, int, double *, int) {}
#endif
