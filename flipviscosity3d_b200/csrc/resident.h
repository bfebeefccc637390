// Persistent ("resident") kernels: the whole grid stays on the SMs for a complete iterative phase and synchronises
// through a grid-wide barrier instead of kernel boundaries.
//
// k_pcg_resident: the ENTIRE diagonal-preconditioned CG of the pressure projection in one launch.  The system of the
// headline scene (0.62 M unknowns at 256^3) is small enough that its per-cell solver state fits ON CHIP: every thread owns
// up to CPT cells of the compact cell list and keeps x, r, p, q of those cells in registers and the seven stencil
// coefficients of each in shared memory (<= 196 KB per CTA, one CTA per SM).  Per iteration only the search direction
// travels - one 8-byte store per cell, six neighbour loads served by L2 - and three grid-wide exchanges replace the launch
// boundaries and the host's convergence polls of the kernel-per-phase path (cg.h): the solve is latency bound (the
// working set of the launch-per-phase version already sat in the 126 MB L2), so what is removed is launch latency.
// Same operator, same preconditioner, same stopping rule (max|r| < tol before maxit, src/pressuresolver.cpp:521-567) and
// the same deterministic fixed-order reductions as run_cg; systems that do not fit (more than CPT_MAX * 512 cells per
// SM) or that are cut into slabs across GPUs take the launch-per-phase path.
#pragma once
#include "cg.h"

// ---- grid-wide exchange -------------------------------------------------------------------------------------------
// One synchronisation point = one two-hop flag exchange that is a barrier AND an all-reduce of up to two doubles:
//   every CTA stores its contribution into its own slot, each double split into two 8-byte words {32 data bits,
//   generation} (8-byte stores are single-copy atomic: a reader that sees the generation also sees the data, so pure
//   reductions need no fence and no atomics);  CTA 0 polls all slots, combines them in a fixed order (bit-reproducible)
//   and publishes the result the same way;  one thread of every CTA polls the result.
// Measured on B200 (dev/grid_sync_bench.cu, 148 CTAs): 1.8 us per exchange, against 2.6-4.7 us for all-to-all polling
// (every CTA reading every slot - and in the real kernels the polling traffic of the early CTAs slowed the late ones
// down several times more) and 4.3 us for an atomic-counter barrier with fences plus a read-back of the partials.
// Slots and result words are double buffered by the parity of the generation: a CTA can only run one exchange ahead of
// CTA 0 and vice versa.  VIS: the exchange also orders every CTA's earlier global stores before every CTA's later loads
// (fences in the publishing / polling threads; consumers read such data with L2 loads, ld_cg).
// All CTAs must be co-resident (one CTA per SM, grid <= SMs).  A wait that exceeds ~2 s sets GridBar::broken and every
// later exchange returns at once with ok = false: a lost CTA cannot hang the device.
#define GRID_MAX_CTAS 1024
#define GRID_SLOT_WORDS 4
struct GridCtx {
    unsigned long long *slots;   // [2][GRID_MAX_CTAS + 1][GRID_SLOT_WORDS]; entry GRID_MAX_CTAS of a buffer = the result
    GridBar *bar;                // gen: generation after the last launch; broken: a wait timed out
    unsigned gen;                // this launch's running generation (per thread copy, uniform)
};

// sum (or max) over a CTA of up to 1024 threads; result valid in every thread.  sm: [32]
template <bool MAX>
FLIP_D double blk_reduce(double v, double *sm) {
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAX ? fmax(v, u) : v + u;
    }
    __syncthreads();  // protect sm from a previous use
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = MAX ? fmax(r, sm[w]) : r + sm[w];
    return r;
}

#ifdef FLIP_CPU_EMU
// the emulator runs CTAs one after another: resident kernels are launched with a single CTA there
FLIP_D void grid_begin(GridCtx &) {}
FLIP_D void grid_end(GridCtx &) {}
template <bool AMAX, bool BMAX, bool VIS>
FLIP_D bool grid_allreduce2(GridCtx &, double &a, double &b, double *sm) {
    a = blk_reduce<AMAX>(a, sm);
    b = blk_reduce<BMAX>(b, sm);
    return true;
}
FLIP_D double ld_cg(const double *p) { return *p; }
FLIP_D float ld_cg(const float *p) { return *p; }
#else
FLIP_D unsigned long long grid_ld(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
FLIP_D void grid_st(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
FLIP_D void grid_begin(GridCtx &c) { c.gen = ((volatile GridBar *)c.bar)->gen; }   // before this CTA's first exchange, hence before anybody's grid_end
FLIP_D void grid_end(GridCtx &c) { if (blockIdx.x == 0 && threadIdx.x == 0) ((volatile GridBar *)c.bar)->gen = c.gen; }
FLIP_D void grid_put(unsigned long long *w, double a, double b, unsigned gen) {
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
    grid_st(w + 0, ((ua & 0xffffffffull) << 32) | gen);
    grid_st(w + 1, ((ua >> 32) << 32) | gen);
    grid_st(w + 2, ((ub & 0xffffffffull) << 32) | gen);
    grid_st(w + 3, ((ub >> 32) << 32) | gen);
}
// false = timed out (GridBar::broken is set)
FLIP_D bool grid_get(const unsigned long long *w, double &a, double &b, unsigned gen, GridBar *bar, bool backoff) {
    unsigned long long w0, w1, w2, w3;
    long long t0 = 0;
    int spins = 0;
    bool ok = true;
    while (true) {
        w0 = grid_ld(w + 0); w1 = grid_ld(w + 1); w2 = grid_ld(w + 2); w3 = grid_ld(w + 3);
        if ((unsigned)w0 == gen && (unsigned)w1 == gen && (unsigned)w2 == gen && (unsigned)w3 == gen) break;
        if (backoff) __nanosleep(20);
        if ((++spins & 255) == 0) {
            if (t0 == 0) t0 = clock64();
            if (((volatile GridBar *)bar)->broken || clock64() - t0 > 4000000000ll) { ((volatile GridBar *)bar)->broken = 1; ok = false; break; }
        }
    }
    a = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
    b = __longlong_as_double((long long)((w2 >> 32) | ((w3 >> 32) << 32)));
    return ok;
}
// a, b: this thread's contribution (combined over the CTA first).  Returns the fixed-order combination over all CTAs in
// every thread; false = the grid is broken.  sm: [32] doubles of shared memory.
template <bool AMAX, bool BMAX, bool VIS>
FLIP_D bool grid_allreduce2(GridCtx &c, double &a, double &b, double *sm) {
    __shared__ int grid_ok_s;
    __shared__ double grid_res_s[2];
    a = blk_reduce<AMAX>(a, sm);     // ends with a block barrier: every thread's earlier stores are ordered before thread 0's fence
    b = blk_reduce<BMAX>(b, sm);
    const unsigned gen = ++c.gen;
    unsigned long long *buf = c.slots + (size_t)(gen & 1u) * (GRID_MAX_CTAS + 1) * GRID_SLOT_WORDS;
    unsigned long long *res = buf + (size_t)GRID_MAX_CTAS * GRID_SLOT_WORDS;
    if (threadIdx.x == 0) {
        if (VIS) __threadfence();
        grid_put(buf + (size_t)blockIdx.x * GRID_SLOT_WORDS, a, b, gen);
        grid_ok_s = 1;
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        double va = 0.0, vb = 0.0;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
            double ua, ub;
            if (!grid_get(buf + (size_t)q * GRID_SLOT_WORDS, ua, ub, gen, c.bar, false)) grid_ok_s = 0;
            va = AMAX ? fmax(va, ua) : va + ua;
            vb = BMAX ? fmax(vb, ub) : vb + ub;
            if (VIS) __threadfence();
        }
        va = blk_reduce<AMAX>(va, sm);
        vb = blk_reduce<BMAX>(vb, sm);
        if (threadIdx.x == 0) {
            if (VIS) __threadfence();
            grid_put(res, va, vb, gen);
        }
    }
    if (threadIdx.x == 0) {
        double ra, rb;
        if (!grid_get(res, ra, rb, gen, c.bar, true)) grid_ok_s = 0;
        if (VIS) __threadfence();
        grid_res_s[0] = ra; grid_res_s[1] = rb;
    }
    __syncthreads();
    a = grid_res_s[0]; b = grid_res_s[1];
    const bool ok = grid_ok_s != 0;
    __syncthreads();   // the shared words are rewritten by the next exchange
    return ok;
}
FLIP_D double ld_cg(const double *p) { return __ldcg(p); }   // L2 only: written by other CTAs during this launch
FLIP_D float ld_cg(const float *p) { return __ldcg(p); }
#endif

#define PCGR_THREADS 512
#define PCGR_CPT_MAX 12

struct PcgResParams {
    Grid g;
    const int *cell_list, *cell_count;
    const float4 *coef;      // {diag, +i, +j, +k}
    const double *b;         // right-hand side (dense padded layout)
    double *p;               // search direction (dense; zero outside the unknowns on entry, like run_cg's s)
    double *x;               // solution out (dense)
    GridCtx grid;
    CGState *st;             // [2]: the final state is written to both slots
    double tol_abs;
    int maxit, strict;
};

template <int CPT>
__global__ void __launch_bounds__(PCGR_THREADS, 1) k_pcg_resident(PcgResParams P) {
#ifdef FLIP_CPU_EMU
    __shared__ float cs[8 * CPT * PCGR_THREADS];
#else
    extern __shared__ float cs[];
#endif
    __shared__ double sm[32];
    const int t = threadIdx.x, G = gridDim.x;
    GridCtx gc = P.grid;
    grid_begin(gc);
    const Grid &g = P.g;
    const int sy = SY(g), sz = SZ(g);
    const int nc = *P.cell_count;
    const int chunk = (nc + G - 1) / G;                       // cells of the list per CTA; host: chunk <= CPT * 512
    const int base = blockIdx.x * chunk;
    int mine = nc - base; mine = mine < 0 ? 0 : (mine > chunk ? chunk : mine);
    // shared-memory planes, [CPT][512] each: id | diag, -i, +i, -j, +j, -k, +k
    int *ids = (int *)cs;
    float *cd = cs + 1 * CPT * PCGR_THREADS, *cxm = cs + 2 * CPT * PCGR_THREADS, *cxp = cs + 3 * CPT * PCGR_THREADS,
          *cym = cs + 4 * CPT * PCGR_THREADS, *cyp = cs + 5 * CPT * PCGR_THREADS, *czm = cs + 6 * CPT * PCGR_THREADS,
          *czp = cs + 7 * CPT * PCGR_THREADS;
    double x[CPT], r[CPT], p[CPT], q[CPT];
    double rz = 0.0, bm = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int o = c * PCGR_THREADS + t;
        int id = -1;
        x[c] = 0.0; r[c] = 0.0; p[c] = 0.0; q[c] = 0.0;
        if (o < mine) {
            id = P.cell_list[base + o];
            const float4 a = P.coef[id];
            if (a.x != 0.0f) {
                cd[o] = a.x; cxp[o] = a.y; cyp[o] = a.z; czp[o] = a.w;
                cxm[o] = P.coef[id - 1].y; cym[o] = P.coef[id - sy].z; czm[o] = P.coef[id - sz].w;
                r[c] = P.b[id];
                p[c] = r[c] / (double)a.x;
                rz += r[c] * p[c];
                bm = fmax(bm, fabs(r[c]));
                P.p[id] = p[c];
            } else id = -1;   // a listed cell without an unknown (cannot happen for the scalar system; kept for safety)
        }
        ids[o] = id;
    }
    bool ok = grid_allreduce2<false, true, true>(gc, rz, bm, sm);            // also publishes p
    const double rho = rz, bmax = bm;
    CGState st;
    st.rho = rho; st.resid = bmax; st.tol = P.tol_abs; st.bmax = bmax; st.iter = 0; st.done = 0; st.converged = 0;
    st.maxit = P.maxit; st.fail = 0; st.first = 0; st.alpha = 0.0;
    if (bmax < P.tol_abs) { st.done = 1; st.converged = 1; }                 // zero pressure (src/pressuresolver.cpp:173-175)
    else if (!ok || rho == 0 || rho != rho) { st.done = 1; st.fail = 1; }
    while (!st.done) {
        // q = A p (row order of k_pressure_apply), p.q
        double pq = 0.0;
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const int o = c * PCGR_THREADS + t;
            const int id = ids[o];
            if (id < 0) continue;
            double val = 0.0;
            val += ld_cg(P.p + id - 1) * (double)cxm[o];
            val += ld_cg(P.p + id + 1) * (double)cxp[o];
            val += ld_cg(P.p + id - sy) * (double)cym[o];
            val += ld_cg(P.p + id + sy) * (double)cyp[o];
            val += ld_cg(P.p + id - sz) * (double)czm[o];
            val += ld_cg(P.p + id + sz) * (double)czp[o];
            val += p[c] * (double)cd[o];
            q[c] = val;
            pq += p[c] * val;
        }
        double none = 0.0;
        ok = grid_allreduce2<false, false, false>(gc, pq, none, sm);
        const double alpha = st.rho / pq;
        double rzn = 0.0, rm = 0.0;
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const int o = c * PCGR_THREADS + t;
            if (ids[o] < 0) continue;
            x[c] += alpha * p[c];
            r[c] -= alpha * q[c];
            rzn += r[c] * (r[c] / (double)cd[o]);
            rm = fmax(rm, fabs(r[c]));
        }
        ok = grid_allreduce2<false, true, false>(gc, rzn, rm, sm) && ok;
        const double rho_new = rzn, rmax = rm;
        const bool conv = P.strict ? (rmax < st.tol) : (rmax <= st.tol);
        const bool bad = !ok || !(rmax == rmax) || !(rho_new == rho_new);
        st.iter++;
        st.resid = rmax;
        st.converged = conv ? 1 : 0;
        if (bad) st.fail = 1;
        if (conv || bad || st.iter >= st.maxit) { st.done = 1; break; }
        const double beta = rho_new / st.rho;
        st.rho = rho_new;
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const int o = c * PCGR_THREADS + t;
            const int id = ids[o];
            if (id < 0) continue;
            p[c] = r[c] / (double)cd[o] + beta * p[c];
            P.p[id] = p[c];
        }
        double n0 = 0.0, n1 = 0.0;
        ok = grid_allreduce2<false, false, true>(gc, n0, n1, sm);   // everybody's new p is visible before the next stencil
    }
    grid_end(gc);
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int id = ids[c * PCGR_THREADS + t];
        if (id >= 0) P.x[id] = x[c];
    }
    if (blockIdx.x == 0 && t == 0) { P.st[0] = st; P.st[1] = st; }
}
