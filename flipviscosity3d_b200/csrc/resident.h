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
// One synchronisation point = one all-to-all flag exchange: every CTA stores up to two doubles into its own slot, each
// split into two 8-byte words {32 data bits, generation} (8-byte stores are single-copy atomic, so a reader that sees the
// generation also sees the data: no fence, no atomics, no master CTA), then polls the slots of all CTAs and re-reduces
// them in a fixed order.  Barrier and all-reduce in one L2 round trip (~2 us on B200 against ~4 us for an atomic counter
// barrier followed by a read-back of the partials, measured on the pressure solve).  Slots are double buffered by the
// parity of the generation: a CTA can only run one exchange ahead of the slowest one.  VIS: the exchange also orders
// this CTA's earlier global stores before the other CTAs' later loads (one fence in thread 0 on each side).
// All CTAs must be co-resident (cooperative launch, grid <= SMs).  A wait that exceeds ~2 s sets GridBar::broken and
// every later exchange returns at once with ok = false: a lost CTA cannot hang the device.
#define GRID_MAX_CTAS 1024
#define GRID_SLOT_WORDS 4
struct GridCtx {
    unsigned long long *slots;   // [2][GRID_MAX_CTAS][GRID_SLOT_WORDS]
    GridBar *bar;                // gen: generation after the last launch; broken: a wait timed out
    unsigned gen;                // this launch's running generation (per thread copy, uniform)
};

#ifdef FLIP_CPU_EMU
// the emulator runs CTAs one after another: resident kernels are launched with a single CTA there
FLIP_D void grid_begin(GridCtx &) {}
FLIP_D void grid_end(GridCtx &) {}
template <bool AMAX, bool BMAX, bool VIS>
FLIP_D bool grid_allreduce2(GridCtx &, double &a, double &b, double *sm) {
    a = cta_reduce<AMAX>(a, sm);
    b = cta_reduce<BMAX>(b, sm);
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
// a, b: this CTA's contribution (any thread's values are combined over the CTA first).  Returns the fixed-order
// combination over all CTAs in every thread; false = the grid is broken.
template <bool AMAX, bool BMAX, bool VIS>
FLIP_D bool grid_allreduce2(GridCtx &c, double &a, double &b, double *sm) {
    __shared__ int grid_ok_s;
    a = cta_reduce<AMAX>(a, sm);     // ends with a block barrier: every thread's earlier stores are ordered before thread 0's fence
    b = cta_reduce<BMAX>(b, sm);
    const unsigned gen = ++c.gen;
    unsigned long long *buf = c.slots + (size_t)(gen & 1u) * GRID_MAX_CTAS * GRID_SLOT_WORDS;
    if (threadIdx.x == 0) {
        if (VIS) __threadfence();
        const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
        unsigned long long *w = buf + (size_t)blockIdx.x * GRID_SLOT_WORDS;
        grid_st(w + 0, ((ua & 0xffffffffull) << 32) | gen);
        grid_st(w + 1, ((ua >> 32) << 32) | gen);
        grid_st(w + 2, ((ub & 0xffffffffull) << 32) | gen);
        grid_st(w + 3, ((ub >> 32) << 32) | gen);
        grid_ok_s = 1;
    }
    __syncthreads();
    double va = 0.0, vb = 0.0;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
        const unsigned long long *w = buf + (size_t)q * GRID_SLOT_WORDS;
        unsigned long long w0, w1, w2, w3;
        long long t0 = 0;
        int spins = 0;
        while (true) {
            w0 = grid_ld(w + 0); w1 = grid_ld(w + 1); w2 = grid_ld(w + 2); w3 = grid_ld(w + 3);
            if ((unsigned)w0 == gen && (unsigned)w1 == gen && (unsigned)w2 == gen && (unsigned)w3 == gen) break;
            if ((++spins & 1023) == 0) {
                if (t0 == 0) t0 = clock64();
                if (((volatile GridBar *)c.bar)->broken || clock64() - t0 > 4000000000ll) { ((volatile GridBar *)c.bar)->broken = 1; grid_ok_s = 0; break; }
            }
        }
        const double ua = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
        const double ub = __longlong_as_double((long long)((w2 >> 32) | ((w3 >> 32) << 32)));
        va = AMAX ? fmax(va, ua) : va + ua;
        vb = BMAX ? fmax(vb, ub) : vb + ub;
        if (VIS) __threadfence();    // acquire: loads after the exchange see what the publishers stored before it
    }
    a = cta_reduce<AMAX>(va, sm);
    b = cta_reduce<BMAX>(vb, sm);
    return grid_ok_s != 0;
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
    __shared__ double sm[PCGR_THREADS / 32];
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
