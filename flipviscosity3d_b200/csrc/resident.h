// Persistent ("resident") kernels: the whole grid stays on the SMs for a complete iterative phase and synchronises
// through a grid-wide barrier instead of kernel boundaries.
//
// k_pcg_resident: the ENTIRE diagonal-preconditioned CG of the pressure projection in one launch.  The system of the
// headline scene (0.62 M unknowns at 256^3) is small enough that its per-cell solver state fits ON CHIP: every thread owns
// up to CPT cells of the compact cell list and keeps x, r, p, q of those cells in registers and the seven stencil
// coefficients of each in shared memory (<= 196 KB per CTA, one CTA per SM).  Per iteration only the search direction
// travels - one 8-byte store per cell, six neighbour loads served by L2 - and three grid barriers replace the launch
// boundaries and the host's convergence polls of the kernel-per-phase path (cg.h): the solve is latency bound (the
// working set of the launch-per-phase version already sat in the 126 MB L2), so what is removed is launch latency.
// Same operator, same preconditioner, same stopping rule (max|r| < tol before maxit, src/pressuresolver.cpp:521-567) and
// the same deterministic fixed-order reductions as run_cg; systems that do not fit (more than CPT_MAX * 512 cells per
// SM) or that are cut into slabs across GPUs take the launch-per-phase path.
#pragma once
#include "cg.h"

#ifdef FLIP_CPU_EMU
// the emulator runs CTAs one after another: resident kernels are launched with a single CTA there
FLIP_D void grid_sync(GridBar *, unsigned &) { __syncthreads(); }
FLIP_D double ld_cg(const double *p) { return *p; }
#else
FLIP_D unsigned grid_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// All CTAs of the grid must be co-resident (cooperative launch).  `gen` is the caller's copy of the generation counter
// (read once at kernel start, before the first arrival, hence before the first barrier can complete).
FLIP_D void grid_sync(GridBar *b, unsigned &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // release: this CTA's stores before its arrival
        if (atomicAdd(&b->count, 1u) == gridDim.x - 1) {
            b->count = 0;                                  // nobody touches count again before gen moves
            __threadfence();
            atomicExch(&b->gen, gen + 1u);
        } else {
            while (grid_ld_acquire(&b->gen) == gen) {}
        }
        __threadfence();                                   // acquire (also drops this SM's stale L1 lines)
    }
    gen++;
    __syncthreads();
}
FLIP_D double ld_cg(const double *p) { return __ldcg(p); }   // L2 only: written by other CTAs during this launch
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
    double *part;            // [3 * gridDim.x] reduction partials
    GridBar *bar;
    CGState *st;             // [2]: the final state is written to both slots
    double tol_abs;
    int maxit, strict;
};

// fixed-order sum (or max) of one partial per CTA; result in every thread
template <bool MAX>
FLIP_D double pcgr_reduce(double mine, double *part, GridBar *bar, unsigned &gen, double *sm) {
    double v = cta_reduce<MAX>(mine, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    grid_sync(bar, gen);
    double a = 0.0;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += PCGR_THREADS) { double u = ld_cg(part + q); a = MAX ? fmax(a, u) : a + u; }
    return cta_reduce<MAX>(a, sm);
}

template <int CPT>
__global__ void __launch_bounds__(PCGR_THREADS, 1) k_pcg_resident(PcgResParams P) {
#ifdef FLIP_CPU_EMU
    __shared__ float cs[8 * CPT * PCGR_THREADS];
#else
    extern __shared__ float cs[];
#endif
    __shared__ double sm[PCGR_THREADS / 32];
    const int t = threadIdx.x, G = gridDim.x;
    unsigned gen = 0;
#ifndef FLIP_CPU_EMU
    gen = grid_ld_acquire(&P.bar->gen);
#endif
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
    double rho = pcgr_reduce<false>(rz, P.part, P.bar, gen, sm);             // also publishes p
    const double bmax = pcgr_reduce<true>(bm, P.part + G, P.bar, gen, sm);
    CGState st;
    st.rho = rho; st.resid = bmax; st.tol = P.tol_abs; st.bmax = bmax; st.iter = 0; st.done = 0; st.converged = 0;
    st.maxit = P.maxit; st.fail = 0; st.first = 0; st.alpha = 0.0;
    if (bmax < P.tol_abs) { st.done = 1; st.converged = 1; }                 // zero pressure (src/pressuresolver.cpp:173-175)
    else if (rho == 0 || rho != rho) { st.done = 1; st.fail = 1; }
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
        pq = pcgr_reduce<false>(pq, P.part, P.bar, gen, sm);
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
        rzn = cta_reduce<false>(rzn, sm);
        if (t == 0) P.part[G + blockIdx.x] = rzn;
        const double rmax = pcgr_reduce<true>(rm, P.part + 2 * G, P.bar, gen, sm);
        double rho_new = 0.0;
        for (int qq = t; qq < G; qq += PCGR_THREADS) rho_new += ld_cg(P.part + G + qq);
        rho_new = cta_reduce<false>(rho_new, sm);
        const bool conv = P.strict ? (rmax < st.tol) : (rmax <= st.tol);
        const bool bad = !(rmax == rmax) || !(rho_new == rho_new);
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
        grid_sync(P.bar, gen);   // everybody's new p is visible before the next stencil
    }
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int id = ids[c * PCGR_THREADS + t];
        if (id >= 0) P.x[id] = x[c];
    }
    if (blockIdx.x == 0 && t == 0) { P.st[0] = st; P.st[1] = st; }
}
