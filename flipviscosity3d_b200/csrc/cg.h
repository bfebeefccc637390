// Shared conjugate-gradient machinery for the pressure and viscosity solves.
//
// The reference runs MIC(0)-PCG on explicitly assembled matrices (src/pressuresolver.cpp:521-567,
// src/pcgsolver/pcgsolver.h:241-295); MIC(0) is a sequential recurrence.  Here the operator is a
// matrix-free stencil over coefficient fields, the preconditioner is the (parallel) diagonal, and
// the vectors are fp64 dense fields addressed through a list of active 8x8x8 blocks, so a solve
// only touches memory near the liquid.  Stopping rules are the reference's (max|r| against an
// absolute (pressure) or relative (viscosity) tolerance); parity is defined on the converged
// solution, not on the iterate path.
//
// One iteration = three kernels (A: q = A s, s.q;  B: x,r update, max|r|, r.z;  C: s = z + beta s).
// Dot products are reduced deterministically: every CTA writes one partial, and every CTA of the
// next kernel re-reduces the (<= FLIP_CG_MAXGRID) partials in a fixed order.  Iteration scalars
// live in two CGState slots used alternately, so no kernel reads a scalar another CTA of the same
// launch writes, and nothing goes back to the host except a poll every `cg_chunk` iterations.
#pragma once
#include "sim.h"
#include "xch.h"

#define CG_THREADS 512  // one thread per cell of an 8x8x8 block

struct BlockCell { int i, j, k; bool inside; };

FLIP_D BlockCell block_cell(const Grid &g, int blk, int t) {
    int bi = blk % g.nbx, r = blk / g.nbx;
    int bj = r % g.nby, bk = r / g.nby;
    BlockCell c;
    c.i = bi * FLIP_B + (t & 7);
    c.j = bj * FLIP_B + ((t >> 3) & 7);
    c.k = bk * FLIP_B + (t >> 6);
    c.inside = c.i <= g.ni && c.j <= g.nj && c.k <= g.nk;
    return c;
}

// Grid stages run over the list of 8x8x8 blocks near the liquid (fields.cu grid_list_ensure) instead of the dense grid:
// one CTA of CG_THREADS threads per listed block, thread = cell.  The body sees (i, j, k) of a cell inside
// [0, ni] x [0, nj] x [0, nk] exactly like the dense (n+1)^3 sweeps it replaces; `continue` leaves the cell.
#define FOR_LIST_CELLS(g, glist, gcount, i, j, k)                                                        \
    for (int _b = blockIdx.x, _nb = *(gcount); _b < _nb; _b += gridDim.x)                                \
        for (BlockCell _c = block_cell(g, (glist)[_b], threadIdx.x); _c.inside; _c.inside = false)      \
            for (int i = _c.i, j = _c.j, k = _c.k, _once = 1; _once; _once = 0)

// sum (or max) over the CTA; result valid in every thread
template <bool MAX>
FLIP_D double cta_reduce(double v, double *sm /*[CG_THREADS/32]*/) {
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAX ? fmax(v, u) : v + u;
    }
    __syncthreads();  // protect sm from a previous use
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < CG_THREADS / 32; w++) r = MAX ? fmax(r, sm[w]) : r + sm[w];
    return r;
}

// fixed-order reduction of n partials by a whole CTA; result valid in every thread
template <bool MAX>
FLIP_D double reduce_partials(const double *__restrict__ part, int n, double *sm) {
    double v = 0.0;
    for (int q = threadIdx.x; q < n; q += CG_THREADS) v = MAX ? fmax(v, part[q]) : v + part[q];
    return cta_reduce<MAX>(v, sm);
}

// Partials of reduction `kind` live at part[kind * nranks * G ...): G slots per rank, rank-major.  Every CTA stores its
// value into its own rank's slots; on a sharded handle the LAST CTA of the kernel then copies the rank's slots to every
// peer (PART_LEAVE: one coalesced burst and ONE system-scope fence per kernel - a fence in each of ~300 CTAs costs ~10 us,
// measured) and posts the hand-shake.  Consumers re-reduce all nranks * G values in that fixed order: bit-identical
// scalars on every rank, no all-reduce kernel.
#define PART_KINDS 6
#define PART_STORE(P, kind, value) (P).part[((size_t)(kind) * (P).X.nranks + (P).X.rank) * gridDim.x + blockIdx.x] = (value)
#define PART_PTR(P, kind) ((P).part + (size_t)(kind) * (P).X.nranks * gridDim.x)
#define PART_N(P) ((P).X.nranks * (int)gridDim.x)

// end of a kernel that produced partials (all threads of every CTA): replaces xch_leave
FLIP_D void part_leave(const Xch &X, double *part, double *const *part_peers) {
    if (X.nranks == 1) return;
    __shared__ int part_last_s;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                  // this CTA's partials before its arrival (device scope)
        unsigned int prev = atomicAdd(&X.local->cta_done, 1u);
        part_last_s = prev == gridDim.x - 1;
    }
    __syncthreads();
    if (!part_last_s) return;
    __threadfence();
    const int G = gridDim.x;
    for (int kind = 0; kind < PART_KINDS; kind++) {       // all kinds: re-sending an unchanged value is harmless
        const size_t o = ((size_t)kind * X.nranks + X.rank) * G;
        for (int q = threadIdx.x; q < G; q += blockDim.x) {
            const double v = ((volatile double *)part)[o + q];
            for (int p = 0; p < X.nranks; p++)
                if (p != X.rank) part_peers[p][o + q] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        xch_fence();                                      // the copies are performed before the flags
        X.local->cta_done = 0;
        unsigned long long n = X.local->done + 1;
        X.local->done = n;
        for (int p = 0; p < X.nranks; p++)
            if (p != X.rank) xch_st_relaxed(&((volatile unsigned long long *)X.peers[p]->arrive)[X.rank], n);
    }
}
#define PART_LEAVE(P) part_leave((P).X, (P).part, (P).part_peers)

struct CGParams {
    Grid g;
    const int *blk_list;
    const int *blk_count;
    const int *cell_list;      // padded indices of the cells of the active blocks that hold >= 1 unknown
    const int *cell_count;
    double *x, *r, *s, *q;     // [NC*total]
    double *z;                 // preconditioned residual (multigrid mode), or null for the diagonal
    double *part;              // reduction partials, kind-major: part[kind * nranks * G + rank * G + cta]
    double *const *part_peers; // device table [nranks]: `part` of every rank (sharded solves push their partials to all)
    Xch X;                     // exchange context (nranks == 1: single GPU)
    CGState *st;               // [2]
    int strict;                // 1: converged when max|r| < tol (pressure), 0: <= tol (viscosity)
    int flexible;              // multigrid mode: Polak-Ribiere beta = -(q.z)/(s.q), robust to an inexact V-cycle
};

// diagonal accessors
struct DiagPressure {
    const float4 *c;
    FLIP_D float operator()(int, int id) const { return c[id].x; }
};
struct DiagViscosity {
    const float *d; int total;
    FLIP_D float operator()(int comp, int id) const { return d[(size_t)comp * total + id]; }
};

// r holds b on entry.  x = 0, s = M^-1 r, partials of r.s and max|r|.
template <int NC, class Diag, bool KEEPX>
__global__ void __launch_bounds__(CG_THREADS) k_cg_init(CGParams P, Diag diag, double *__restrict__ bmax_part) {
    __shared__ double sm[CG_THREADS / 32];
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    int nc = *P.cell_count;
    double rz = 0.0, bm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            float d = diag(m, id);
            double r = d != 0.0f ? P.r[o] : 0.0;
            double z = d != 0.0f ? r / (double)d : 0.0;
            P.r[o] = r;
            if (!KEEPX) P.x[o] = 0.0;
            P.s[o] = z;
            rz += r * z;
            bm = fmax(bm, fabs(r));
        }
    }
    rz = cta_reduce<false>(rz, sm);
    bm = cta_reduce<true>(bm, sm);
    if (threadIdx.x == 0) {
        PART_STORE(P, 1, rz);
        // with a warm start max|b| (the tolerance reference) was measured before r became b - A x0
        if (!KEEPX) PART_STORE(P, 2, bm);
        else PART_STORE(P, 3, bm);   // max|r0|, informational
    }
    PART_LEAVE(P);
}

// max|b| partials only (warm start: the relative tolerance refers to b, not to r0)
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg_bmax(CGParams P, Diag diag) {
    __shared__ double sm[CG_THREADS / 32];
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    int nc = *P.cell_count;
    double bm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++)
            if (diag(m, id) != 0.0f) bm = fmax(bm, fabs(P.r[(size_t)m * g.total + id]));
    }
    bm = cta_reduce<true>(bm, sm);
    if (threadIdx.x == 0) PART_STORE(P, 2, bm);
    PART_LEAVE(P);
}

// Warm start (x0 != 0): k_cg_guess puts the guess into s (so the phase-A kernel computes q = A x0),
// k_cg_guess_residual then sets x = x0 and r = b - q; k_cg_init<KEEPX> continues from there.
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg_guess(CGParams P, Diag diag, const float *__restrict__ guess) {
    const Grid &g = P.g;
    int nc = *P.cell_count;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            P.s[o] = diag(m, id) != 0.0f ? (double)guess[o] : 0.0;
        }
    }
}
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg_guess_residual(CGParams P, Diag diag) {
    const Grid &g = P.g;
    int nc = *P.cell_count;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            bool unk = diag(m, id) != 0.0f;
            P.x[o] = unk ? P.s[o] : 0.0;
            P.r[o] = unk ? P.r[o] - P.q[o] : 0.0;
        }
    }
}

// single CTA: first CGState.  tol_rel > 0: tol = tol_rel * max|b| (pcgsolver.h:254-259), else tol_abs.
static __global__ void __launch_bounds__(CG_THREADS) k_cg_begin(CGParams P, int nparts, double tol_abs, double tol_rel, int maxit,
                                                                 int warm) {
    __shared__ double sm[CG_THREADS / 32];
    if (!xch_enter(P.X)) return;
    const int NP = P.X.nranks * nparts;   // launched with one CTA: the producers' grid size is passed in
    double rho = reduce_partials<false>(P.part + (size_t)NP, NP, sm);
    double bmax = reduce_partials<true>(P.part + 2 * (size_t)NP, NP, sm);
    double r0max = warm ? reduce_partials<true>(P.part + 3 * (size_t)NP, NP, sm) : bmax;
    if (threadIdx.x == 0) {
        CGState st;
        st.rho = rho; st.resid = r0max; st.bmax = bmax;
        st.tol = tol_rel > 0 ? tol_rel * bmax : tol_abs;
        st.iter = 0; st.maxit = maxit; st.fail = 0; st.first = 1; st.alpha = 0.0;
        st.done = 0; st.converged = 0;
        if (tol_rel > 0) {
            // viscosity: zero rhs -> solution 0, success (pcgsolver.h:254-258)
            if (bmax == 0) { st.done = 1; st.converged = 1; }
        } else {
            // pressure: max|b| < tol -> zero pressure (src/pressuresolver.cpp:173-175)
            if (bmax < tol_abs) { st.done = 1; st.converged = 1; }
        }
        // a warm start may already satisfy the stopping rule
        if (!st.done && warm && (P.strict ? r0max < st.tol : r0max <= st.tol)) { st.done = 1; st.converged = 1; }
        if (!st.done && (rho == 0 || rho != rho)) { st.done = 1; st.fail = 1; }
        P.st[0] = st;
        P.st[1] = st;
    }
    xch_leave(P.X, false);
}

// phase B: alpha = rho / s.q;  x += alpha s;  r -= alpha q;  partials of r.(M^-1 r) and max|r|
template <int NC, class Diag, bool MG>
__global__ void __launch_bounds__(CG_THREADS) k_cg_update(CGParams P, Diag diag, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    const CGState st = P.st[parity];
    if (st.done) return;
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    double sq = reduce_partials<false>(PART_PTR(P, 0), PART_N(P), sm);
    double alpha = st.rho / sq;
    int nc = *P.cell_count;
    double rz = 0.0, rm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        // all loads first (independent of the diagonal test): one memory round trip per cell
        float dd[NC];
        double sv[NC], qv[NC], xv[NC], rv[NC];
#pragma unroll
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            dd[m] = diag(m, id);
            sv[m] = P.s[o]; qv[m] = P.q[o]; xv[m] = P.x[o]; rv[m] = P.r[o];
        }
#pragma unroll
        for (int m = 0; m < NC; m++) {
            float d = dd[m];
            if (d == 0.0f) continue;
            size_t o = (size_t)m * g.total + id;
            P.x[o] = xv[m] + alpha * sv[m];
            double r = rv[m] - alpha * qv[m];
            P.r[o] = r;
            if (!MG) rz += r * (r / (double)d);
            rm = fmax(rm, fabs(r));
        }
    }
    rz = cta_reduce<false>(rz, sm);
    rm = cta_reduce<true>(rm, sm);
    if (threadIdx.x == 0) {
        if (!MG) PART_STORE(P, 1, rz);
        PART_STORE(P, 2, rm);
    }
    PART_LEAVE(P);
}

// multigrid mode: partials of r.z after the V-cycle
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg_dot(CGParams P, Diag diag, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    if (parity >= 0 && P.st[parity].done) return;
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    int nc = *P.cell_count;
    // at start-up (parity < 0) there is no q yet, and the fourth partial array still holds max|r0| of a warm start
    const bool flex = P.flexible && parity >= 0;
    double rz = 0.0, qz = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            if (diag(m, id) == 0.0f) continue;
            size_t o = (size_t)m * g.total + id;
            double z = P.z[o];
            rz += P.r[o] * z;
            if (flex) qz += P.q[o] * z;
        }
    }
    rz = cta_reduce<false>(rz, sm);
    if (flex) qz = cta_reduce<false>(qz, sm);
    if (threadIdx.x == 0) {
        PART_STORE(P, 1, rz);
        if (flex) PART_STORE(P, 3, qz);
    }
    PART_LEAVE(P);
}

// multigrid mode, start-up: x = 0, r = masked b, partial max|b|   (then V-cycle, k_cg_dot, k_cg_start)
// KEEPX (warm start): x and r = b - A x0 are already in place (k_cg_guess_residual); only max|r0| is reduced, into
// the fourth partial array, because the tolerance stays relative to max|b| (k_cg_bmax).
template <int NC, class Diag, bool KEEPX = false>
__global__ void __launch_bounds__(CG_THREADS) k_cg_init_mg(CGParams P, Diag diag) {
    __shared__ double sm[CG_THREADS / 32];
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    int nc = *P.cell_count;
    double bm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            double r = diag(m, id) != 0.0f ? P.r[o] : 0.0;
            P.r[o] = r;
            if (!KEEPX) P.x[o] = 0.0;
            bm = fmax(bm, fabs(r));
        }
    }
    bm = cta_reduce<true>(bm, sm);
    if (threadIdx.x == 0) PART_STORE(P, (KEEPX ? 3 : 2), bm);
    PART_LEAVE(P);
}

template <int NC>
__global__ void __launch_bounds__(CG_THREADS) k_cg_start_mg(CGParams P) {
    const Grid &g = P.g;
    int nc = *P.cell_count;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            P.s[o] = P.z[o];
        }
    }
}

// phase C: convergence bookkeeping into the other CGState slot, then s = M^-1 r + beta s
template <int NC, class Diag, bool MG>
__global__ void __launch_bounds__(CG_THREADS) k_cg_direction(CGParams P, Diag diag, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    const CGState st = P.st[parity];
    if (st.done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) P.st[parity ^ 1] = st;
        return;
    }
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    double rho_new = reduce_partials<false>(PART_PTR(P, 1), PART_N(P), sm);
    double rmax = reduce_partials<true>(PART_PTR(P, 2), PART_N(P), sm);
    bool conv = P.strict ? (rmax < st.tol) : (rmax <= st.tol);
    bool bad = !(rmax == rmax) || !(rho_new == rho_new);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CGState nx = st;
        nx.iter = st.iter + 1;
        nx.resid = rmax;
        nx.rho = rho_new;
        nx.converged = conv ? 1 : 0;
        nx.done = (conv || bad || nx.iter >= st.maxit) ? 1 : 0;
        if (bad) nx.fail = 1;
        P.st[parity ^ 1] = nx;
    }
    if (conv || bad) { xch_leave(P.X, false); return; }
    double beta = rho_new / st.rho;
    if (MG && P.flexible) {
        // beta = z_new.(r_new - r_old) / rho_old with r_new - r_old = -alpha q and alpha = rho_old / s.q
        double sq = reduce_partials<false>(PART_PTR(P, 0), PART_N(P), sm);
        double qz = reduce_partials<false>(PART_PTR(P, 3), PART_N(P), sm);
        beta = -qz / sq;
    }
    int nc = *P.cell_count;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        float dd[NC];
        double zv[NC], sv[NC];
#pragma unroll
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            dd[m] = diag(m, id);
            zv[m] = MG ? P.z[o] : P.r[o];
            sv[m] = P.s[o];
        }
#pragma unroll
        for (int m = 0; m < NC; m++) {
            float d = dd[m];
            if (d == 0.0f) continue;
            size_t o = (size_t)m * g.total + id;
            P.s[o] = (MG ? zv[m] : zv[m] / (double)d) + beta * sv[m];
        }
    }
    xch_leave(P.X, false);
}

// zero a [NC*total] double field on the cells of the listed blocks
template <int NC>
__global__ void __launch_bounds__(CG_THREADS) k_clear_blocks(Grid g, const int *__restrict__ list,
                                                              const int *__restrict__ count, double *__restrict__ f) {
    int nb = *count;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        BlockCell c = block_cell(g, list[b], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < NC; m++) f[(size_t)m * g.total + id] = 0.0;
    }
}

// ---- active block list -------------------------------------------------------------------
// flag[b] = 1 if any of the NC diagonal fields is non-zero inside block b
// unknowns (optional): total number of non-zero diagonals of the whole system (integer atomics: exact)
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_flag_blocks(Grid g, Diag diag, int *__restrict__ flag, int *__restrict__ unknowns) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    BlockCell c = block_cell(g, blockIdx.x, threadIdx.x);
    int mine = 0;
    if (c.inside) {
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < NC; m++) mine += diag(m, id) != 0.0f ? 1 : 0;
    }
    if (mine) any = 1;  // benign same-value race
    __syncthreads();
    if (threadIdx.x == 0) flag[blockIdx.x] = any;
    if (unknowns && any) {   // uniform across the CTA
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(unknowns, mine);
    }
}

// single CTA, ordered compaction (ascending block id => deterministic CTA/block assignment)
static __global__ void __launch_bounds__(1024) k_compact_blocks(const int *__restrict__ flag, int n, int *__restrict__ list,
                                                         int *__restrict__ count) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int lo = 0, hi = n;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        int id = base + threadIdx.x;
        int v = (id < n && id >= lo && id < hi && flag[id]) ? 1 : 0;   // [lo,hi): this rank's slab of blocks
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        int pos = carry_s + inc - v + (wid > 0 ? warp_sums[wid - 1] : 0);
        if (v) list[pos] = id;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pos + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry_s;
}

// ---- compact cell list -------------------------------------------------------------------
// On the bench scene only ~23 % of the cells of an active block hold an unknown (ncu: warps half
// empty, kernels latency bound).  The CG kernels therefore iterate over a compact, block-ordered
// list of the cells that hold at least one unknown: every thread has work, loads stay localised
// because the order is (block, then cell inside the block).
template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cell_counts(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                             Diag diag, int *__restrict__ per_block, const Cuts *__restrict__ cuts, int rank) {
    __shared__ int wsum[CG_THREADS / 32];
    int nb = *count;
    const int k0 = cuts ? cuts->c[0][rank] : 0, k1 = cuts ? cuts->c[0][rank + 1] : g.nk + 1;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        BlockCell c = block_cell(g, list[b], threadIdx.x);
        bool has = false;
        if (c.inside && c.k >= k0 && c.k < k1) {
            int id = gidx(g, c.i, c.j, c.k);
            for (int m = 0; m < NC; m++) has = has || diag(m, id) != 0.0f;
        }
        unsigned bal = __ballot_sync(0xffffffffu, has);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < CG_THREADS / 32; w++) t += wsum[w];
            per_block[b] = t;
        }
    }
}

// single CTA: exclusive scan of per_block[0..*count) in place, total to *total
static __global__ void __launch_bounds__(1024) k_scan_small(int *__restrict__ v, const int *__restrict__ count, int *__restrict__ total) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    int n = *count;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        int id = base + threadIdx.x;
        int x = id < n ? v[id] : 0;
        int inc = x;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        int excl = carry_s + inc - x + (wid > 0 ? warp_sums[wid - 1] : 0);
        if (id < n) v[id] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cell_fill(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                           Diag diag, const int *__restrict__ offset, int *__restrict__ cells,
                                                           const Cuts *__restrict__ cuts, int rank) {
    __shared__ int wsum[CG_THREADS / 32];
    int nb = *count;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k0 = cuts ? cuts->c[0][rank] : 0, k1 = cuts ? cuts->c[0][rank + 1] : g.nk + 1;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        BlockCell c = block_cell(g, list[b], threadIdx.x);
        bool has = false;
        int id = 0;
        if (c.inside && c.k >= k0 && c.k < k1) {
            id = gidx(g, c.i, c.j, c.k);
            for (int m = 0; m < NC; m++) has = has || diag(m, id) != 0.0f;
        }
        unsigned bal = __ballot_sync(0xffffffffu, has);
        __syncthreads();
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < wid; w++) before += wsum[w];
        int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (has) cells[offset[b] + rank] = id;
    }
}

static inline int list_grid(const Sim &s) {   // CTAs of a kernel over the near-liquid block list
    int gsz = s.num_sms * 4;
    return gsz > s.g.nblocks ? s.g.nblocks : gsz;
}

static inline int cg_grid(const Sim &s) {
    int gsz = s.num_sms * s.cg_grid_mult;
    return gsz > FLIP_CG_MAXGRID ? FLIP_CG_MAXGRID : gsz;
}

// block list of a whole (replicated) level: every rank builds the same list
template <int NC, class Diag>
static void build_block_list_on(Sim &s, const Grid &g, Diag diag, int *flag, int *list, int *count) {
    auto kflag = &k_flag_blocks<NC, Diag>;
    FLIP_LAUNCH_SYNC(kflag, g.nblocks, CG_THREADS, s.stream, g, diag, flag, (int *)nullptr);
    FLIP_LAUNCH_SYNC(k_compact_blocks, 1, 1024, s.stream, (const int *)flag, g.nblocks, list, count);
    s.kernel_launches += 2;
    KERNEL_CHECK();
}

// Work lists of a solve: the active 8x8x8 blocks of the WHOLE system (every rank builds the same list) and the compact,
// block-ordered list of the cells that hold at least one unknown inside this rank's slab of k-planes (all of them on one
// GPU).  s.unk_count = unknowns of the whole system.
template <int NC, class Diag>
static void build_block_list(Sim &s, Diag diag) {
    const Grid &g = s.g;
    const Cuts *cuts = xch_cuts(s);          // null on one GPU
    const int rank = xch_rank(s);
    auto kflag = &k_flag_blocks<NC, Diag>;
    CUDA_CHECK(cudaMemsetAsync(s.unk_count, 0, sizeof(int), s.stream));
    FLIP_LAUNCH_SYNC(kflag, g.nblocks, CG_THREADS, s.stream, g, diag, s.blk_flag, s.unk_count);
    FLIP_LAUNCH_SYNC(k_compact_blocks, 1, 1024, s.stream, (const int *)s.blk_flag, g.nblocks, s.blk_list, s.blk_count);
    auto kcc = &k_cell_counts<NC, Diag>;
    auto kcf = &k_cell_fill<NC, Diag>;
    FLIP_LAUNCH_SYNC(kcc, cg_grid(s), CG_THREADS, s.stream, g, (const int *)s.blk_list, (const int *)s.blk_count, diag, s.blk_flag, cuts, rank);
    FLIP_LAUNCH_SYNC(k_scan_small, 1, 1024, s.stream, s.blk_flag, (const int *)s.blk_count, s.cell_count);
    FLIP_LAUNCH_SYNC(kcf, cg_grid(s), CG_THREADS, s.stream, g, (const int *)s.blk_list, (const int *)s.blk_count, diag,
                     (const int *)s.blk_flag, s.cell_list, cuts, rank);
    s.kernel_launches += 5;
    KERNEL_CHECK();
}

// CGParams of a solve on the handle's own buffers
static CGParams cg_params(Sim &s, int strict) {
    CGParams P;
    P.g = s.g; P.blk_list = s.blk_list; P.blk_count = s.blk_count; P.cell_list = s.cell_list; P.cell_count = s.cell_count;
    P.x = s.cg_x; P.r = s.cg_r; P.s = s.cg_s; P.q = s.cg_q; P.z = nullptr;
    P.part = s.part; P.part_peers = s.part_peers; P.X = xch_of(s); P.st = s.cgst; P.strict = strict; P.flexible = 0;
    return P;
}

// A CUDA graph of an iteration chunk is only valid for the exchange set-up it was captured under
static inline unsigned long long cg_graph_tag(Sim &s, int chunk, int variant) {
    return (unsigned long long)chunk * 8 + variant * 2 + (s.sharded ? 1 : 0) + ((unsigned long long)s.xch_epoch << 20);
}

// Generic driver.  `apply(parity)` launches the phase-A kernel (q = A s and the s.q partials); on several ranks it also
// pushes the ghost planes of s first (the closure decides).  A chunk of `cg_chunk` iterations (3 kernels each) is
// captured ONCE into a CUDA graph per solver (all kernel arguments are pointers into the handle's own buffers and never
// change) and replayed; convergence is decided on the device, the host only polls the 64-byte state per chunk.
template <int NC, class Diag, class ApplyFn>
static CGState run_cg(Sim &s, CGParams P, Diag diag, double tol_abs, double tol_rel, int maxit, ApplyFn apply,
                      int graph_slot = -1, const float *guess = nullptr) {
    int G = cg_grid(s);
    auto kinit = &k_cg_init<NC, Diag, false>;
    auto kinit_keep = &k_cg_init<NC, Diag, true>;
    auto kupdate = &k_cg_update<NC, Diag, false>;
    auto kdir = &k_cg_direction<NC, Diag, false>;
    if (guess) {
        // r0 = b - A x0 with x0 = the guess; the tolerance stays relative to max|b|
        auto kguess = &k_cg_guess<NC, Diag>;
        auto kres = &k_cg_guess_residual<NC, Diag>;
        auto kbmax = &k_cg_bmax<NC, Diag>;
        FLIP_LAUNCH_SYNC(kbmax, G, CG_THREADS, s.stream, P, diag);
        FLIP_LAUNCH(kguess, G, CG_THREADS, s.stream, P, diag, guess);
        // the phase-A kernel tests st[parity].done: make sure slot 0 says "not done"
        CUDA_CHECK(cudaMemsetAsync(s.cgst, 0, 2 * sizeof(CGState), s.stream));
        apply(0);
        FLIP_LAUNCH(kres, G, CG_THREADS, s.stream, P, diag);
        FLIP_LAUNCH_SYNC(kinit_keep, G, CG_THREADS, s.stream, P, diag, (double *)nullptr);
        s.kernel_launches += 5;
    } else {
        FLIP_LAUNCH_SYNC(kinit, G, CG_THREADS, s.stream, P, diag, (double *)nullptr);
        s.kernel_launches += 1;
    }
    FLIP_LAUNCH_SYNC(k_cg_begin, 1, CG_THREADS, s.stream, P, G, tol_abs, tol_rel, maxit, guess ? 1 : 0);
    s.kernel_launches += 1;
    KERNEL_CHECK();
    int chunk = s.cg_chunk < 2 ? 2 : (s.cg_chunk & ~1);
    auto launch_chunk = [&]() {
        for (int it = 0; it < chunk; it++) {
            int parity = it & 1;
            apply(parity);
            FLIP_LAUNCH_SYNC(kupdate, G, CG_THREADS, s.stream, P, diag, parity);
            FLIP_LAUNCH_SYNC(kdir, G, CG_THREADS, s.stream, P, diag, parity);
        }
    };
#ifndef FLIP_CPU_EMU
    bool use_graph = s.use_graphs && graph_slot >= 0 && graph_slot < 2;
    const unsigned long long tag = cg_graph_tag(s, chunk, 0);
    if (use_graph && (!s.cg_graph[graph_slot] || s.cg_graph_tag[graph_slot] != tag)) {
        if (s.cg_graph[graph_slot]) { cudaGraphExecDestroy((cudaGraphExec_t)s.cg_graph[graph_slot]); s.cg_graph[graph_slot] = nullptr; }
        cudaGraph_t graph = nullptr;
        long long keep = s.kernel_launches;
        CUDA_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        launch_chunk();
        CUDA_CHECK(cudaStreamEndCapture(s.stream, &graph));
        s.cg_graph_launches[graph_slot] = s.kernel_launches - keep + 2 * chunk;
        s.kernel_launches = keep;
        cudaGraphExec_t exec = nullptr;
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        s.cg_graph[graph_slot] = (void *)exec;
        s.cg_graph_tag[graph_slot] = tag;
    }
#else
    bool use_graph = false;
#endif
    CGState h;
    int launched = 0;
    while (true) {
        CUDA_CHECK(cudaMemcpyAsync(s.cgst_host, s.cgst, sizeof(CGState), cudaMemcpyDeviceToHost, s.stream));
        xch_status_fetch(s);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        h = *s.cgst_host;
        if (xch_status_bad(s)) { h.fail = 1; h.converged = 0; break; }   // broken exchange: stop launching, xch_check reports it
        if (h.done || launched >= maxit + chunk) break;
#ifndef FLIP_CPU_EMU
        if (use_graph) { CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)s.cg_graph[graph_slot], s.stream)); s.kernel_launches += s.cg_graph_launches[graph_slot]; }
        else
#endif
        { launch_chunk(); s.kernel_launches += 2 * chunk; }
        KERNEL_CHECK();
        launched += chunk;
    }
    return h;
}

// ------------------------------------------------------------------------------------------
// Single-reduction CG (Chronopoulos & Gear recurrences), diagonal preconditioner.
//
//   u = M^-1 r,  w = A u,  gamma = r.u,  delta = w.u
//   beta = gamma/gamma_old,  alpha = gamma / (delta - beta*gamma/alpha_old)
//   p = u + beta p,  s = w + beta s (= A p),  x += alpha p,  r -= alpha s
//
// Two kernels per iteration instead of three and ONE point where scalars are needed:
//   K1 (k_cg2_step)  reduces the partials of gamma, delta and max|r| left by the previous iteration,
//                    decides convergence, updates p, s, x, r, u in one pass and leaves the partials
//                    of the new gamma and max|r| (ping-pong kinds: other CTAs still read the old);
//   K2 (phase-A stencil kernel, unchanged, with s:=u and q:=w) computes w = A u and the partials of
//                    delta.
// Same operator, same stopping rule (max|r| against the tolerance, tested on the residual BEFORE each
// update), same converged solution as run_cg.
// Partial kinds: 0 delta | 1 gamma(0) | 2 rmax(0) | 3 gamma(1) | 4 rmax(1)
// ------------------------------------------------------------------------------------------
struct CG2Params {
    CGParams P;      // x, r, s(=p), q(=s=Ap) as in CGParams; z = u
    double *w;       // w = A u
};

template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg2_init(CG2Params Q, Diag diag) {
    __shared__ double sm[CG_THREADS / 32];
    const CGParams &P = Q.P;
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    int nc = *P.cell_count;
    double gam = 0.0, bm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
#pragma unroll
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            float d = diag(m, id);
            double r = d != 0.0f ? P.r[o] : 0.0;
            double u = d != 0.0f ? r / (double)d : 0.0;
            P.r[o] = r;
            P.x[o] = 0.0;
            P.z[o] = u;
            gam += r * u;
            bm = fmax(bm, fabs(r));
        }
    }
    gam = cta_reduce<false>(gam, sm);
    bm = cta_reduce<true>(bm, sm);
    if (threadIdx.x == 0) {
        PART_STORE(P, 1, gam);       // gamma(0)
        PART_STORE(P, 2, bm);        // rmax(0) = max|b|
    }
    PART_LEAVE(P);
}

template <int NC, class Diag>
__global__ void __launch_bounds__(CG_THREADS) k_cg2_step(CG2Params Q, Diag diag, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    const CGParams &P = Q.P;
    const CGState st = P.st[parity];
    const int G = gridDim.x;
    if (st.done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) P.st[parity ^ 1] = st;
        return;
    }
    if (!xch_enter(P.X)) return;
    double rmax = reduce_partials<true>(PART_PTR(P, parity ? 4 : 2), PART_N(P), sm);
    double gam = reduce_partials<false>(PART_PTR(P, parity ? 3 : 1), PART_N(P), sm);
    double del = reduce_partials<false>(PART_PTR(P, 0), PART_N(P), sm);
    bool conv = P.strict ? (rmax < st.tol) : (rmax <= st.tol);
    double beta = st.first ? 0.0 : gam / st.rho;
    double alpha = st.first ? gam / del : gam / (del - beta * gam / st.alpha);
    bool bad = !(alpha == alpha) || !(rmax == rmax) || (st.first && (gam == 0 || !(gam == gam)));
    bool stop = conv || bad || st.iter >= st.maxit;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CGState nx = st;
        nx.resid = rmax;
        nx.converged = conv ? 1 : 0;
        nx.done = stop ? 1 : 0;
        if (bad && !conv) nx.fail = 1;
        if (!stop) { nx.iter = st.iter + 1; nx.rho = gam; nx.alpha = alpha; nx.first = 0; }
        P.st[parity ^ 1] = nx;
    }
    if (stop) { xch_leave(P.X, false); return; }
    const Grid &g = P.g;
    const bool first = st.first != 0;
    int nc = *P.cell_count;
    double gnew = 0.0, rm = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += G * CG_THREADS) {
        int id = P.cell_list[qq];
        float dd[NC];
        double uv[NC], wv[NC], pv[NC], sv[NC], xv[NC], rv[NC];
#pragma unroll
        for (int m = 0; m < NC; m++) {
            size_t o = (size_t)m * g.total + id;
            dd[m] = diag(m, id);
            uv[m] = P.z[o]; wv[m] = Q.w[o]; xv[m] = P.x[o]; rv[m] = P.r[o];
            pv[m] = first ? 0.0 : P.s[o];
            sv[m] = first ? 0.0 : P.q[o];
        }
#pragma unroll
        for (int m = 0; m < NC; m++) {
            float d = dd[m];
            if (d == 0.0f) continue;
            size_t o = (size_t)m * g.total + id;
            double pn = uv[m] + beta * pv[m];
            double sn = wv[m] + beta * sv[m];
            double r = rv[m] - alpha * sn;
            double u = r / (double)d;
            P.s[o] = pn;
            P.q[o] = sn;
            P.x[o] = xv[m] + alpha * pn;
            P.r[o] = r;
            P.z[o] = u;
            gnew += r * u;
            rm = fmax(rm, fabs(r));
        }
    }
    gnew = cta_reduce<false>(gnew, sm);
    rm = cta_reduce<true>(rm, sm);
    if (threadIdx.x == 0) { PART_STORE(P, parity ? 1 : 3, gnew); PART_STORE(P, parity ? 2 : 4, rm); }
    PART_LEAVE(P);
}

// `apply_uw(parity)` must launch the phase-A kernel on a CGParams whose s is u (P.z) and q is w: it computes w = A u and
// leaves the partials of u.w in kind 0; on several ranks it pushes the ghost planes of u first.
template <int NC, class Diag, class ApplyFn>
static CGState run_cg2(Sim &s, CGParams P, Diag diag, double tol_abs, double tol_rel, int maxit, ApplyFn apply_uw, int graph_slot) {
    int G = cg_grid(s);
    CG2Params Q;
    Q.P = P; Q.P.z = s.cg_z; Q.w = s.cg_w;
    auto kinit = &k_cg2_init<NC, Diag>;
    auto kstep = &k_cg2_step<NC, Diag>;
    FLIP_LAUNCH_SYNC(kinit, G, CG_THREADS, s.stream, Q, diag);
    // first CGState: reuses k_cg_begin (rho := gamma0, tolerance, trivial-rhs exits)
    FLIP_LAUNCH_SYNC(k_cg_begin, 1, CG_THREADS, s.stream, P, G, tol_abs, tol_rel, maxit, 0);
    apply_uw(0);
    s.kernel_launches += 2;
    KERNEL_CHECK();
    int chunk = s.cg_chunk < 2 ? 2 : (s.cg_chunk & ~1);
    auto launch_chunk = [&]() {
        for (int it = 0; it < chunk; it++) {
            int parity = it & 1;
            FLIP_LAUNCH_SYNC(kstep, G, CG_THREADS, s.stream, Q, diag, parity);
            apply_uw(parity ^ 1);
        }
    };
#ifndef FLIP_CPU_EMU
    bool use_graph = s.use_graphs && graph_slot >= 0 && graph_slot < 2;
    const unsigned long long tag = cg_graph_tag(s, chunk, 1);
    if (use_graph && (!s.cg_graph[graph_slot] || s.cg_graph_tag[graph_slot] != tag)) {
        if (s.cg_graph[graph_slot]) { cudaGraphExecDestroy((cudaGraphExec_t)s.cg_graph[graph_slot]); s.cg_graph[graph_slot] = nullptr; }
        cudaGraph_t graph = nullptr;
        long long keep = s.kernel_launches;
        CUDA_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        launch_chunk();
        CUDA_CHECK(cudaStreamEndCapture(s.stream, &graph));
        s.cg_graph_launches[graph_slot] = s.kernel_launches - keep + chunk;
        s.kernel_launches = keep;
        cudaGraphExec_t exec = nullptr;
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        s.cg_graph[graph_slot] = (void *)exec;
        s.cg_graph_tag[graph_slot] = tag;
    }
#else
    bool use_graph = false;
#endif
    CGState h;
    int launched = 0;
    while (true) {
        CUDA_CHECK(cudaMemcpyAsync(s.cgst_host, s.cgst, sizeof(CGState), cudaMemcpyDeviceToHost, s.stream));
        xch_status_fetch(s);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        h = *s.cgst_host;
        if (xch_status_bad(s)) { h.fail = 1; h.converged = 0; break; }   // broken exchange: stop launching, xch_check reports it
        if (h.done || launched >= maxit + 2 * chunk) break;
#ifndef FLIP_CPU_EMU
        if (use_graph) { CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)s.cg_graph[graph_slot], s.stream)); s.kernel_launches += s.cg_graph_launches[graph_slot]; }
        else
#endif
        { launch_chunk(); s.kernel_launches += chunk; }
        KERNEL_CHECK();
        launched += chunk;
    }
    return h;
}
