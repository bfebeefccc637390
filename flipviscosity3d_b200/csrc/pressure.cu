// Variational free-surface pressure projection.
//
// Reference behaviour being reproduced (relative to /root/reference):
//   unknown set, rhs, coefficients   src/pressuresolver.cpp:196-322
//   CG stopping rule                  src/pressuresolver.cpp:521-567 (max|r| < 1e-9 absolute)
//   result cast to float              src/pressuresolver.cpp:186-191
//   _project                          src/fluidsimulation.cpp:522-531
//
// The reference assembles {diag,+i,+j,+k} per unknown and addresses neighbours through a dense
// int keymap with a sequential MIC(0) preconditioner.  Here the same four float coefficients are
// a dense float4 stencil field (one 16-byte load per cell), the operator is applied matrix-free
// on the active 8x8x8 blocks, and the preconditioner is the diagonal (see cg.h).
#include "cg.h"
#include "levelset_math.h"
#include "resident.h"

// rhs (into r) and stencil coefficients for every cell; non-unknown cells get zeros.
__global__ void __launch_bounds__(CG_THREADS) k_pressure_build(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                        const float *__restrict__ phi, const float *__restrict__ vel,
                                                        const float *__restrict__ weight, float4 *__restrict__ coef,
                                                        double *__restrict__ rhs, float scale, float minfrac) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    double b = 0.0;
    // unknowns: interior cells with phi < 0 (src/pressuresolver.cpp:206-215)
    bool unknown = i >= 1 && i < g.ni - 1 && j >= 1 && j < g.nj - 1 && k >= 1 && k < g.nk - 1 && phi[id] < 0;
    if (unknown) {
        const float *wu = weight, *wv = weight + T, *ww = weight + 2 * T;
        const float *u = vel, *v = vel + T, *w = vel + 2 * T;
        // negative divergence (src/pressuresolver.cpp:235-244): float products, double accumulation
        b -= (double)(wu[id + 1] * u[id + 1]);
        b += (double)(wu[id] * u[id]);
        b -= (double)(wv[id + sy] * v[id + sy]);
        b += (double)(wv[id] * v[id]);
        b -= (double)(ww[id + sz] * w[id + sz]);
        b += (double)(ww[id] * w[id]);
        b /= g.dxd;
        float p0 = phi[id];
        // coefficients (src/pressuresolver.cpp:259-320), float accumulation in the reference's order
        float term = wu[id + 1] * scale;
        float pn = phi[id + 1];
        if (pn < 0) { c.x += term; c.y -= term; }
        else { c.x += term / fmaxf(frac_inside2(p0, pn), minfrac); }
        term = wu[id] * scale;
        pn = phi[id - 1];
        if (pn < 0) { c.x += term; }
        else { c.x += term / fmaxf(frac_inside2(pn, p0), minfrac); }
        term = wv[id + sy] * scale;
        pn = phi[id + sy];
        if (pn < 0) { c.x += term; c.z -= term; }
        else { c.x += term / fmaxf(frac_inside2(p0, pn), minfrac); }
        term = wv[id] * scale;
        pn = phi[id - sy];
        if (pn < 0) { c.x += term; }
        else { c.x += term / fmaxf(frac_inside2(pn, p0), minfrac); }
        term = ww[id + sz] * scale;
        pn = phi[id + sz];
        if (pn < 0) { c.x += term; c.w -= term; }
        else { c.x += term / fmaxf(frac_inside2(p0, pn), minfrac); }
        term = ww[id] * scale;
        pn = phi[id - sz];
        if (pn < 0) { c.x += term; }
        else { c.x += term / fmaxf(frac_inside2(pn, p0), minfrac); }
    }
    coef[id] = c;
    rhs[id] = b;
  }
}

// phase A: q = A s on the active blocks (row order of src/pressuresolver.cpp:464-499) and s.q partials
__global__ void __launch_bounds__(CG_THREADS) k_pressure_apply(CGParams P, const float4 *__restrict__ coef, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    if (P.st[parity].done) return;
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    const int sy = SY(g), sz = SZ(g);
    const double *__restrict__ s = P.s;
    int nc = *P.cell_count;
    double sq = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        float4 a = coef[id];
        if (a.x == 0.0f) continue;
        double sc = s[id];
        double val = 0.0;
        val += s[id - 1] * (double)coef[id - 1].y;
        val += s[id + 1] * (double)a.y;
        val += s[id - sy] * (double)coef[id - sy].z;
        val += s[id + sy] * (double)a.z;
        val += s[id - sz] * (double)coef[id - sz].w;
        val += s[id + sz] * (double)a.w;
        val += sc * (double)a.x;
        P.q[id] = val;
        sq += sc * val;
    }
    sq = cta_reduce<false>(sq, sm);
    if (threadIdx.x == 0) PART_STORE(P, 0, sq);
    PART_LEAVE(P);
}

__global__ void __launch_bounds__(CG_THREADS) k_pressure_store(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                                const float4 *__restrict__ coef, const double *__restrict__ x, float *__restrict__ pr) {
    FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
        if (i >= g.ni || j >= g.nj || k >= g.nk) continue;
        int id = gidx(g, i, j, k);
        pr[id] = coef[id].x != 0.0f ? (float)x[id] : 0.0f;
    }
}

// The whole CG in one persistent launch (resident.h).  Returns false when the system does not fit on chip (or the handle
// is sharded): the caller then runs the launch-per-phase path.
template <int CPT>
static void launch_pcg_resident(Sim &s, PcgResParams &P, int G) {
    auto k = &k_pcg_resident<CPT>;
#ifdef FLIP_CPU_EMU
    FLIP_LAUNCH_SYNC(k, 1, PCGR_THREADS, s.stream, P);
#else
    const size_t smem = (size_t)8 * CPT * PCGR_THREADS * sizeof(float);
    CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&P};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k, dim3(G), dim3(PCGR_THREADS), args, smem, s.stream));
#endif
}
static bool solve_pressure_resident(Sim &s, int pmaxit, CGState &h) {
    if (!s.pres_resident || xch_of(s).nranks != 1) return false;
    CUDA_CHECK(cudaMemcpyAsync(s.count_host, s.cell_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    const int nc = s.count_host[0];
#ifdef FLIP_CPU_EMU
    const int G = 1;
#else
    const int G = s.num_sms;
#endif
    const int need = cdiv(cdiv(nc, G), PCGR_THREADS);
    if (need > PCGR_CPT_MAX) return false;
    PcgResParams P;
    P.g = s.g; P.cell_list = s.cell_list; P.cell_count = s.cell_count; P.coef = s.pcoef; P.b = s.cg_r; P.p = s.cg_s; P.x = s.cg_x;
    P.grid.slots = s.grid_slots; P.grid.bar = s.grid_bar; P.grid.gen = 0; P.st = s.cgst; P.tol_abs = s.pressure_tol; P.maxit = pmaxit; P.strict = 1;
    if (need <= 2) launch_pcg_resident<2>(s, P, G);
    else if (need <= 4) launch_pcg_resident<4>(s, P, G);
    else if (need <= 6) launch_pcg_resident<6>(s, P, G);
    else if (need <= 8) launch_pcg_resident<8>(s, P, G);
    else if (need <= 10) launch_pcg_resident<10>(s, P, G);
    else launch_pcg_resident<PCGR_CPT_MAX>(s, P, G);
    s.kernel_launches++;
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(s.cgst_host, s.cgst, sizeof(CGState), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    h = *s.cgst_host;
    return true;
}

void solve_pressure(Sim &s, float dt) {
    const Grid &g = s.g;
    // small systems are not worth their hand-shakes: every rank solves the whole system (sim.h shard_min_unknowns)
    ReplicateGuard replicate(s, s.sharded && (long long)s.pres_last_unknowns < s.shard_min_unknowns);
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    CUDA_CHECK(cudaEventRecord(e0, s.stream));
    xch_update_cuts(s);   // k-slabs balanced by liquid cells (no-op on one GPU)
    // scale = deltaTime / (dx*dx) in double, used as (float)scale (src/pressuresolver.cpp:250, 259)
    double scale = (double)dt / (g.dxd * g.dxd);
    grid_list_ensure(s);
    FLIP_LAUNCH(k_pressure_build, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.phi_liq, (const float *)s.vel, (const float *)s.weight, s.pcoef, s.cg_r, (float)scale, s.minfrac);
    s.kernel_launches++;
    DiagPressure diag{s.pcoef};
    build_block_list<1>(s, diag);
    // the search direction is read with a one-cell halo: it must be zero outside the active blocks
    CUDA_CHECK(cudaMemsetAsync(s.cg_s, 0, sizeof(double) * (size_t)g.total, s.stream));
    CGParams P = cg_params(s, 1);
    int G = cg_grid(s);
    const float4 *coef = s.pcoef;
    cudaStream_t st = s.stream;
    int pmaxit = s.pressure_maxit * s.pressure_maxit_scale;
    CGState h;
    // phase A of an iteration: one ghost plane of the search direction each way, then the 7-point stencil
    auto apply_on = [&](CGParams &Q) {
        return [&s, &Q, G, st, coef](int parity) {
            xch_push_halo(s, s.g, Q.s, sizeof(double), 1, 0, 1);
            FLIP_LAUNCH_SYNC(k_pressure_apply, G, CG_THREADS, st, Q, coef, parity);
            s.kernel_launches++;
        };
    };
    if (solve_pressure_resident(s, pmaxit, h)) {
        // done: the whole solve ran in one persistent launch
    } else if (s.cg_variant_pressure == 1) {
        // the search direction of the stencil kernel is u = M^-1 r here: it needs the zero halo too
        CUDA_CHECK(cudaMemsetAsync(s.cg_z, 0, sizeof(double) * (size_t)g.total, s.stream));
        CGParams Pu = P;
        Pu.s = s.cg_z; Pu.q = s.cg_w;
        h = run_cg2<1>(s, P, diag, s.pressure_tol, 0.0, pmaxit, apply_on(Pu), 0);
    } else {
        h = run_cg<1>(s, P, diag, s.pressure_tol, 0.0, pmaxit, apply_on(P), 0);
    }
    // every rank gets every slab of the solution
    xch_push_gather(s, g, s.cg_x, sizeof(double), 1, 0);
    xch_barrier(s);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    xch_check(s);
    FLIP_LAUNCH(k_pressure_store, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float4 *)s.pcoef, (const double *)s.cg_x, s.pressure);
    s.kernel_launches++;
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(s.count_host, s.blk_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaMemcpyAsync(s.count_host + 1, s.unk_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaEventRecord(e1, s.stream));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1));
    s.pres_stats.iters = h.iter; s.pres_stats.converged = h.converged; s.pres_stats.resid = h.resid;
    s.pres_stats.bmax = h.bmax; s.pres_stats.skipped = (h.iter == 0 && h.converged) ? 1 : 0;
    s.pres_stats.blocks = s.count_host[0]; s.pres_stats.unknowns = s.count_host[1];
    s.pres_last_unknowns = s.count_host[1];
    s.pres_stats.ms = ms;
    if (s.verbose) {
        printf("\tpressure: %d iterations, max|r| %.3e, %s (%d active blocks, %.3f ms)\n", h.iter, h.resid,
               h.converged ? "converged" : "NOT converged", *s.count_host, ms);
    }
}

// _project (src/fluidsimulation.cpp:522-531); the weights are static and precomputed
void stage_project(Sim &s, float dt) {
    solve_pressure(s, dt);
    apply_pressure(s, dt);
    extrapolate_velocity(s);
}
