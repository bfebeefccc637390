// Geometric multigrid V-cycle used as the preconditioner of the viscosity CG.
//
// The reference preconditions with MIC(0) (src/pcgsolver/pcgsolver.h:62-214), a sequential
// triangular recurrence that still needs > 700 iterations at 256^3.  A parallel
// diagonal preconditioner needs ~5000.  This V-cycle is built for the hardware instead:
//   * every level is the SAME matrix-free coupled-face stencil, on a 2x coarser MAC grid, driven
//     by coarsened coefficient fields (volume-weighted averages of the fine cc / edge coefficients
//     and face volumes; factor dt/dx^2 -> /4);
//   * transfers follow the staggering: linear along the face normal (the coarse face coincides
//     with every second fine face), piecewise constant across it; restriction = prolongation^T / 8;
//   * smoother = damped Jacobi, the same number of sweeps before and after the coarse correction,
//     so the V-cycle is a symmetric positive definite operator and plain CG stays valid;
//   * fp32 vectors inside the cycle (the outer CG vectors stay fp64).
// Only the preconditioner changes: operator, right-hand side, stopping rule and the converged
// solution are those of cg.h / viscosity.cu.
#pragma once
#include "cg.h"

#define VMG_MAX_LEVELS 8

struct VLevel {
    Grid g;
    float *coef;     // [4T] cc, cu, cv, cw
    float *diag;     // [3T] row diagonals, 0 = not an unknown
    float *vol;      // [3T] face control-volume fractions U,V,W
    float *x[2];     // [3T] ping-pong iterate
    float *b;        // [3T] right-hand side
    float *r;        // [3T] residual
    unsigned char *wall;   // [3T] 1 = solid (Dirichlet) face
    int *blk_flag, *blk_list, *blk_count;
    bool owns;       // level 0 borrows coef/diag/vol/block list from the solver
};

struct VMG {
    int nlevels = 0;
    VLevel lv[VMG_MAX_LEVELS];
    int pre = 2;           // sweeps before = sweeps after
    int coarse_sweeps = 24;
    float omega = 0.5f;
    float alpha = 1.0f;
    int nalloc = 0;
    bool allocated = false;
};

// (A v)[comp] at cell index id for the coupled face stencil; v has 3 components of stride T
template <class TV>
FLIP_D float vmg_row(int comp, int id, int sy, int sz, size_t T, const float *__restrict__ coef,
                     float d, const TV *__restrict__ v) {
    const float *cc = coef, *cu = coef + T, *cv = coef + 2 * T, *cw = coef + 3 * T;
    const TV *u = v, *w1 = v + T, *w2 = v + 2 * T;  // U, V, W
    if (comp == 0) {
        float fR = cc[id], fL = cc[id - 1], fT = cw[id + sy], fB = cw[id], fF = cv[id + sz], fK = cv[id];
        return d * (float)u[id] - fR * (float)u[id + 1] - fL * (float)u[id - 1] - fT * (float)u[id + sy] -
               fB * (float)u[id - sy] - fF * (float)u[id + sz] - fK * (float)u[id - sz] - fT * (float)w1[id + sy] +
               fT * (float)w1[id - 1 + sy] + fB * (float)w1[id] - fB * (float)w1[id - 1] - fF * (float)w2[id + sz] +
               fF * (float)w2[id - 1 + sz] + fK * (float)w2[id] - fK * (float)w2[id - 1];
    } else if (comp == 1) {
        float fR = cw[id + 1], fL = cw[id], fT = cc[id], fB = cc[id - sy], fF = cu[id + sz], fK = cu[id];
        return d * (float)w1[id] - fR * (float)w1[id + 1] - fL * (float)w1[id - 1] - fT * (float)w1[id + sy] -
               fB * (float)w1[id - sy] - fF * (float)w1[id + sz] - fK * (float)w1[id - sz] - fR * (float)u[id + 1] +
               fR * (float)u[id + 1 - sy] + fL * (float)u[id] - fL * (float)u[id - sy] - fF * (float)w2[id + sz] +
               fF * (float)w2[id - sy + sz] + fK * (float)w2[id] - fK * (float)w2[id - sy];
    } else {
        float fR = cv[id + 1], fL = cv[id], fT = cu[id + sy], fB = cu[id], fF = cc[id], fK = cc[id - sz];
        return d * (float)w2[id] - fR * (float)w2[id + 1] - fL * (float)w2[id - 1] - fT * (float)w2[id + sy] -
               fB * (float)w2[id - sy] - fF * (float)w2[id + sz] - fK * (float)w2[id - sz] - fR * (float)u[id + 1] +
               fR * (float)u[id + 1 - sz] + fL * (float)u[id] - fL * (float)u[id - sz] - fT * (float)w1[id + sy] +
               fT * (float)w1[id + sy - sz] + fB * (float)w1[id] - fB * (float)w1[id - sz];
    }
}

struct VLevelDev {  // what the kernels need of a level
    Grid g;
    const float *coef, *diag;
    const int *blk_list, *blk_count;
};

// one axis of the prolongation: fine index n -> up to two coarse indices and weights
FLIP_D void vmg_parents(bool own, int n, int &p0, int &p1, float &w0, float &w1) {
    if (own) {
        p0 = n >> 1;
        if ((n & 1) == 0) { p1 = p0; w0 = 1.0f; w1 = 0.0f; }
        else { p1 = p0 + 1; w0 = 0.5f; w1 = 0.5f; }
    } else {
        p0 = n >> 1;
        p1 = (n & 1) ? p0 + 1 : p0 - 1;
        w0 = 0.75f; w1 = 0.25f;
    }
}

// sum of the interpolation weights of a fine face's coarse parents that are unknowns.  The
// interpolation is renormalised by it, so that faces next to the free surface (whose outer parents
// carry no equation) are extrapolated from the liquid side instead of being pulled towards 0.
FLIP_D float vmg_pnorm(int m, int i, int j, int k, const Grid &gc, const float *__restrict__ diag_c) {
    int pi[2], pj[2], pk[2];
    float wi[2], wj[2], wk[2];
    vmg_parents(m == 0, i, pi[0], pi[1], wi[0], wi[1]);
    vmg_parents(m == 1, j, pj[0], pj[1], wj[0], wj[1]);
    vmg_parents(m == 2, k, pk[0], pk[1], wk[0], wk[1]);
    const float *d = diag_c + (size_t)m * gc.total;
    float sum = 0.0f;
    for (int c2 = 0; c2 < 2; c2++)
        for (int b2 = 0; b2 < 2; b2++)
            for (int a = 0; a < 2; a++) {
                float w = wi[a] * wj[b2] * wk[c2];
                if (w == 0.0f) continue;
                int I = pi[a], J = pj[b2], K = pk[c2];
                if (I < 0 || J < 0 || K < 0 || I > gc.ni || J > gc.nj || K > gc.nk) continue;
                if (d[gidx(gc, I, J, K)] != 0.0f) sum += w;
            }
    return sum;
}

// x_out = omega * b / d   (first sweep from a zero iterate).  TB = float or double right-hand side.
template <class TB>
__global__ void __launch_bounds__(CG_THREADS) k_vmg_first(VLevelDev L, const TB *__restrict__ b, float *__restrict__ xo,
                                                           float omega, const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = L.g;
    size_t T = (size_t)g.total;
    int nb = *L.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, L.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) {
            float d = L.diag[m * T + id];
            if (d != 0.0f) xo[m * T + id] = omega * (float)b[m * T + id] / d;
        }
    }
}

// damped Jacobi sweep: x_out = x_in + omega * (b - A x_in) / d
template <class TB>
__global__ void __launch_bounds__(CG_THREADS) k_vmg_smooth(VLevelDev L, const TB *__restrict__ b, const float *__restrict__ xi,
                                                            float *__restrict__ xo, float omega, const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = L.g;
    const int sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    int nb = *L.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, L.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) {
            float d = L.diag[m * T + id];
            if (d == 0.0f) continue;
            float ax = vmg_row<float>(m, id, sy, sz, T, L.coef, d, xi);
            xo[m * T + id] = xi[m * T + id] + omega * ((float)b[m * T + id] - ax) / d;
        }
    }
}

// r = b - A x
// (scaled by 1/pnorm: the restriction is the transpose of the renormalised prolongation)
template <class TB>
__global__ void __launch_bounds__(CG_THREADS) k_vmg_residual(VLevelDev L, const TB *__restrict__ b, const float *__restrict__ x,
                                                              float *__restrict__ r, Grid gc, const float *__restrict__ diag_c,
                                                              const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = L.g;
    const int sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    int nb = *L.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, L.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) {
            float d = L.diag[m * T + id];
            if (d == 0.0f) continue;
            float nrm = vmg_pnorm(m, c.i, c.j, c.k, gc, diag_c);
            float res = (float)b[m * T + id] - vmg_row<float>(m, id, sy, sz, T, L.coef, d, x);
            r[m * T + id] = nrm > 0.0f ? res / nrm : 0.0f;
        }
    }
}

// Transfer weights.  Along a component's own axis the coarse sample coincides with fine index 2I:
// fine 2I-1, 2I, 2I+1 carry 1/2, 1, 1/2.  Across it the samples are cell-centred: coarse J sits
// between fine 2J and 2J+1, linear interpolation gives fine 2J-1, 2J, 2J+1, 2J+2 the weights
// 1/4, 3/4, 3/4, 1/4.  (Piecewise-constant transfers made the rediscretised coarse operator too
// soft by ~2x: the interpolated coarse function had twice the transverse energy.)
FLIP_D void vmg_axis(bool own, int n, int &lo, int &cnt, float w[4]) {
    if (own) { lo = 2 * n - 1; cnt = 3; w[0] = 0.5f; w[1] = 1.0f; w[2] = 0.5f; w[3] = 0.0f; }
    else { lo = 2 * n - 1; cnt = 4; w[0] = 0.25f; w[1] = 0.75f; w[2] = 0.75f; w[3] = 0.25f; }
}

// coarse b = P^T r / 8 over the coarse level's active blocks
__global__ void __launch_bounds__(CG_THREADS) k_vmg_restrict(VLevelDev C, Grid gf, const float *__restrict__ rf,
                                                              float *__restrict__ bc, const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = C.g;
    size_t Tc = (size_t)g.total, Tf = (size_t)gf.total;
    int nb = *C.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, C.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) {
            if (C.diag[m * Tc + id] == 0.0f) continue;
            const float *r = rf + m * Tf;
            int li, ci, lj, cj, lk, ck;
            float wi[4], wj[4], wk[4];
            vmg_axis(m == 0, c.i, li, ci, wi);
            vmg_axis(m == 1, c.j, lj, cj, wj);
            vmg_axis(m == 2, c.k, lk, ck, wk);
            float acc = 0.0f;
            for (int c2 = 0; c2 < ck; c2++) {
                int fk = lk + c2;
                if (fk < 0 || fk > gf.nk) continue;
                for (int b2 = 0; b2 < cj; b2++) {
                    int fj = lj + b2;
                    if (fj < 0 || fj > gf.nj) continue;
                    float wjk = wj[b2] * wk[c2];
                    for (int a = 0; a < ci; a++) {
                        int fi = li + a;
                        if (fi < 0 || fi > gf.ni) continue;
                        acc += wi[a] * wjk * r[gidx(gf, fi, fj, fk)];
                    }
                }
            }
            bc[m * Tc + id] = 0.125f * acc;
        }
    }
}

// fine x += alpha * P x_coarse over the fine level's active blocks
__global__ void __launch_bounds__(CG_THREADS) k_vmg_prolong(VLevelDev F, Grid gc, const float *__restrict__ diag_c,
                                                             const float *__restrict__ xc, float *__restrict__ xf, float alpha,
                                                             const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = F.g;
    size_t Tf = (size_t)g.total, Tc = (size_t)gc.total;
    int nb = *F.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, F.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) {
            if (F.diag[m * Tf + id] == 0.0f) continue;
            const float *x = xc + m * Tc;
            int pi[2], pj[2], pk[2];
            float wi[2], wj[2], wk[2];
            vmg_parents(m == 0, c.i, pi[0], pi[1], wi[0], wi[1]);
            vmg_parents(m == 1, c.j, pj[0], pj[1], wj[0], wj[1]);
            vmg_parents(m == 2, c.k, pk[0], pk[1], wk[0], wk[1]);
            float v = 0.0f;
            for (int c2 = 0; c2 < 2; c2++)
                for (int b2 = 0; b2 < 2; b2++)
                    for (int a = 0; a < 2; a++) {
                        float w = wi[a] * wj[b2] * wk[c2];
                        if (w == 0.0f) continue;
                        int I = pi[a], J = pj[b2], K = pk[c2];
                        if (I < 0 || J < 0 || K < 0 || I > gc.ni || J > gc.nj || K > gc.nk) continue;
                        v += w * x[gidx(gc, I, J, K)];
                    }
            float nrm = vmg_pnorm(m, c.i, c.j, c.k, gc, diag_c);
            if (nrm > 0.0f) xf[m * Tf + id] += alpha * v / nrm;
        }
    }
}

// z (double) = x (float) on the fine level's unknowns
__global__ void __launch_bounds__(CG_THREADS) k_vmg_export(VLevelDev F, const float *__restrict__ x, double *__restrict__ z,
                                                            const CGState *__restrict__ st) {
    if (st && st->done) return;
    const Grid &g = F.g;
    size_t T = (size_t)g.total;
    int nb = *F.blk_count;
    for (int bl = blockIdx.x; bl < nb; bl += gridDim.x) {
        BlockCell c = block_cell(g, F.blk_list[bl], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++) z[m * T + id] = F.diag[m * T + id] != 0.0f ? (double)x[m * T + id] : 0.0;
    }
}

// ---- coarsening of the operator -------------------------------------------------------------
FLIP_D float vmg_fetch(const Grid &gf, const float *__restrict__ f, int i, int j, int k) {
    return (i >= 0 && j >= 0 && k >= 0 && i <= gf.ni && j <= gf.nj && k <= gf.nk) ? f[gidx(gf, i, j, k)] : 0.0f;
}

// coarse coefficient fields and face volumes: weighted means over the coarse control volume.
// `node` flags which axes of the sample are node-like (weights 1/2,1,1/2 over 2I-1..2I+1) rather
// than cell-like (weights 1,1 over 2I, 2I+1).
FLIP_D float vmg_average(const Grid &gf, const float *__restrict__ f, int I, int J, int K, bool nx, bool ny, bool nz) {
    float acc = 0.0f;
    for (int a = (nx ? -1 : 0); a <= 1; a++)
        for (int b = (ny ? -1 : 0); b <= 1; b++)
            for (int c = (nz ? -1 : 0); c <= 1; c++) {
                float w = ((nx && a != 0) ? 0.5f : 1.0f) * ((ny && b != 0) ? 0.5f : 1.0f) * ((nz && c != 0) ? 0.5f : 1.0f);
                acc += w * vmg_fetch(gf, f, 2 * I + a, 2 * J + b, 2 * K + c);
            }
    return 0.125f * acc;
}

__global__ void __launch_bounds__(256) k_vmg_coarsen_coefs(Grid gc, Grid gf, const float *__restrict__ coef_f,
                                                           const float *__restrict__ vol_f, const float *__restrict__ diag_f,
                                                           float *__restrict__ coef_c, float *__restrict__ vol_c,
                                                           float *__restrict__ mask_c, const unsigned char *__restrict__ wall_f,
                                                           unsigned char *__restrict__ wall_c) {
    int I, J, K;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, gc.ni + 1, gc.nj + 1, gc.nk + 1, I, J, K)) return;
    size_t Tf = (size_t)gf.total, Tc = (size_t)gc.total;
    int id = gidx(gc, I, J, K);
    // factor dt/dx^2 is baked into the fine coefficients: dx -> 2dx divides it by 4
    coef_c[id] = 0.25f * vmg_average(gf, coef_f, I, J, K, false, false, false);                    // cc: cell
    coef_c[Tc + id] = 0.25f * vmg_average(gf, coef_f + Tf, I, J, K, false, true, true);            // cu: edge along x
    coef_c[2 * Tc + id] = 0.25f * vmg_average(gf, coef_f + 2 * Tf, I, J, K, true, false, true);    // cv: edge along y
    coef_c[3 * Tc + id] = 0.25f * vmg_average(gf, coef_f + 3 * Tf, I, J, K, true, true, false);    // cw: edge along z
    vol_c[id] = vmg_average(gf, vol_f, I, J, K, true, false, false);
    vol_c[Tc + id] = vmg_average(gf, vol_f + Tf, I, J, K, false, true, false);
    vol_c[2 * Tc + id] = vmg_average(gf, vol_f + 2 * Tf, I, J, K, false, false, true);
    // a coarse face is an unknown if any fine face it interpolates to is one
    for (int m = 0; m < 3; m++) {
        float any = 0.0f;
        unsigned char wall = 0;
        for (int a = (m == 0 ? -1 : 0); a <= 1; a++)
            for (int b = (m == 1 ? -1 : 0); b <= 1; b++)
                for (int c = (m == 2 ? -1 : 0); c <= 1; c++) {
                    int fi = 2 * I + a, fj = 2 * J + b, fk = 2 * K + c;
                    if (fi < 0 || fj < 0 || fk < 0 || fi > gf.ni || fj > gf.nj || fk > gf.nk) continue;
                    int fid = gidx(gf, fi, fj, fk);
                    if (diag_f[m * Tf + fid] != 0.0f) any = 1.0f;
                    if (wall_f[m * Tf + fid]) wall = 1;
                }
        mask_c[m * Tc + id] = any;
        wall_c[m * Tc + id] = wall;
    }
}

// Coarse faces are unknowns if a fine child is and their control volume holds real liquid mass.  A
// face that is not an unknown is either a wall (some child is solid: Dirichlet, its couplings stay
// in the neighbours' diagonals) or air.  Strain terms that touch an AIR face are removed from the
// coarse operator altogether (next kernel) instead of pinning that face to zero: rigid motions of
// a free-floating body of liquid then keep zero strain energy on every level, as they have on
// the fine one.
__global__ void __launch_bounds__(256) k_vmg_classify(Grid gc, const float *__restrict__ vol, float *__restrict__ diag,
                                                      unsigned char *__restrict__ wall, float minvol) {
    int I, J, K;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, gc.ni + 1, gc.nj + 1, gc.nk + 1, I, J, K)) return;
    size_t T = (size_t)gc.total;
    int id = gidx(gc, I, J, K);
    bool interior = I >= 1 && I < gc.ni && J >= 1 && J < gc.nj && K >= 1 && K < gc.nk;
    for (int m = 0; m < 3; m++) {
        bool unk = interior && diag[m * T + id] != 0.0f && vol[m * T + id] >= minvol;
        diag[m * T + id] = unk ? 1.0f : 0.0f;
        // faces on the rim of the index box count as walls (the domain boundary is solid)
        bool rim = !interior;
        wall[m * T + id] = (!unk && (wall[m * T + id] || rim)) ? 1 : 0;
    }
}

FLIP_D bool vmg_air(const Grid &g, const float *__restrict__ diag, const unsigned char *__restrict__ wall, int m, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i > g.ni || j > g.nj || k > g.nk) return false;  // outside: wall
    size_t o = (size_t)m * g.total + gidx(g, i, j, k);
    return diag[o] == 0.0f && !wall[o];
}

__global__ void __launch_bounds__(256) k_vmg_prune(Grid gc, float *__restrict__ coef, const float *__restrict__ diag,
                                                   const unsigned char *__restrict__ wall) {
    int I, J, K;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, gc.ni + 1, gc.nj + 1, gc.nk + 1, I, J, K)) return;
    size_t T = (size_t)gc.total;
    int id = gidx(gc, I, J, K);
    // cell term: pairs U(I),U(I+1) / V(J),V(J+1) / W(K),W(K+1)
    if (vmg_air(gc, diag, wall, 0, I, J, K) || vmg_air(gc, diag, wall, 0, I + 1, J, K) || vmg_air(gc, diag, wall, 1, I, J, K) ||
        vmg_air(gc, diag, wall, 1, I, J + 1, K) || vmg_air(gc, diag, wall, 2, I, J, K) || vmg_air(gc, diag, wall, 2, I, J, K + 1))
        coef[id] = 0.0f;
    // edge along x at node (J,K): V(I,J,K), V(I,J,K-1), W(I,J,K), W(I,J-1,K)
    if (vmg_air(gc, diag, wall, 1, I, J, K) || vmg_air(gc, diag, wall, 1, I, J, K - 1) || vmg_air(gc, diag, wall, 2, I, J, K) ||
        vmg_air(gc, diag, wall, 2, I, J - 1, K))
        coef[T + id] = 0.0f;
    // edge along y at node (I,K): U(I,J,K), U(I,J,K-1), W(I,J,K), W(I-1,J,K)
    if (vmg_air(gc, diag, wall, 0, I, J, K) || vmg_air(gc, diag, wall, 0, I, J, K - 1) || vmg_air(gc, diag, wall, 2, I, J, K) ||
        vmg_air(gc, diag, wall, 2, I - 1, J, K))
        coef[2 * T + id] = 0.0f;
    // edge along z at node (I,J): U(I,J,K), U(I,J-1,K), V(I,J,K), V(I-1,J,K)
    if (vmg_air(gc, diag, wall, 0, I, J, K) || vmg_air(gc, diag, wall, 0, I, J - 1, K) || vmg_air(gc, diag, wall, 1, I, J, K) ||
        vmg_air(gc, diag, wall, 1, I - 1, J, K))
        coef[3 * T + id] = 0.0f;
}

// coarse row diagonals (the 0/1 unknown flag arrives in diag and is overwritten)
__global__ void __launch_bounds__(256) k_vmg_coarsen_rows(Grid gc, const float *__restrict__ coef, const float *__restrict__ vol,
                                                          float *__restrict__ diag) {
    int I, J, K;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, gc.ni + 1, gc.nj + 1, gc.nk + 1, I, J, K)) return;
    size_t T = (size_t)gc.total;
    int id = gidx(gc, I, J, K), sy = SY(gc), sz = SZ(gc);
    const float *cc = coef, *cu = coef + T, *cv = coef + 2 * T, *cw = coef + 3 * T;
    float dU = 0, dV = 0, dW = 0;
    if (diag[id] != 0.0f) dU = vol[id] + cc[id] + cc[id - 1] + cw[id + sy] + cw[id] + cv[id + sz] + cv[id];
    if (diag[T + id] != 0.0f) dV = vol[T + id] + cw[id + 1] + cw[id] + cc[id] + cc[id - sy] + cu[id + sz] + cu[id];
    if (diag[2 * T + id] != 0.0f) dW = vol[2 * T + id] + cv[id + 1] + cv[id] + cu[id + sy] + cu[id] + cc[id] + cc[id - sz];
    diag[id] = dU; diag[T + id] = dV; diag[2 * T + id] = dW;
}
