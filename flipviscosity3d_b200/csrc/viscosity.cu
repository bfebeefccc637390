// Variational viscosity (Batty & Bridson 2008) as a matrix-free coupled U/V/W face stencil.
//
// Reference behaviour being reproduced (relative to /root/reference):
//   face states           src/viscositysolver.cpp:80-123   (static, see fields.cu)
//   volume fractions      src/viscositysolver.cpp:135-270
//   unknown set           src/viscositysolver.cpp:276-366
//   rows / rhs            src/viscositysolver.cpp:374-664
//   solve + acceptance    src/viscositysolver.cpp:666-690, src/pcgsolver/pcgsolver.h:241-295
//   write-back            src/viscositysolver.cpp:692-727
//
// The reference assembles a vector-of-vectors sparse matrix (<= 15 nnz/row).  Here every row is
// recomputed from four coefficient fields (2*f*mu*vol at cell centres, f*mu_bar*vol on the three
// edge families) shared by the three unknowns of a cell, plus one diagonal per face.  Off-diagonal
// masks are not needed: the search direction is kept at exactly 0 on non-unknown faces, which is
// what the reference's "drop FLUID neighbours without an index, move SOLID neighbours to the rhs"
// amounts to inside A*s.
#include "gmg.h"
#include "levelset_math.h"

// volume grid ids
enum { VC = 0, VU = 1, VV = 2, VW = 3, VEU = 4, VEV = 5, VEW = 6 };

struct VolGrid { int w, h, d; float csx, csy, csz; };

__device__ __host__ inline VolGrid vol_grid(const Grid &g, int v) {
    // dims: src/viscositysolver.h:68-76; centerStart offsets: src/viscositysolver.cpp:170-177
    VolGrid r;
    const float h = g.hdx;
    switch (v) {
        case VC:  r.w = g.ni;     r.h = g.nj;     r.d = g.nk;     r.csx = h; r.csy = h; r.csz = h; break;
        case VU:  r.w = g.ni + 1; r.h = g.nj;     r.d = g.nk;     r.csx = 0; r.csy = h; r.csz = h; break;
        case VV:  r.w = g.ni;     r.h = g.nj + 1; r.d = g.nk;     r.csx = h; r.csy = 0; r.csz = h; break;
        case VW:  r.w = g.ni;     r.h = g.nj;     r.d = g.nk + 1; r.csx = h; r.csy = h; r.csz = 0; break;
        case VEU: r.w = g.ni;     r.h = g.nj + 1; r.d = g.nk + 1; r.csx = h; r.csy = 0; r.csz = 0; break;
        case VEV: r.w = g.ni + 1; r.h = g.nj;     r.d = g.nk + 1; r.csx = 0; r.csy = h; r.csz = 0; break;
        default:  r.w = g.ni + 1; r.h = g.nj + 1; r.d = g.nk;     r.csx = 0; r.csy = 0; r.csz = h; break;
    }
    return r;
}

// liquid cells dilated twice in 6-connectivity on the (ni+1)(nj+1)(nk+1) index box
// (src/viscositysolver.cpp:138-168) = cells within L1 distance 2 of a cell with phi < 0
__global__ void __launch_bounds__(CG_THREADS) k_visc_valid(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount, const float *__restrict__ phi, unsigned char *__restrict__ vvalid) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    bool v = false;
    for (int dk = -2; dk <= 2 && !v; dk++) {
        int rj = 2 - abs(dk);
        for (int dj = -rj; dj <= rj && !v; dj++) {
            int ri = rj - abs(dj);
            for (int di = -ri; di <= ri; di++) {
                int a = i + di, b = j + dj, c = k + dk;
                if (a >= 0 && b >= 0 && c >= 0 && a < g.ni && b < g.nj && c < g.nk && phi[gidx(g, a, b, c)] < 0) {
                    v = true;
                    break;
                }
            }
        }
    }
    vvalid[gidx(g, i, j, k)] = v ? 1 : 0;
  }
}

// nodal phi of the 7 control-volume families.  The reference caches node values first-come in
// k,j,i scan order, each computed from the position arithmetic of the cell that got there first
// (src/viscositysolver.cpp:188-252); the same cell and the same float expression are used here.
__global__ void __launch_bounds__(256) k_visc_nodes(Grid g, const float *__restrict__ phi, const unsigned char *__restrict__ vvalid,
                                                    float *__restrict__ vnode) {
    int i, j, k;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, g.ni + 2, g.nj + 2, g.nk + 2, i, j, k)) return;
    int id = gidx(g, i, j, k);
    // A node value is only ever read by k_visc_volumes through a VALID cell that touches it, and a node next to a
    // valid cell is recomputed here every substep: nodes with no valid cell around them (~90 % of the 256^3 grid)
    // are skipped altogether.
    {
        bool any = false;
        for (int n = 0; n < 8 && !any; n++) {
            int ci = i - (n & 1), cj = j - ((n >> 1) & 1), ck = k - (n >> 2);
            if (ci < 0 || cj < 0 || ck < 0 || ci > g.ni || cj > g.nj || ck > g.nk) continue;
            any = vvalid[gidx(g, ci, cj, ck)] != 0;
        }
        if (!any) return;
    }
    for (int v = 0; v < 7; v++) {
        VolGrid vg = vol_grid(g, v);
        if (i > vg.w || j > vg.h || k > vg.d) continue;
        float val = 0.0f;
        bool found = false;
        for (int n = 0; n < 8 && !found; n++) {
            // scan order: smallest k first, then j, then i  ->  offsets (1,1,1),(0,1,1),(1,0,1),...
            int a = (n & 1) ? 0 : 1, b = (n & 2) ? 0 : 1, c = (n & 4) ? 0 : 1;
            int ci = i - a, cj = j - b, ck = k - c;
            if (ci < 0 || cj < 0 || ck < 0 || ci >= vg.w || cj >= vg.h || ck >= vg.d) continue;
            if (!vvalid[gidx(g, ci, cj, ck)]) continue;
            found = true;
            // centre = centerStart + GridIndexToCellCenter(ci,cj,ck); node = centre +- hdx
            float cx = vg.csx + index_to_center(ci, g.dxd);
            float cy = vg.csy + index_to_center(cj, g.dxd);
            float cz = vg.csz + index_to_center(ck, g.dxd);
            float px = cx + (a ? g.hdx : -g.hdx);
            float py = cy + (b ? g.hdx : -g.hdx);
            float pz = cz + (c ? g.hdx : -g.hdx);
            // ParticleLevelSet::trilinearInterpolate (src/particlelevelset.cpp:88-92)
            val = (float)trilinear_grid(g, phi, g.ni, g.nj, g.nk, px - g.hdx, py - g.hdx, pz - g.hdx);
        }
        vnode[(size_t)v * g.total + id] = val;
    }
}

__global__ void __launch_bounds__(CG_THREADS) k_visc_volumes(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount, const unsigned char *__restrict__ vvalid,
                                                      const float *__restrict__ vnode, float *__restrict__ vvol) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    bool valid = vvalid[id] != 0;
    for (int v = 0; v < 7; v++) {
        VolGrid vg = vol_grid(g, v);
        if (i >= vg.w || j >= vg.h || k >= vg.d) continue;
        float out = 0.0f;
        if (valid) {
            const float *nd = vnode + (size_t)v * g.total;
            float p000 = nd[id], p100 = nd[id + 1], p010 = nd[id + sy], p110 = nd[id + 1 + sy];
            float p001 = nd[id + sz], p101 = nd[id + 1 + sz], p011 = nd[id + sy + sz], p111 = nd[id + 1 + sy + sz];
            if (p000 < 0 && p001 < 0 && p010 < 0 && p011 < 0 && p100 < 0 && p101 < 0 && p110 < 0 && p111 < 0) out = 1.0f;
            else if (p000 >= 0 && p001 >= 0 && p010 >= 0 && p011 >= 0 && p100 >= 0 && p101 >= 0 && p110 >= 0 && p111 >= 0) out = 0.0f;
            else out = cube_fraction(p000, p100, p010, p110, p001, p101, p011, p111);
        }
        vvol[(size_t)v * g.total + id] = out;
    }
  }
}

void viscosity_volumes(Sim &s) {
    const Grid &g = s.g;
    long long n2 = (long long)(g.ni + 2) * (g.nj + 2) * (g.nk + 2);
    grid_list_ensure(s);
    const int *gl = s.grid_list, *gc = s.grid_count;
    FLIP_LAUNCH(k_visc_valid, list_grid(s), CG_THREADS, s.stream, g, gl, gc, (const float *)s.phi_liq, s.vvalid);
    // nodes: the index box is one wider than the blocks cover; the kernel leaves at once where no valid cell is near
    FLIP_LAUNCH(k_visc_nodes, cdiv(n2, 256), 256, s.stream, g, (const float *)s.phi_liq, (const unsigned char *)s.vvalid, s.vnode);
    FLIP_LAUNCH(k_visc_volumes, list_grid(s), CG_THREADS, s.stream, g, gl, gc, (const unsigned char *)s.vvalid, (const float *)s.vnode, s.vvol);
    s.kernel_launches += 3;
    KERNEL_CHECK();
}

// coefficient fields: cc = 2*f*mu*vol_center, ceu/cev/cew = f*mu_bar*vol_edge{U,V,W}
__global__ void __launch_bounds__(CG_THREADS) k_visc_coefs(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount, const float *__restrict__ visc, const float *__restrict__ vvol,
                                                    float *__restrict__ vcoef, float factor) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    float mu = visc[id];
    // 2 * factor * visc * vol (src/viscositysolver.cpp:422-423)
    vcoef[id] = 2 * factor * mu * vvol[VC * T + id];
    // edge means: the four cells sharing the edge (src/viscositysolver.cpp:397-413 for U rows)
    float mu_w = 0.25f * (visc[id - 1] + visc[id - 1 - sy] + mu + visc[id - sy]);
    float mu_v = 0.25f * (visc[id - 1] + visc[id - 1 - sz] + mu + visc[id - sz]);
    float mu_u = 0.25f * (visc[id - sy] + visc[id - sy - sz] + mu + visc[id - sz]);
    vcoef[1 * T + id] = factor * mu_u * vvol[VEU * T + id];
    vcoef[2 * T + id] = factor * mu_v * vvol[VEV * T + id];
    vcoef[3 * T + id] = factor * mu_w * vvol[VEW * T + id];
  }
}

// rows: unknown test, diagonal, rhs (rhs goes to r)
// vmass = the mass term the CG operator uses for every row (see k_visc_apply): the face volume itself (exact == 1, the
// default), or what is left of it in the reference's fp32 diagonal, d_fp32 - (sum of the six factors) (exact == 0).
__global__ void __launch_bounds__(CG_THREADS) k_visc_rows(Grid g, const float *__restrict__ vvol, const float *__restrict__ vcoef,
                                                   const unsigned char *__restrict__ fstate, const float *__restrict__ vel,
                                                   float *__restrict__ vdiag, double *__restrict__ rhs, float *__restrict__ vmass,
                                                   int exact, const int *__restrict__ glist, const int *__restrict__ gcount) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    const float *cc = vcoef, *cu = vcoef + T, *cv = vcoef + 2 * T, *cw = vcoef + 3 * T;
    const float *vc = vvol + VC * T, *veu = vvol + VEU * T, *vev = vvol + VEV * T, *vew = vvol + VEW * T;
    const unsigned char *su = fstate, *sv = fstate + T, *sw = fstate + 2 * T;
    const float *u = vel, *v = vel + T, *w = vel + 2 * T;
    bool interior = i >= 1 && i < g.ni && j >= 1 && j < g.nj && k >= 1 && k < g.nk;
    float dU = 0, dV = 0, dW = 0;
    float mU = 0, mV = 0, mW = 0;
    double bU = 0, bV = 0, bW = 0;
    if (interior && su[id] == 0) {
        float vol = vvol[VU * T + id];
        if (vol > 0 || vc[id] > 0 || vc[id - 1] > 0 || vew[id + sy] > 0 || vew[id] > 0 || vev[id + sz] > 0 || vev[id] > 0) {
            float fR = cc[id], fL = cc[id - 1], fT = cw[id + sy], fB = cw[id], fF = cv[id + sz], fK = cv[id];
            dU = vol + fR + fL + fT + fB + fF + fK;
            mU = exact ? vol : (float)((double)dU - ((double)fR + (double)fL + (double)fT + (double)fB + (double)fF + (double)fK));
            float r = vol * u[id];
            if (su[id + 1]) r -= -fR * u[id + 1];
            if (su[id - 1]) r -= -fL * u[id - 1];
            if (su[id + sy]) r -= -fT * u[id + sy];
            if (su[id - sy]) r -= -fB * u[id - sy];
            if (su[id + sz]) r -= -fF * u[id + sz];
            if (su[id - sz]) r -= -fK * u[id - sz];
            if (sv[id + sy]) r -= -fT * v[id + sy];
            if (sv[id - 1 + sy]) r -= fT * v[id - 1 + sy];
            if (sv[id]) r -= fB * v[id];
            if (sv[id - 1]) r -= -fB * v[id - 1];
            if (sw[id + sz]) r -= -fF * w[id + sz];
            if (sw[id - 1 + sz]) r -= fF * w[id - 1 + sz];
            if (sw[id]) r -= fK * w[id];
            if (sw[id - 1]) r -= -fK * w[id - 1];
            bU = (double)r;
        }
    }
    if (interior && sv[id] == 0) {
        float vol = vvol[VV * T + id];
        if (vol > 0 || vew[id + 1] > 0 || vew[id] > 0 || vc[id] > 0 || vc[id - sy] > 0 || veu[id + sz] > 0 || veu[id] > 0) {
            float fR = cw[id + 1], fL = cw[id], fT = cc[id], fB = cc[id - sy], fF = cu[id + sz], fK = cu[id];
            dV = vol + fR + fL + fT + fB + fF + fK;
            mV = exact ? vol : (float)((double)dV - ((double)fR + (double)fL + (double)fT + (double)fB + (double)fF + (double)fK));
            float r = vol * v[id];
            if (sv[id + 1]) r -= -fR * v[id + 1];
            if (sv[id - 1]) r -= -fL * v[id - 1];
            if (sv[id + sy]) r -= -fT * v[id + sy];
            if (sv[id - sy]) r -= -fB * v[id - sy];
            if (sv[id + sz]) r -= -fF * v[id + sz];
            if (sv[id - sz]) r -= -fK * v[id - sz];
            if (su[id + 1]) r -= -fR * u[id + 1];
            if (su[id + 1 - sy]) r -= fR * u[id + 1 - sy];
            if (su[id]) r -= fL * u[id];
            if (su[id - sy]) r -= -fL * u[id - sy];
            if (sw[id + sz]) r -= -fF * w[id + sz];
            if (sw[id - sy + sz]) r -= fF * w[id - sy + sz];
            if (sw[id]) r -= fK * w[id];
            if (sw[id - sy]) r -= -fK * w[id - sy];
            bV = (double)r;
        }
    }
    if (interior && sw[id] == 0) {
        float vol = vvol[VW * T + id];
        if (vol > 0 || vev[id + 1] > 0 || vev[id] > 0 || veu[id + sy] > 0 || veu[id] > 0 || vc[id] > 0 || vc[id - sz] > 0) {
            float fR = cv[id + 1], fL = cv[id], fT = cu[id + sy], fB = cu[id], fF = cc[id], fK = cc[id - sz];
            dW = vol + fR + fL + fT + fB + fF + fK;
            mW = exact ? vol : (float)((double)dW - ((double)fR + (double)fL + (double)fT + (double)fB + (double)fF + (double)fK));
            float r = vol * w[id];
            if (sw[id + 1]) r -= -fR * w[id + 1];
            if (sw[id - 1]) r -= -fL * w[id - 1];
            if (sw[id + sy]) r -= -fT * w[id + sy];
            if (sw[id - sy]) r -= -fB * w[id - sy];
            if (sw[id + sz]) r -= -fF * w[id + sz];
            if (sw[id - sz]) r -= -fK * w[id - sz];
            if (su[id + 1]) r -= -fR * u[id + 1];
            if (su[id + 1 - sz]) r -= fR * u[id + 1 - sz];
            if (su[id]) r -= fL * u[id];
            if (su[id - sz]) r -= -fL * u[id - sz];
            if (sv[id + sy]) r -= -fT * v[id + sy];
            if (sv[id + sy - sz]) r -= fT * v[id + sy - sz];
            if (sv[id]) r -= fB * v[id];
            if (sv[id - sz]) r -= -fB * v[id - sz];
            bW = (double)r;
        }
    }
    vdiag[id] = dU; vdiag[T + id] = dV; vdiag[2 * T + id] = dW;
    vmass[id] = mU; vmass[T + id] = mV; vmass[2 * T + id] = mW;
    rhs[id] = dU != 0.0f ? bU : 0.0;
    rhs[T + id] = dV != 0.0f ? bV : 0.0;
    rhs[2 * T + id] = dW != 0.0f ? bW : 0.0;
  }
}

// phase A: q = A s for the three face families of each cell index
//
// Every row is evaluated in DIFFERENCE form,  q = mass * s0 + sum_k f_k (s0 - s_k) + cross differences,  which is
// the same row as  d * s0 - sum_k f_k s_k - ...  with d = vol + sum_k f_k  (src/viscositysolver.cpp:429) in exact
// arithmetic.  The reference forms d in fp32; at 256^3 the six factors add up to ~2e4, ulp(2e4) = 2e-3, so the
// face-volume ("mass") term vol <= 1 of every row is rounded to within +-1e-3 and comes out NEGATIVE on thousands of thin
// free-surface faces: the assembled matrix loses its diagonal dominance there and CG needs several times more iterations
// (measured on a scipy copy of the system: 64 multigrid-PCG iterations with the exact diagonal, 307 with the fp32 one,
// Jacobi-PCG 3484 vs 5686; the reference's own MICCG(0) needs 8693 at 256^3, far beyond its 700-iteration cap).  The
// difference form takes the mass term as a separate field, `vmass` (k_visc_rows):
//   viscosity_operator = 0 (default)  vmass = vol, the exact row;
//   viscosity_operator = 1            vmass = d_fp32 - sum of the factors: bit for bit the reference's rounded row, for
//                                     strict parity runs (tests/test_parity_regime.py).
// The two differ by the reference's rounding noise only (<= 1e-7 relative per diagonal entry), which in the stiff regime
// moves the SOLUTION by ~3e-5 (measured: profiles/r2_parity_*.json).
__global__ void __launch_bounds__(CG_THREADS) k_visc_apply(CGParams P, const float *__restrict__ vcoef,
                                                           const float *__restrict__ vdiag, const float *__restrict__ vmass, int parity) {
    __shared__ double sm[CG_THREADS / 32];
    if (P.st[parity].done) return;
    if (!xch_enter(P.X)) return;
    const Grid &g = P.g;
    const int sy = SY(g), sz = SZ(g);
    const size_t T = (size_t)g.total;
    const float *__restrict__ cc = vcoef, *__restrict__ cu = vcoef + T, *__restrict__ cv = vcoef + 2 * T,
                *__restrict__ cw = vcoef + 3 * T;
    const double *__restrict__ su = P.s, *__restrict__ sv = P.s + T, *__restrict__ sw = P.s + 2 * T;
    int nc = *P.cell_count;
    double sq = 0.0;
    for (int qq = blockIdx.x * CG_THREADS + threadIdx.x; qq < nc; qq += gridDim.x * CG_THREADS) {
        int id = P.cell_list[qq];
        // Every value the three rows of this cell index need is loaded once, unconditionally and up
        // front (27 fp64 + 13 fp32 + 3 diagonals): the loads are independent, so they are all in
        // flight together, and values shared between the U, V and W rows are not re-requested from
        // L1 (the ncu capture of round 1 showed this kernel at 74 % of L1 throughput with the loads
        // nested inside the per-row branches).
        const float dU = vdiag[id], dV = vdiag[T + id], dW = vdiag[2 * T + id];
        const double u0 = su[id], u_xp = su[id + 1], u_xm = su[id - 1], u_yp = su[id + sy], u_ym = su[id - sy],
                     u_zp = su[id + sz], u_zm = su[id - sz], u_xp_ym = su[id + 1 - sy], u_xp_zm = su[id + 1 - sz];
        const double v0 = sv[id], v_xp = sv[id + 1], v_xm = sv[id - 1], v_yp = sv[id + sy], v_ym = sv[id - sy],
                     v_zp = sv[id + sz], v_zm = sv[id - sz], v_xm_yp = sv[id - 1 + sy], v_yp_zm = sv[id + sy - sz];
        const double w0 = sw[id], w_xp = sw[id + 1], w_xm = sw[id - 1], w_yp = sw[id + sy], w_ym = sw[id - sy],
                     w_zp = sw[id + sz], w_zm = sw[id - sz], w_xm_zp = sw[id - 1 + sz], w_ym_zp = sw[id - sy + sz];
        const double c0 = cc[id], c_xm = cc[id - 1], c_ym = cc[id - sy], c_zm = cc[id - sz];
        const double eu0 = cu[id], eu_yp = cu[id + sy], eu_zp = cu[id + sz];
        const double ev0 = cv[id], ev_xp = cv[id + 1], ev_zp = cv[id + sz];
        const double ew0 = cw[id], ew_xp = cw[id + 1], ew_yp = cw[id + sy];
        if (dU != 0.0f) {
            const double fR = c0, fL = c_xm, fT = ew_yp, fB = ew0, fF = ev_zp, fK = ev0;
            double q = (double)vmass[id] * u0 + fR * (u0 - u_xp) + fL * (u0 - u_xm) + fT * (u0 - u_yp) + fB * (u0 - u_ym) +
                       fF * (u0 - u_zp) + fK * (u0 - u_zm) - fT * (v_yp - v_xm_yp) + fB * (v0 - v_xm) - fF * (w_zp - w_xm_zp) +
                       fK * (w0 - w_xm);
            P.q[id] = q;
            sq += u0 * q;
        }
        if (dV != 0.0f) {
            const double fR = ew_xp, fL = ew0, fT = c0, fB = c_ym, fF = eu_zp, fK = eu0;
            double q = (double)vmass[T + id] * v0 + fR * (v0 - v_xp) + fL * (v0 - v_xm) + fT * (v0 - v_yp) + fB * (v0 - v_ym) +
                       fF * (v0 - v_zp) + fK * (v0 - v_zm) - fR * (u_xp - u_xp_ym) + fL * (u0 - u_ym) - fF * (w_zp - w_ym_zp) +
                       fK * (w0 - w_ym);
            P.q[T + id] = q;
            sq += v0 * q;
        }
        if (dW != 0.0f) {
            const double fR = ev_xp, fL = ev0, fT = eu_yp, fB = eu0, fF = c0, fK = c_zm;
            double q = (double)vmass[2 * T + id] * w0 + fR * (w0 - w_xp) + fL * (w0 - w_xm) + fT * (w0 - w_yp) + fB * (w0 - w_ym) +
                       fF * (w0 - w_zp) + fK * (w0 - w_zm) - fR * (u_xp - u_xp_zm) + fL * (u0 - u_zm) - fT * (v_yp - v_yp_zm) +
                       fB * (v0 - v_zm);
            P.q[2 * T + id] = q;
            sq += w0 * q;
        }
    }
    sq = cta_reduce<false>(sq, sm);
    if (threadIdx.x == 0) PART_STORE(P, 0, sq);
    PART_LEAVE(P);
}

static void gmg_free(Sim &s);
void viscosity_free(Sim &s) { gmg_free(s); }

// ------------------------------------------------------------------------------------------
// Galerkin multigrid hierarchy (gmg.h): allocation, per-solve set-up, V-cycle launcher
// ------------------------------------------------------------------------------------------
static GMG *gmg_get(Sim &s) {
    if (s.gmg) return (GMG *)s.gmg;
    GMG *M = new GMG();
    s.gmg = M;
    CUDA_CHECK(cudaMallocHost((void **)&M->count_host, 4 * sizeof(int)));
    heap_alloc(s, M->dense, (size_t)GMG_DENSE_MAX * GMG_DENSE_MAX);
    heap_alloc(s, M->Ainv, (size_t)GMG_DENSE_MAX * GMG_DENSE_MAX);
    Grid g = s.g;
    for (int l = 0; l < GMG_MAX_LEVELS; l++) {
        GLevel &L = M->lv[l];
        L.g = g;
        size_t T = (size_t)g.total;
        heap_alloc(s, L.x[0], 3 * T);
        heap_alloc(s, L.x[1], 3 * T);
        heap_alloc(s, L.r, 3 * T);
        heap_alloc(s, L.pn, 3 * T);
        heap_alloc(s, L.rng, 4);
        if (l == 0) {
            L.diag = s.vdiag; L.blk_flag = s.blk_flag; L.blk_list = s.blk_list; L.blk_count = s.blk_count;
            L.owns = false;
        } else {
            heap_alloc(s, L.diag, 3 * T); heap_alloc(s, L.b, 3 * T);
            heap_alloc(s, L.blk_flag, (size_t)g.nblocks); heap_alloc(s, L.blk_list, (size_t)g.nblocks);
            heap_alloc(s, L.blk_count, 1);
            heap_alloc(s, L.rows, 3 * T); heap_alloc(s, L.rowmap, 3 * T); heap_alloc(s, L.nrows_dev, 1);
            L.ntiles = FLIP_B * g.nbz * g.nby * g.nbx;
            heap_alloc(s, L.tile_off, (size_t)L.ntiles + 1);
            heap_alloc(s, L.scan_tmp, (size_t)L.ntiles / 2048 + 2);
            heap_alloc(s, L.offs, 3 * GMG_STRIDE);
            std::vector<int> offs(3 * GMG_STRIDE, 0);
            for (int m = 0; m < 3; m++)
                for (int slot = 0; slot < GMG_SLOTS; slot++) offs[m * GMG_STRIDE + slot] = gmg_slot_offset(g, m, slot);
            CUDA_CHECK(cudaMemcpyAsync(L.offs, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice, s.stream));
            if (l == 1) {   // x-groups: the restriction from level 0 is the expensive one (gmg.h k_gmg_restrict_x)
                heap_alloc(s, L.groups, 3 * T / 8 + 8);
                heap_alloc(s, L.gtile_off, (size_t)L.ntiles + 1);
                heap_alloc(s, L.grng, 4);
                // compact rows (gmg.h k_gmg_compact_rows); the row array itself follows S
                heap_alloc(s, L.coffs, 3 * GMG_CSTRIDE);
                heap_alloc(s, L.cslot, 3 * GMG_CSTRIDE);
                heap_alloc(s, L.cmeta, 4);
                heap_alloc(s, L.cmask, 3 * 8);
            }
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
            L.owns = true;
        }
        M->nalloc = l + 1;
        int mn = g.ni < g.nj ? g.ni : g.nj;
        mn = mn < g.nk ? mn : g.nk;
        if (mn <= 4) break;
        g = make_grid((g.ni + 1) / 2, (g.nj + 1) / 2, (g.nk + 1) / 2, g.dx * 2.0f);
    }
    return M;
}

static void gmg_free(Sim &s) {
    GMG *M = (GMG *)s.gmg;
    if (!M) return;
    // the buffers live in the handle's heap, which is released as a whole (sim_free)
    cudaFreeHost(M->count_host);
#ifndef FLIP_CPU_EMU
    if (M->exec) cudaGraphExecDestroy((cudaGraphExec_t)M->exec);
#endif
    delete M;
    s.gmg = nullptr;
}

// device view of a level; `own`: restricted to this rank's slab of rows and hand-shaking with the other ranks
static GLevelDev gmg_dev(Sim &s, const GLevel &L, bool own) {
    GLevelDev d;
    d.g = L.g; d.diag = L.diag; d.rows = L.rows; d.nrows = L.nrows_dev; d.S = L.S; d.wj = L.wj; d.offs = L.offs;
    const bool compact = s.mg_compact && L.cmeta != nullptr && L.Sc != nullptr;
    d.Sc = compact ? L.Sc : nullptr; d.coffs = compact ? L.coffs : nullptr; d.cmeta = compact ? L.cmeta : nullptr;
    d.rng = own ? L.rng : L.rng + 2;
    d.groups = L.groups;
    d.grng = L.grng ? (own ? L.grng : L.grng + 2) : nullptr;
    d.X = xch_of(s);
    if (!own) d.X.nranks = 1;
    return d;
}
static int gmg_grid(const Sim &s, const GLevel &L) {
    int G = cg_grid(s);
    return L.g.nblocks < G ? L.g.nblocks : G;
}
static int gmg_row_grid(const Sim &s, const GLevel &L) {
    // one warp per row, 8 warps per CTA.  Sized from the row CAPACITY (which only changes when the level is
    // re-allocated), not the row count of this solve, so the captured launch sequence stays valid across solves.
    long long G = cdiv((long long)L.cap, 8), cap = s.num_sms * 8;
    return (int)(G < 1 ? 1 : (G > cap ? cap : G));
}

// Levels whose rows are cut into k-slabs when the handle is sharded: the sweeps of level 1 and the Galerkin products and
// restrictions of levels 1 and 2 (their cuts are whole planes, xch.h).  Deeper levels are small and latency bound: every
// rank runs them in full on its own copy.
#define GMG_SHARD_SWEEP_LEVELS 1
#define GMG_SHARD_BUILD_LEVELS 2

// Galerkin operators for this solve: unknown flags, row lists, transfer normalisers, A_c = P^T A P / 8
static void gmg_build(Sim &s, GMG &M) {
    M.pre = s.mg_sweeps < 1 ? 1 : s.mg_sweeps;
    for (int l = 0; l < GMG_MAX_LEVELS; l++) M.pre_l[l] = M.pre;
    if (s.mg_sweeps_l0 > 0) M.pre_l[0] = s.mg_sweeps_l0;
    if (s.mg_sweeps_l1 > 0) M.pre_l[1] = s.mg_sweeps_l1;
    M.coarse_sweeps = s.mg_coarse_sweeps;
    M.omega = s.mg_omega;
    int want = s.mg_levels < M.nalloc ? (s.mg_levels < 1 ? 1 : s.mg_levels) : M.nalloc;
    M.nlevels = 1;
    M.dense_last = false;
    const Cuts *cuts = xch_cuts(s);
    for (int l = 0; l < want; l++) {
        GLevel &L = M.lv[l];
        size_t T = (size_t)L.g.total;
        // vectors are read with a halo: zero outside the unknowns of THIS solve
        CUDA_CHECK(cudaMemsetAsync(L.x[0], 0, 3 * T * sizeof(float), s.stream));
        CUDA_CHECK(cudaMemsetAsync(L.x[1], 0, 3 * T * sizeof(float), s.stream));
        CUDA_CHECK(cudaMemsetAsync(L.r, 0, 3 * T * sizeof(float), s.stream));
        CUDA_CHECK(cudaMemsetAsync(L.pn, 0, 3 * T * sizeof(float), s.stream));
        if (l == 0) continue;
        GLevel &F = M.lv[l - 1];
        const bool shard = cuts && l <= GMG_SHARD_BUILD_LEVELS;
        CUDA_CHECK(cudaMemsetAsync(L.b, 0, 3 * T * sizeof(float), s.stream));
        CUDA_CHECK(cudaMemsetAsync(L.rowmap, 0xFF, 3 * T * sizeof(int), s.stream));
        CUDA_CHECK(cudaMemsetAsync(L.tile_off, 0, ((size_t)L.ntiles + 1) * sizeof(int), s.stream));
        long long n = (long long)(L.g.ni + 1) * (L.g.nj + 1) * (L.g.nk + 1);
        FLIP_LAUNCH(k_gmg_flags, cdiv(n, 256), 256, s.stream, L.g, F.g, (const float *)F.diag, L.diag);
        DiagViscosity d{L.diag, L.g.total};
        build_block_list_on<3>(s, L.g, d, L.blk_flag, L.blk_list, L.blk_count);
        int G = gmg_grid(s, L);
        const bool xg = L.groups != nullptr;
        if (xg) CUDA_CHECK(cudaMemsetAsync(L.gtile_off, 0, ((size_t)L.ntiles + 1) * sizeof(int), s.stream));
        FLIP_LAUNCH_SYNC(k_gmg_row_counts, G, CG_THREADS, s.stream, L.g, (const int *)L.blk_list, (const int *)L.blk_count,
                         (const float *)L.diag, L.tile_off, xg ? L.gtile_off : (int *)nullptr);
        exclusive_scan(s, L.tile_off, L.tile_off, L.scan_tmp, L.ntiles);
        if (xg) exclusive_scan(s, L.gtile_off, L.gtile_off, L.scan_tmp, L.ntiles);
        FLIP_LAUNCH_SYNC(k_gmg_row_fill, G, CG_THREADS, s.stream, L.g, (const int *)L.blk_list, (const int *)L.blk_count,
                         (const float *)L.diag, (const int *)L.tile_off, L.rows, L.rowmap, (const int *)L.gtile_off, L.groups);
        FLIP_LAUNCH(k_gmg_ranges, 1, 32, s.stream, (const int *)L.tile_off, L.ntiles, L.g.nbx * L.g.nby,
                    l < XCH_LEVELS ? cuts : (const Cuts *)nullptr, l, xch_rank(s), L.rng, L.nrows_dev, (const int *)L.gtile_off, L.grng);
        s.kernel_launches += 4;
        CUDA_CHECK(cudaMemcpyAsync(M.count_host, L.nrows_dev, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        L.nrows = M.count_host[0];
        if (L.nrows == 0) break;
        if ((size_t)L.nrows > L.cap) {
            heap_free(s, L.S); heap_free(s, L.wj);
            if (L.cmeta) heap_free(s, L.Sc);
            L.cap = (size_t)L.nrows + (size_t)L.nrows / 4 + 1024;
            heap_alloc(s, L.S, L.cap * GMG_STRIDE);   // zero filled: padding slots stay 0
            heap_alloc(s, L.wj, L.cap);
            if (L.cmeta) heap_alloc(s, L.Sc, L.cap * GMG_CSTRIDE);
            // every rank took this branch (same row count).  The zero fill runs on each rank's own stream: no peer may
            // push rows into the new array before it is done.  A barrier only proves that the peers finished the kernel
            // BEFORE their previous one, so it takes two to order their fills before this rank's push.
            if (shard) { xch_barrier(s); xch_barrier(s); }
        }
        // transfer normaliser of the fine level, then the Galerkin product
        FLIP_LAUNCH(k_gmg_pnorm, gmg_grid(s, F), CG_THREADS, s.stream, F.g, (const int *)F.blk_list, (const int *)F.blk_count,
                    (const float *)F.diag, L.g, (const float *)L.diag, F.pn);
        int GB = cdiv(3LL * L.nrows, GMG_BUILD_WARPS);
        const int *rng = shard ? L.rng : L.rng + 2;
        if (l == 1) {
            auto kb = s.mg_build ? &k_gmg_build_g<true> : &k_gmg_build<true>;
            FLIP_LAUNCH_SYNC(kb, GB, 32 * GMG_BUILD_WARPS, s.stream, L.g, F.g, (const int *)L.rows, rng, (const float *)L.diag, L.S,
                        (const float *)F.diag, (const float *)F.pn, (const float *)s.vcoef, (const int *)nullptr,
                        (const float *)s.vvol, 0);
        } else {
            auto kb = s.mg_build ? &k_gmg_build_g<false> : &k_gmg_build<false>;
            FLIP_LAUNCH_SYNC(kb, GB, 32 * GMG_BUILD_WARPS, s.stream, L.g, F.g, (const int *)L.rows, rng, (const float *)L.diag, L.S,
                        (const float *)F.diag, (const float *)F.pn, (const float *)nullptr, (const int *)F.rowmap,
                        (const float *)F.S, F.nrows);
        }
        if (shard) {
            // every rank computed the rows of its own slab: deliver them to everybody (the next level's products and the
            // mirror pass read rows of any slab), then make sure everybody's rows have arrived
            xch_push_rows(s, L.rng, L.S, GMG_STRIDE * sizeof(float));
            xch_barrier(s);
        }
        FLIP_LAUNCH(k_gmg_mirror, cdiv(160LL * L.nrows, 256), 256, s.stream, L.g, (const int *)L.rows, L.nrows, (const int *)L.rowmap, L.S);
        FLIP_LAUNCH_SYNC(k_gmg_diag, cdiv(L.nrows, 8), 256, s.stream, L.g, (const int *)L.rows, L.nrows, (const float *)L.S, L.diag, L.wj, M.omega);
        s.kernel_launches += 4;
        M.nlevels = l + 1;
        if (s.mg_dense && L.nrows <= (s.mg_dense_rows < GMG_DENSE_MAX ? s.mg_dense_rows : GMG_DENSE_MAX)) {
            // small enough for an exact solve: this is the last level
            GLevelDev D = gmg_dev(s, L, false);
            FLIP_LAUNCH_SYNC(k_gmg_dense_inverse, 1, 1024, s.stream, D, (const int *)L.rowmap, M.dense, M.Ainv);
            s.kernel_launches++;
            M.dense_last = true;
            break;
        }
    }
    // compact rows of the first explicit level for the sweeps (gmg.h): after every reader of the full rows above, on every
    // rank over all rows (the ranks hold identical rows by now; the pass is ~0.1 ms)
    if (M.nlevels >= 2 && M.lv[1].cmeta) {
        GLevel &L = M.lv[1];
        CUDA_CHECK(cudaMemsetAsync(L.cmeta, 0, 4 * sizeof(int), s.stream));
        if (s.mg_compact && L.Sc) {
            CUDA_CHECK(cudaMemsetAsync(L.cmask, 0, 3 * 8 * sizeof(unsigned), s.stream));
            int GR = gmg_row_grid(s, L);
            FLIP_LAUNCH_SYNC(k_gmg_slot_mask, GR, 256, s.stream, (const int *)L.rows, (const int *)L.nrows_dev, L.g.total, (const float *)L.S, L.cmask);
            FLIP_LAUNCH_SYNC(k_gmg_compact_table, 1, 96, s.stream, (const unsigned *)L.cmask, (const int *)L.offs, L.coffs, L.cslot, L.cmeta);
            FLIP_LAUNCH_SYNC(k_gmg_compact_rows, GR, 256, s.stream, (const int *)L.rows, (const int *)L.nrows_dev, L.g.total, (const float *)L.S,
                             (const int *)L.cslot, (const int *)L.cmeta, L.Sc);
            s.kernel_launches += 3;
        }
    }
    KERNEL_CHECK();
}

// z = Vcycle(r): r is the CG residual (fp64) on level 0.  All grid sizes depend only on what the host
// knows after gmg_build (row capacities), every other argument is a pointer into the hierarchy, so the whole
// launch sequence can be captured into a CUDA graph.  On a sharded handle levels 0 and 1 work on this rank's slab:
// every sweep is followed by a push of the ghost planes its successor reads (1 plane of level 0, 2 planes of level 1
// - the Galerkin window); the restriction to level 2 is cut by rows and gathered, deeper levels run replicated.
static void gmg_vcycle(Sim &s, GMG &M, const double *r_in, double *z_out, const CGState *st) {
    const float w = M.omega;
    int cur[GMG_MAX_LEVELS];
    auto l0_first = &k_gmg0_sweep<0>; auto l0_smooth = &k_gmg0_sweep<1>; auto l0_resid = &k_gmg0_sweep<2>; auto l0_last = &k_gmg0_sweep<3>;
#ifndef FLIP_CPU_EMU
    // coefficient rows staged by TMA bulk copies (mg_tma, default) or read through L1 like the other operands
    auto sweep1 = s.mg_tma ? &k_gmg_sweep_tma<1> : &k_gmg_sweep<1>;
    auto sweep2 = s.mg_tma ? &k_gmg_sweep_tma<2> : &k_gmg_sweep<2>;
#else
    auto sweep1 = &k_gmg_sweep<1>; auto sweep2 = &k_gmg_sweep<2>;
#endif
    const int last = M.nlevels - 1;
    const bool sh = xch_cuts(s) != nullptr;
    GLevel &L0 = M.lv[0];
    G0Params P0;
    P0.g = L0.g; P0.cell_list = s.cell_list; P0.cell_count = s.cell_count; P0.coef = s.vcoef; P0.diag = L0.diag; P0.pn = L0.pn; P0.vol = s.vvol;
    P0.X = xch_of(s);
    const int G0 = s.num_sms * 8;
    const float *nof = nullptr;
    auto halo0 = [&](float *x) { xch_push_halo(s, L0.g, x, sizeof(float), 3, 0, 1); };
    // level 0, downstroke
    FLIP_LAUNCH(l0_first, G0, 256, s.stream, P0, r_in, nof, L0.x[0], (double *)nullptr, w, st);
    cur[0] = 0;
    halo0(L0.x[0]);
    if (last == 0) {
        for (int k = 1; k < 2 * M.pre_l[0] - 1; k++) {
            FLIP_LAUNCH_X(sh, l0_smooth, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], L0.x[cur[0] ^ 1], (double *)nullptr, w, st);
            cur[0] ^= 1;
            halo0(L0.x[cur[0]]);
        }
        FLIP_LAUNCH_X(sh, l0_last, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], (float *)nullptr, z_out, w, st);
        s.kernel_launches += 2 * M.pre_l[0];
        return;
    }
    for (int k = 1; k < M.pre_l[0]; k++) {
        FLIP_LAUNCH_X(sh, l0_smooth, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], L0.x[cur[0] ^ 1], (double *)nullptr, w, st);
        cur[0] ^= 1;
        halo0(L0.x[cur[0]]);
    }
    FLIP_LAUNCH_X(sh, l0_resid, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], L0.r, (double *)nullptr, w, st);
    if (sh) xch_push_halo(s, L0.g, L0.r, sizeof(float), 3, 0, 2);
    s.kernel_launches += M.pre_l[0] + 1;
    // explicit levels, downstroke
    for (int l = 1; l <= last; l++) {
        GLevel &L = M.lv[l];
        // sweeps on this rank's rows only (never the dense last level: its solve needs the whole right-hand side)
        const bool own = sh && l <= GMG_SHARD_SWEEP_LEVELS && !(l == last && M.dense_last);
        // the restriction INTO level l is cut by rows while level l-1 is (the rows' children live in this rank's slab)
        const bool own_restrict = sh && l <= GMG_SHARD_SWEEP_LEVELS + 1 && l < XCH_LEVELS;
        GLevelDev D = gmg_dev(s, L, own), Dr = gmg_dev(s, L, own_restrict);
        int GR = gmg_row_grid(s, L);
        const bool xg = s.mg_xgroup && L.groups != nullptr;   // first explicit level: restriction over x-groups (gmg.h)
        if (xg) FLIP_LAUNCH_SYNC(k_gmg_restrict_x, GR, 256, s.stream, Dr, M.lv[l - 1].g, (const float *)M.lv[l - 1].r, L.b, L.x[0], st);
        else FLIP_LAUNCH_SYNC(k_gmg_restrict_first, GR, 256, s.stream, Dr, M.lv[l - 1].g, (const float *)M.lv[l - 1].r, L.b, L.x[0], st);
        cur[l] = 0;
        if (own) xch_push_halo(s, L.g, L.x[0], sizeof(float), 3, l, 2);
        else if (own_restrict) {
            // first replicated level: gather the restricted right-hand side and first iterate from all slabs
            xch_push_gather(s, L.g, L.b, sizeof(float), 3, l);
            xch_push_gather(s, L.g, L.x[0], sizeof(float), 3, l);
            xch_barrier(s);
        }
        if (l == last && M.dense_last) {
            FLIP_LAUNCH_SYNC(k_gmg_dense_apply, 8, 256, s.stream, D, (const float *)M.Ainv, (const float *)L.b, L.x[0], st);
            s.kernel_launches += 2;
            continue;
        }
        int sweeps = l == last ? 1 + M.coarse_sweeps : M.pre_l[l];
        for (int k = 1; k < sweeps; k++) {
            FLIP_LAUNCH_SYNC(sweep1, GR, 256, s.stream, D, (const float *)L.b, (const float *)L.x[cur[l]], L.x[cur[l] ^ 1], nof, w, st);
            cur[l] ^= 1;
            if (own) xch_push_halo(s, L.g, L.x[cur[l]], sizeof(float), 3, l, 2);
        }
        if (l < last) {
            FLIP_LAUNCH_SYNC(sweep2, GR, 256, s.stream, D, (const float *)L.b, (const float *)L.x[cur[l]], L.r, (const float *)L.pn, w, st);
            if (own) xch_push_halo(s, L.g, L.r, sizeof(float), 3, l, 2);
        }
        s.kernel_launches += sweeps + (l < last ? 1 : 0);
    }
    // upstroke
    for (int l = last - 1; l >= 1; l--) {
        GLevel &L = M.lv[l];
        const bool own = sh && l <= GMG_SHARD_SWEEP_LEVELS;   // l < last here: never the dense level
        GLevelDev D = gmg_dev(s, L, own);
        int GR = gmg_row_grid(s, L), GT = cdiv((long long)L.cap, 256) < s.num_sms * 8 ? cdiv((long long)L.cap, 256) : s.num_sms * 8;
        // the coarser level is replicated (or, for a sharded one, its ghost planes were pushed after its last sweep)
        FLIP_LAUNCH(k_gmg_prolong, GT, 256, s.stream, D, (const float *)L.pn, M.lv[l + 1].g, (const float *)M.lv[l + 1].x[cur[l + 1]], L.x[cur[l]], st);
        if (own) xch_push_halo(s, L.g, L.x[cur[l]], sizeof(float), 3, l, 2);
        for (int k = 0; k < M.pre_l[l]; k++) {
            FLIP_LAUNCH_SYNC(sweep1, GR, 256, s.stream, D, (const float *)L.b, (const float *)L.x[cur[l]], L.x[cur[l] ^ 1], nof, w, st);
            cur[l] ^= 1;
            if (own) xch_push_halo(s, L.g, L.x[cur[l]], sizeof(float), 3, l, 2);
        }
        s.kernel_launches += 1 + M.pre_l[l];
    }
    FLIP_LAUNCH_X(sh, k_gmg0_prolong, G0, 256, s.stream, P0, M.lv[1].g, (const float *)M.lv[1].x[cur[1]], L0.x[cur[0]], st);
    halo0(L0.x[cur[0]]);
    for (int k = 0; k < M.pre_l[0] - 1; k++) {
        FLIP_LAUNCH_X(sh, l0_smooth, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], L0.x[cur[0] ^ 1], (double *)nullptr, w, st);
        cur[0] ^= 1;
        halo0(L0.x[cur[0]]);
    }
    FLIP_LAUNCH_X(sh, l0_last, G0, 256, s.stream, P0, r_in, (const float *)L0.x[cur[0]], (float *)nullptr, z_out, w, st);
    s.kernel_launches += 1 + M.pre_l[0];
}

// Multigrid-preconditioned CG: the iteration chunk (stencil apply, update, V-cycle, dot, direction: ~70 launches per
// iteration on one GPU, ~100 with the exchange pushes of a sharded handle) is replayed from a CUDA graph that is
// re-captured only when a level is re-allocated, a parameter changes or the exchange set-up changes.
template <class ApplyFn>
static CGState run_cg_gmg(Sim &s, GMG &M, CGParams P, DiagViscosity diag, double tol_rel, int maxit, ApplyFn apply,
                          const float *guess) {
    int G = cg_grid(s);
    auto kinit = &k_cg_init_mg<3, DiagViscosity, false>;
    auto kinit_keep = &k_cg_init_mg<3, DiagViscosity, true>;
    auto kstart = &k_cg_start_mg<3>;
    auto kdot = &k_cg_dot<3, DiagViscosity>;
    auto kupdate = &k_cg_update<3, DiagViscosity, true>;
    auto kdir = &k_cg_direction<3, DiagViscosity, true>;
    if (guess) {
        // warm start from the current velocity: r0 = b - A x0; the stopping rule stays max|r| <= tol * max|b|
        auto kguess = &k_cg_guess<3, DiagViscosity>;
        auto kres = &k_cg_guess_residual<3, DiagViscosity>;
        auto kbmax = &k_cg_bmax<3, DiagViscosity>;
        FLIP_LAUNCH_SYNC(kbmax, G, CG_THREADS, s.stream, P, diag);
        FLIP_LAUNCH(kguess, G, CG_THREADS, s.stream, P, diag, guess);
        CUDA_CHECK(cudaMemsetAsync(s.cgst, 0, 2 * sizeof(CGState), s.stream));   // the stencil kernel tests st[0].done
        apply(0);
        FLIP_LAUNCH(kres, G, CG_THREADS, s.stream, P, diag);
        FLIP_LAUNCH_SYNC(kinit_keep, G, CG_THREADS, s.stream, P, diag);
        s.kernel_launches += 4;
    } else
        FLIP_LAUNCH_SYNC(kinit, G, CG_THREADS, s.stream, P, diag);
    gmg_vcycle(s, M, (const double *)P.r, P.z, nullptr);
    FLIP_LAUNCH_SYNC(kdot, G, CG_THREADS, s.stream, P, diag, -1);
    FLIP_LAUNCH(kstart, G, CG_THREADS, s.stream, P);
    FLIP_LAUNCH_SYNC(k_cg_begin, 1, CG_THREADS, s.stream, P, G, 0.0, tol_rel, maxit, guess ? 1 : 0);
    s.kernel_launches += 4;
    KERNEL_CHECK();
    int chunk = s.mg_chunk < 2 ? 2 : (s.mg_chunk & ~1);
    long long per_chunk = 0;
    auto launch_chunk = [&]() {
        long long before = s.kernel_launches;
        for (int it = 0; it < chunk; it++) {
            int parity = it & 1;
            apply(parity);
            FLIP_LAUNCH_SYNC(kupdate, G, CG_THREADS, s.stream, P, diag, parity);
            gmg_vcycle(s, M, (const double *)P.r, P.z, (const CGState *)(P.st + parity));
            FLIP_LAUNCH_SYNC(kdot, G, CG_THREADS, s.stream, P, diag, parity);
            FLIP_LAUNCH_SYNC(kdir, G, CG_THREADS, s.stream, P, diag, parity);
            s.kernel_launches += 3;
        }
        per_chunk = s.kernel_launches - before;
    };
#ifndef FLIP_CPU_EMU
    // The graph is keyed by everything the launch sequence depends on (level count, sweep counts, every pointer
    // and capacity-derived grid size, the exchange epoch); row counts and slab cuts are read on the device.
    cudaGraphExec_t exec = nullptr;
    if (s.use_graphs) {
        unsigned long long sig = 1469598103934665603ull;
        auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
        mix((unsigned long long)M.nlevels); mix((unsigned long long)M.dense_last); mix((unsigned long long)chunk); mix((unsigned long long)M.coarse_sweeps);
        mix((unsigned long long)(M.omega * 1e6f)); mix((unsigned long long)P.flexible); mix((unsigned long long)G);
        mix(s.xch_epoch); mix((unsigned long long)s.sharded); mix((unsigned long long)s.mg_tma); mix((unsigned long long)s.mg_xgroup); mix((unsigned long long)s.mg_compact);
        for (int l = 0; l < M.nlevels; l++) {
            const GLevel &L = M.lv[l];
            mix((unsigned long long)M.pre_l[l]); mix((unsigned long long)L.cap); mix((unsigned long long)(size_t)L.S);
            mix((unsigned long long)(size_t)L.wj); mix((unsigned long long)(size_t)L.x[0]); mix((unsigned long long)(size_t)L.Sc);
        }
        if (M.exec && M.exec_sig != sig) { cudaGraphExecDestroy((cudaGraphExec_t)M.exec); M.exec = nullptr; }
        if (!M.exec) {
            cudaGraph_t graph = nullptr;
            long long keep = s.kernel_launches;
            CUDA_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
            launch_chunk();
            CUDA_CHECK(cudaStreamEndCapture(s.stream, &graph));
            s.kernel_launches = keep;   // captured, not launched
            cudaGraphExec_t e = nullptr;
            CUDA_CHECK(cudaGraphInstantiate(&e, graph, 0));
            CUDA_CHECK(cudaGraphDestroy(graph));
            M.exec = (void *)e; M.exec_sig = sig; M.exec_launches = per_chunk;
        }
        exec = (cudaGraphExec_t)M.exec;
        per_chunk = M.exec_launches;
    }
#endif
    CGState h;
    int launched = 0;
    while (true) {
        CUDA_CHECK(cudaMemcpyAsync(s.cgst_host, s.cgst, sizeof(CGState), cudaMemcpyDeviceToHost, s.stream));
        xch_status_fetch(s);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        h = *s.cgst_host;
        if (xch_status_bad(s)) { h.fail = 1; h.converged = 0; break; }   // broken exchange: stop launching, xch_check reports it
        if (s.verbose > 1) printf("\t\tmg-pcg iteration %d max|r| %.3e (tol %.3e)\n", h.iter, h.resid, h.tol), fflush(stdout);
        if (h.done || launched >= maxit + chunk) break;
#ifndef FLIP_CPU_EMU
        if (exec) { CUDA_CHECK(cudaGraphLaunch(exec, s.stream)); s.kernel_launches += per_chunk; }
        else
#endif
            launch_chunk();
        KERNEL_CHECK();
        launched += chunk;
    }
    return h;
}

// _applySolutionToVelocityField: the whole field is cleared, unknowns get (float)soln
__global__ void __launch_bounds__(CG_THREADS) k_visc_store(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount, const float *__restrict__ vdiag, const double *__restrict__ x, float *__restrict__ vel) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k);
    size_t T = (size_t)g.total;
    for (int c = 0; c < 3; c++) {
        int w = g.ni + (c == 0), h = g.nj + (c == 1), d = g.nk + (c == 2);
        if (i >= w || j >= h || k >= d) continue;
        vel[c * T + id] = vdiag[c * T + id] != 0.0f ? (float)x[c * T + id] : 0.0f;
    }
  }
}

void stage_apply_viscosity(Sim &s, float dt) {
    s.visc_stats = SolveStats{0, 0, 0, 1, 0, 0, 0, 0};
    if (!s.viscosity_nonzero) return;  // src/fluidsimulation.cpp:171-184
    const Grid &g = s.g;
    cudaEvent_t e0, e1, es;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); CUDA_CHECK(cudaEventCreate(&es));
    CUDA_CHECK(cudaEventRecord(e0, s.stream));
    xch_update_cuts(s);   // k-slabs of this substep, balanced by liquid cells (no-op on one GPU)
    viscosity_volumes(s);
    // factor = dt * invdx * invdx in float (src/viscositysolver.cpp:379-380)
    float invdx = 1.0f / g.dx;
    float factor = dt * invdx * invdx;
    long long n1 = (long long)(g.ni + 1) * (g.nj + 1) * (g.nk + 1);
    const int *gl = s.grid_list, *gc = s.grid_count;   // built by viscosity_volumes
    FLIP_LAUNCH(k_visc_coefs, list_grid(s), CG_THREADS, s.stream, g, gl, gc, (const float *)s.viscosity, (const float *)s.vvol, s.vcoef, factor);
    FLIP_LAUNCH(k_visc_rows, list_grid(s), CG_THREADS, s.stream, g, (const float *)s.vvol, (const float *)s.vcoef,
                (const unsigned char *)s.fstate, (const float *)s.vel, s.vdiag, s.cg_r, s.vmass, s.visc_operator == 0 ? 1 : 0, gl, gc);
    s.kernel_launches += 2;
    DiagViscosity diag{s.vdiag, g.total};
    build_block_list<3>(s, diag);
    CUDA_CHECK(cudaMemsetAsync(s.cg_s, 0, sizeof(double) * 3 * (size_t)g.total, s.stream));
    CGParams P = cg_params(s, 0);
    int G = cg_grid(s);
    const float *vcoef = s.vcoef, *vdiag = s.vdiag, *vmass = s.vmass;
    cudaStream_t st = s.stream;
    int maxit = s.visc_maxit * s.visc_maxit_scale;
    CGState h;
    // phase A of an iteration: ghost planes of the search direction (U, V and W: rows couple components at k +- 1),
    // then the stencil
    auto apply_on = [&](CGParams &Q) {
        return [&s, &Q, G, st, vcoef, vdiag, vmass](int parity) {
            xch_push_halo(s, s.g, Q.s, sizeof(double), 3, 0, 1);
            FLIP_LAUNCH_SYNC(k_visc_apply, G, CG_THREADS, st, Q, vcoef, vdiag, vmass, parity);
            s.kernel_launches++;
        };
    };
    if (s.visc_precond == 2) {
        GMG *M = gmg_get(s);
        gmg_build(s, *M);
        CUDA_CHECK(cudaEventRecord(es, s.stream));
        P.z = s.cg_z;
        P.flexible = s.mg_flexible;
        h = run_cg_gmg(s, *M, P, diag, s.visc_tol, maxit, apply_on(P), s.visc_warm_start ? (const float *)s.vel : nullptr);
    } else if (s.cg_variant_viscosity == 1) {
        CUDA_CHECK(cudaMemsetAsync(s.cg_z, 0, sizeof(double) * 3 * (size_t)g.total, s.stream));
        CGParams Pu = P;
        Pu.s = s.cg_z; Pu.q = s.cg_w;
        h = run_cg2<3>(s, P, diag, 0.0, s.visc_tol, maxit, apply_on(Pu), 1);
    } else {
        h = run_cg<3>(s, P, diag, 0.0, s.visc_tol, maxit, apply_on(P), 1, s.visc_warm_start ? (const float *)s.vel : nullptr);
    }
    // every rank gets every slab of the solution
    xch_push_gather(s, g, s.cg_x, sizeof(double), 3, 0);
    xch_barrier(s);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    xch_check(s);   // a broken exchange must not reach the velocity field
    // acceptance rule of src/viscositysolver.cpp:676-689
    bool accept = !h.fail && (h.converged || (h.iter >= maxit && h.resid < s.visc_accept));
    if (accept) {
        FLIP_LAUNCH(k_visc_store, list_grid(s), CG_THREADS, s.stream, g, gl, gc, (const float *)s.vdiag, (const double *)s.cg_x, s.vel);
        s.kernel_launches++;
    }
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(s.count_host, s.blk_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaMemcpyAsync(s.count_host + 1, s.unk_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaEventRecord(e1, s.stream));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    s.visc_setup_ms = 0;
    if (s.visc_precond == 2) CUDA_CHECK(cudaEventElapsedTime(&s.visc_setup_ms, e0, es));
    CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1)); CUDA_CHECK(cudaEventDestroy(es));
    s.visc_stats.iters = h.iter; s.visc_stats.converged = h.converged; s.visc_stats.resid = h.resid;
    s.visc_stats.bmax = h.bmax; s.visc_stats.skipped = accept ? 0 : 2; s.visc_stats.blocks = s.count_host[0]; s.visc_stats.unknowns = s.count_host[1];
    s.visc_stats.ms = ms;
    if (s.verbose) {
        printf("\tviscosity: %d iterations, max|r| %.3e (tol %.3e), %s (%d active blocks, %.3f ms)\n", h.iter, h.resid,
               h.tol, accept ? (h.converged ? "converged" : "accepted") : "FAILED", *s.count_host, ms);
    }
}

// ---- debug access (tests / dev tools only) -------------------------------------------------------
// y = A x for the viscosity system of the last solve, on raw padded fp64 arrays [3*total] (tests /
// solver prototyping only).  Also: raw access to the row diagonals and face volumes.
extern "C" int flip_debug_visc_apply(void *hsim, const double *x_host, double *y_host) {
    Sim &s = *(Sim *)hsim;
    const Grid &g = s.g;
    size_t n = 3 * (size_t)g.total;
    try {
        CUDA_CHECK(cudaMemcpy(s.cg_z, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemsetAsync(s.cg_w, 0, n * sizeof(double), s.stream));
        CUDA_CHECK(cudaMemsetAsync(s.cgst, 0, 2 * sizeof(CGState), s.stream));
        CGParams P = cg_params(s, 0);
        P.s = s.cg_z; P.q = s.cg_w;
        P.X.nranks = 1;   // debug helper: whole field on this rank, no hand-shakes
        FLIP_LAUNCH_SYNC(k_visc_apply, cg_grid(s), CG_THREADS, s.stream, P, (const float *)s.vcoef, (const float *)s.vdiag, (const float *)s.vmass, 0);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        CUDA_CHECK(cudaMemcpy(y_host, s.cg_w, n * sizeof(double), cudaMemcpyDeviceToHost));
    } catch (...) { return -2; }
    return 0;
}
extern "C" int flip_debug_visc_rhs(void *hsim, double *b_host, float *diag_host, float *vol_host /*3T: U,V,W*/) {
    Sim &s = *(Sim *)hsim;
    size_t T = (size_t)s.g.total;
    // the rhs of the last solve is gone (r was consumed); recompute rows
    try {
        long long n1 = (long long)(s.g.ni + 1) * (s.g.nj + 1) * (s.g.nk + 1);
        grid_list_ensure(s);
        FLIP_LAUNCH(k_visc_rows, list_grid(s), CG_THREADS, s.stream, s.g, (const float *)s.vvol, (const float *)s.vcoef,
                    (const unsigned char *)s.fstate, (const float *)s.vel, s.vdiag, s.cg_r, s.vmass, s.visc_operator == 0 ? 1 : 0,
                    (const int *)s.grid_list, (const int *)s.grid_count);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        CUDA_CHECK(cudaMemcpy(b_host, s.cg_r, 3 * T * sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(diag_host, s.vdiag, 3 * T * sizeof(float), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(vol_host, s.vvol + T, 3 * T * sizeof(float), cudaMemcpyDeviceToHost));
    } catch (...) { return -2; }
    return 0;
}

// compact rows of the first explicit level of the last solve (tests / dev tools): meta = {valid, kept slots of component
// 0, 1, 2}; Sc_out [nrows * GMG_CSTRIDE] and cslot_out [3 * GMG_CSTRIDE] may be null
extern "C" int flip_debug_gmg_compact(void *hsim, int *meta /*[4]*/, float *Sc_out, int *cslot_out) {
    Sim &s = *(Sim *)hsim;
    GMG *M = (GMG *)s.gmg;
    if (!M || M->nlevels < 2 || !M->lv[1].cmeta) return -1;
    GLevel &L = M->lv[1];
    cudaStreamSynchronize(s.stream);
    if (cudaMemcpy(meta, L.cmeta, 4 * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    if (Sc_out && L.Sc && cudaMemcpy(Sc_out, L.Sc, (size_t)L.nrows * GMG_CSTRIDE * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    if (cslot_out && cudaMemcpy(cslot_out, L.cslot, 3 * GMG_CSTRIDE * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -4;
    return 0;
}

// explicit Galerkin level of the last solve (tests / dev tools): rows (m*T + padded id) and S [nrows * GMG_STRIDE]
extern "C" int flip_debug_gmg_level(void *hsim, int level, int *info /*[10]: ni,nj,nk,ax,ay,az,total,nrows,nlevels,stride*/,
                                    int *rows_out, float *S_out, float *diag_out) {
    Sim &s = *(Sim *)hsim;
    GMG *M = (GMG *)s.gmg;
    if (!M || level < 1 || level >= M->nlevels) return -1;
    GLevel &L = M->lv[level];
    const Grid &g = L.g;
    info[0] = g.ni; info[1] = g.nj; info[2] = g.nk; info[3] = g.ax; info[4] = g.ay; info[5] = g.az; info[6] = g.total;
    info[7] = L.nrows; info[8] = M->nlevels; info[9] = GMG_STRIDE;
    cudaStreamSynchronize(s.stream);
    if (rows_out && cudaMemcpy(rows_out, L.rows, (size_t)L.nrows * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    if (S_out && cudaMemcpy(S_out, L.S, (size_t)L.nrows * GMG_STRIDE * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    if (diag_out && cudaMemcpy(diag_out, L.diag, 3 * (size_t)g.total * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -4;
    return 0;
}

// Average duration (CUDA events on the library's stream) of `reps` back-to-back launches of a hot kernel on the
// hierarchy of the last viscosity solve, and its algorithmic bytes per launch:
//   "gmg_sweep_l1"  one damped-Jacobi sweep on the first explicit level: rows x (235 coefficients + row index +
//                   weight + b + x in + x out) x 4 B
//   "visc_apply"    the level-0 matrix-free stencil apply of the CG (120 B per unknown share of SURVEY 8d: 2 vector
//                   transfers x 8 B + 16 B coefficients per unknown)
int viscosity_time_kernel(Sim &s, const char *name, int reps, float *ms_per_launch, unsigned long long *alg_bytes) {
    std::string n(name ? name : "");
    if (reps < 1) reps = 1;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    if (n == "gmg_sweep_l1") {
        GMG *M = (GMG *)s.gmg;
        if (!M || M->nlevels < 2 || M->lv[1].nrows == 0) return -1;
        GLevel &L = M->lv[1];
        GLevelDev D = gmg_dev(s, L, false);
#ifndef FLIP_CPU_EMU
        auto sweep1 = s.mg_tma ? &k_gmg_sweep_tma<1> : &k_gmg_sweep<1>;
#else
        auto sweep1 = &k_gmg_sweep<1>;
#endif
        int GR = gmg_row_grid(s, L);
        const float *nof = nullptr;
        auto one = [&](int k) {
            FLIP_LAUNCH_SYNC(sweep1, GR, 256, s.stream, D, (const float *)L.b, (const float *)L.x[k & 1], L.x[(k & 1) ^ 1], nof, M->omega, (const CGState *)nullptr);
        };
        one(0);
        CUDA_CHECK(cudaEventRecord(e0, s.stream));
        for (int k = 0; k < reps; k++) one(k);
        CUDA_CHECK(cudaEventRecord(e1, s.stream));
        int meta[4] = {0, 0, 0, 0};
        if (D.cmeta) CUDA_CHECK(cudaMemcpy(meta, L.cmeta, sizeof(meta), cudaMemcpyDeviceToHost));
        // stored slots of a row (compact rows: 160, else 240 of which 5 are alignment padding) + row index, weight, b, x in, x out
        *alg_bytes = (unsigned long long)L.nrows * ((meta[0] ? GMG_CSTRIDE : GMG_SLOTS) + 5) * 4ull;
    } else if (n == "visc_apply") {
        if (s.visc_stats.unknowns == 0) return -1;
        CGParams P = cg_params(s, 0);
        P.X.nranks = 1;   // timing helper: this rank's cells, no hand-shakes
        CUDA_CHECK(cudaMemsetAsync(s.cgst, 0, 2 * sizeof(CGState), s.stream));
        int G = cg_grid(s);
        FLIP_LAUNCH_SYNC(k_visc_apply, G, CG_THREADS, s.stream, P, (const float *)s.vcoef, (const float *)s.vdiag, (const float *)s.vmass, 0);
        CUDA_CHECK(cudaEventRecord(e0, s.stream));
        for (int k = 0; k < reps; k++)
            FLIP_LAUNCH_SYNC(k_visc_apply, G, CG_THREADS, s.stream, P, (const float *)s.vcoef, (const float *)s.vdiag, (const float *)s.vmass, 0);
        CUDA_CHECK(cudaEventRecord(e1, s.stream));
        *alg_bytes = (unsigned long long)s.visc_stats.unknowns * (2 * 8 + 16);
    } else {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return -2;
    }
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    KERNEL_CHECK();
    s.kernel_launches += reps + 1;
    *ms_per_launch = ms / reps;
    return 0;
}
