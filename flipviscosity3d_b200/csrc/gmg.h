// Galerkin geometric multigrid: the preconditioner of the viscosity CG.
//
// The reference preconditions with MIC(0) (src/pcgsolver/pcgsolver.h:62-214), a sequential
// recurrence that barely beats the diagonal on this coupled system (64^3: 364 vs 468 iterations)
// and hits its 700-iteration cap from 128^3 up.  A rediscretised V-cycle (round 1, removed)
// gains 5-13x; its weak point, measured on a scipy prototype of the same operator
// (dev/visc_proto.py), is the coarse operator next to the free surface.  Here the coarse operators
// are the exact Galerkin products  A_c = P^T A P / 8  with staggered trilinear
// transfers (linear along the face normal, cell-centred linear across it, renormalised at the free
// surface): 20-25x fewer iterations than the diagonal at 64^3-128^3, and nothing to tune.
//
//   level 0       the solver's own matrix-free coupled stencil (k_visc_apply)
//   levels >= 1   explicit windowed stencils.  A product of trilinear transfers with a 15-point
//                 operator closes on a fixed window per (row component m, column component m'):
//                 m == m': 3 x 5 x 5 (3 along the face normal), m != m': 4 x 4 x 5  ->  235 slots
//                 per row on every level.  Stored row-major over a compact row list
//                 (S[row * 240 + slot]); one WARP applies one row (coalesced 960-byte read, 8 slots
//                 per lane, shuffle reduction), so even the 100-row levels finish in one memory
//                 round trip.  Vectors stay in the dense padded layout: neighbours are plain offsets.
//   set-up        one warp per (coarse row, column component) accumulates its <= 80 slots in fp64 in shared
//                 memory in a fixed order (no atomics, bit-reproducible); blocks below the diagonal are mirrored.
//   cycle         V-cycle with 3 / 1 / 2 damped-Jacobi sweeps before = after the coarse correction on level 0 /
//                 level 1 / deeper levels, exact dense solve on the last level (<= 128 rows), fp32 inside (the outer
//                 CG stays fp64), symmetric, so plain CG remains valid.
#pragma once
#include "cg.h"
#include "resident.h"
#include "scan.h"

#define GMG_MAX_LEVELS 8
#define GMG_SLOTS 235
#define GMG_STRIDE 240     // floats per stored row (235 slots + zero padding; 960 B)
#define GMG_CSTRIDE 160    // floats per COMPACT row of the first explicit level (see k_gmg_compact_rows; 640 B)

// one axis of the prolongation: fine index n -> up to two coarse indices and weights
FLIP_D void gmg_parents(bool own, int n, int &p0, int &p1, float &w0, float &w1) {
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
FLIP_D float gmg_pnorm(int m, int i, int j, int k, const Grid &gc, const float *__restrict__ diag_c) {
    int pi[2], pj[2], pk[2];
    float wi[2], wj[2], wk[2];
    gmg_parents(m == 0, i, pi[0], pi[1], wi[0], wi[1]);
    gmg_parents(m == 1, j, pj[0], pj[1], wj[0], wj[1]);
    gmg_parents(m == 2, k, pk[0], pk[1], wk[0], wk[1]);
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

struct GWin { int lo[3], n[3], base, size; };

// window of column component mp inside a row of component m
FLIP_HD GWin gmg_window(int m, int mp) {
    GWin w;
    w.base = 0;
    for (int q = 0; q < mp; q++) w.base += (q == m) ? 75 : 80;
    for (int a = 0; a < 3; a++) {
        if (m == mp) { w.lo[a] = (a == m) ? -1 : -2; w.n[a] = (a == m) ? 3 : 5; }
        else if (a == m) { w.lo[a] = -2; w.n[a] = 4; }
        else if (a == mp) { w.lo[a] = -1; w.n[a] = 4; }
        else { w.lo[a] = -2; w.n[a] = 5; }
    }
    w.size = w.n[0] * w.n[1] * w.n[2];
    return w;
}

struct GLevel {
    Grid g;
    float *diag = 0;      // [3T] > 0 on unknowns (level 0: the solver's vdiag; else Galerkin diagonal)
    float *pn = 0;        // [3T] sum of the prolongation weights of a face's coarse parents that exist
    float *x[2] = {0, 0}, *b = 0, *r = 0;   // [3T]
    int *blk_flag = 0, *blk_list = 0, *blk_count = 0;
    int *rows = 0;        // [nrows] m*T + id, k-PLANE major (a rank's slab of planes is one contiguous row range), then
                          //         8x8 tile of the plane, component, cell of the tile
    int *rowmap = 0;      // [3T] inverse of rows (-1 = not an unknown)
    int *tile_off = 0;    // [ntiles + 1] first row of every (plane, tile); ntiles = 8 nbz * nby * nbx
    int *scan_tmp = 0;
    int ntiles = 0;
    int *rng = 0;         // device [4]: {first, end} row of this rank's slab, then {0, nrows}
    int *nrows_dev = 0;
    int nrows = 0;
    size_t cap = 0;       // rows S is allocated for
    float *S = 0;         // [nrows * GMG_STRIDE]
    float *wj = 0;        // [nrows] smoothing weight per row
    int *offs = 0;        // [3 * GMG_STRIDE] slot -> element offset table (set once: it only depends on the grid)
    // compact rows (first explicit level only, k_gmg_compact_rows): the sweeps read these instead of S
    float *Sc = 0;        // [nrows * GMG_CSTRIDE]
    int *coffs = 0;       // [3 * GMG_CSTRIDE] compact slot -> element offset
    int *cslot = 0;       // [3 * GMG_CSTRIDE] compact slot -> slot of the full row (-1: padding)
    int *cmeta = 0;       // [4] {compact rows valid for this solve, kept slots of component 0, 1, 2}
    unsigned *cmask = 0;  // [3 * 8] union over the rows of a component of their non-zero slots
    // x-groups (first explicit level only): the rows of one 8-wide x-run of a tile (same component, j, k) are consecutive
    // in `rows`; a group = {first row, 8-bit mask of the cells of the run that are unknowns}.  k_gmg_restrict_x
    int2 *groups = 0;     // [<= 3T/8]
    int *gtile_off = 0;   // [ntiles + 1] first group of every (plane, tile)
    int *grng = 0;        // device [4]: {first, end} group of this rank's slab, then {0, ngroups}
    bool owns = false;
};

struct GMG {
    int nlevels = 0, nalloc = 0;
    GLevel lv[GMG_MAX_LEVELS];
    int pre = 2, coarse_sweeps = 24;
    int pre_l[GMG_MAX_LEVELS] = {2, 1, 2, 2, 2, 2, 2, 2};   // sweeps before = after the coarse correction, per level
    float omega = 0.5f;
    int *count_host = 0;  // pinned
    double *dense = 0;    // [GMG_DENSE_MAX^2] scratch of the dense inversion
    float *Ainv = 0;      // [GMG_DENSE_MAX^2] inverse of the coarsest operator
    bool dense_last = false;
    void *exec = 0;       // cudaGraphExec_t of one chunk of multigrid-PCG iterations
    unsigned long long exec_sig = 0;
    long long exec_launches = 0;
};

struct GLevelDev {
    Grid g;
    const float *diag;
    const int *rows;
    const int *nrows;
    const int *rng;      // {first, end} row this launch works on (the rank's slab on a sharded level, else all rows)
    Xch X;               // exchange context; nranks == 1 on replicated levels and on one GPU
    const float *S;
    const float *wj;     // [nrows] smoothing weights
    const int *offs;     // [3 * GMG_STRIDE] element offset of every slot, per row component (0 for padding)
    const float *Sc;     // compact rows / table / flag of the level, or null (then the sweeps read S)
    const int *coffs;
    const int *cmeta;
    const int2 *groups;  // x-groups of the level (null: none built)
    const int *grng;     // {first, end} group this launch works on
};

FLIP_D void gmg_unflatten(const Grid &g, int id, int &i, int &j, int &k) {
    int x = id % g.ax, r = id / g.ax;
    i = x - FLIP_PX; j = (r % g.ay) - FLIP_PY; k = (r / g.ay) - FLIP_PZ;
}

// coarse unknown flags: a coarse face is an unknown if one of its nearest fine faces is one
__global__ void __launch_bounds__(256) k_gmg_flags(Grid gc, Grid gf, const float *__restrict__ diag_f, float *__restrict__ diag_c) {
    int I, J, K;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, gc.ni + 1, gc.nj + 1, gc.nk + 1, I, J, K)) return;
    size_t Tf = (size_t)gf.total, Tc = (size_t)gc.total;
    int id = gidx(gc, I, J, K);
    bool interior = I >= 1 && I < gc.ni && J >= 1 && J < gc.nj && K >= 1 && K < gc.nk;
    for (int m = 0; m < 3; m++) {
        bool any = false;
        if (interior)
            for (int a = (m == 0 ? -1 : 0); a <= 1 && !any; a++)
                for (int b = (m == 1 ? -1 : 0); b <= 1 && !any; b++)
                    for (int c = (m == 2 ? -1 : 0); c <= 1; c++) {
                        int fi = 2 * I + a, fj = 2 * J + b, fk = 2 * K + c;
                        if (fi < 0 || fj < 0 || fk < 0 || fi > gf.ni || fj > gf.nj || fk > gf.nk) continue;
                        if (diag_f[m * Tf + gidx(gf, fi, fj, fk)] != 0.0f) { any = true; break; }
                    }
        diag_c[m * Tc + id] = any ? 1.0f : 0.0f;
    }
}

// ---- row list of an explicit level, k-plane major ---------------------------------------------------------------
// tile = one k-plane of an 8x8x8 block (64 cells, two warps); tile id = plane * (nby * nbx) + bj * nbx + bi.
// per_tile[tile] = unknowns in the tile (all three components); zero for tiles of inactive blocks (memset by the host).
__global__ void __launch_bounds__(CG_THREADS) k_gmg_row_counts(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                                const float *__restrict__ diag, int *__restrict__ per_tile,
                                                                int *__restrict__ per_tile_g /* x-groups per tile, or null */) {
    __shared__ int wsum[CG_THREADS / 32];
    __shared__ int gsum[CG_THREADS / 32];
    int nb = *count;
    size_t T = (size_t)g.total;
    const int tpp = g.nbx * g.nby;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int blk = list[b];
        BlockCell c = block_cell(g, blk, threadIdx.x);
        int n = 0, ng = 0;
        const int id = c.inside ? gidx(g, c.i, c.j, c.k) : 0;
        for (int m = 0; m < 3; m++) {
            const bool has = c.inside && diag[m * T + id] != 0.0f;
            n += has ? 1 : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, has);   // a warp = 4 x-runs of 8 cells
            for (int yy = 0; yy < 4; yy++) ng += ((bal >> (8 * yy)) & 0xffu) != 0u ? 1 : 0;
        }
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = n; gsum[threadIdx.x >> 5] = ng; }
        __syncthreads();
        if (threadIdx.x < 8) {   // thread p: plane p of the block = warps 2p, 2p+1
            int bk = blk / tpp, rest = blk - bk * tpp;
            per_tile[(bk * FLIP_B + threadIdx.x) * tpp + rest] = wsum[2 * threadIdx.x] + wsum[2 * threadIdx.x + 1];
            if (per_tile_g) per_tile_g[(bk * FLIP_B + threadIdx.x) * tpp + rest] = gsum[2 * threadIdx.x] + gsum[2 * threadIdx.x + 1];
        }
    }
}

__global__ void __launch_bounds__(CG_THREADS) k_gmg_row_fill(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                              const float *__restrict__ diag, const int *__restrict__ tile_off,
                                                              int *__restrict__ rows, int *__restrict__ rowmap,
                                                              const int *__restrict__ gtile_off, int2 *__restrict__ groups /* or null */) {
    __shared__ int wsum[CG_THREADS / 32];
    __shared__ int gsum[CG_THREADS / 32];
    int nb = *count;
    size_t T = (size_t)g.total;
    const int tpp = g.nbx * g.nby;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int blk = list[b];
        BlockCell c = block_cell(g, blk, threadIdx.x);
        int id = c.inside ? gidx(g, c.i, c.j, c.k) : 0;
        const int bk = blk / tpp, rest = blk - bk * tpp;
        int base = tile_off[(bk * FLIP_B + (wid >> 1)) * tpp + rest];   // first row of this thread's plane tile
        int gbase = groups ? gtile_off[(bk * FLIP_B + (wid >> 1)) * tpp + rest] : 0;
        for (int m = 0; m < 3; m++) {
            bool has = c.inside && diag[m * T + id] != 0.0f;
            unsigned bal = __ballot_sync(0xffffffffu, has);
            int ng = 0;
            for (int yy = 0; yy < 4; yy++) ng += ((bal >> (8 * yy)) & 0xffu) != 0u ? 1 : 0;
            __syncthreads();
            if (lane == 0) { wsum[wid] = __popc(bal); gsum[wid] = ng; }
            __syncthreads();
            const int before = (wid & 1) ? wsum[wid - 1] : 0, total = wsum[wid & ~1] + wsum[wid | 1];
            const int r = base + before + __popc(bal & ((1u << lane) - 1u));
            if (has) {
                rows[r] = (int)(m * T + id);
                rowmap[m * T + id] = r;
            }
            if (groups) {
                // the first unknown of every non-empty 8-cell x-run writes the run's group (tile order: component, warp, run)
                const int yy = lane >> 3;
                const unsigned run = (bal >> (8 * yy)) & 0xffu;
                if (has && (run & ((1u << (lane & 7)) - 1u)) == 0u) {
                    int gi = gbase + ((wid & 1) ? gsum[wid - 1] : 0);
                    for (int y2 = 0; y2 < yy; y2++) gi += ((bal >> (8 * y2)) & 0xffu) != 0u ? 1 : 0;
                    int2 G; G.x = r; G.y = (int)run;
                    groups[gi] = G;
                }
                gbase += gsum[wid & ~1] + gsum[wid | 1];
            }
            base += total;
        }
    }
}

// row ranges of a level: rng[0..1] = this rank's slab of planes (cuts of `level`, or everything when cuts == null),
// rng[2..3] = all rows
__global__ void k_gmg_ranges(const int *__restrict__ tile_off, int ntiles, int tpp, const Cuts *__restrict__ cuts, int level, int rank,
                             int *__restrict__ rng, int *__restrict__ nrows, const int *__restrict__ gtile_off, int *__restrict__ grng) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = tile_off[ntiles];
    int a = 0, b = n;
    int t0 = 0, t1 = ntiles;
    if (cuts) {
        t0 = cuts->c[level][rank] * tpp; t1 = cuts->c[level][rank + 1] * tpp;
        t0 = t0 < ntiles ? t0 : ntiles; t1 = t1 < ntiles ? t1 : ntiles;
        a = tile_off[t0];
        b = tile_off[t1];
    }
    rng[0] = a; rng[1] = b; rng[2] = 0; rng[3] = n;
    *nrows = n;
    if (grng) { grng[0] = gtile_off[t0]; grng[1] = gtile_off[t1]; grng[2] = 0; grng[3] = gtile_off[ntiles]; }
}

// pn: prolongation normaliser of every unknown of the fine level
__global__ void __launch_bounds__(CG_THREADS) k_gmg_pnorm(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                           const float *__restrict__ diag, Grid gc, const float *__restrict__ diag_c,
                                                           float *__restrict__ pn) {
    int nb = *count;
    size_t T = (size_t)g.total;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        BlockCell c = block_cell(g, list[b], threadIdx.x);
        if (!c.inside) continue;
        int id = gidx(g, c.i, c.j, c.k);
        for (int m = 0; m < 3; m++)
            pn[m * T + id] = diag[m * T + id] != 0.0f ? gmg_pnorm(m, c.i, c.j, c.k, gc, diag_c) : 0.0f;
    }
}

// --- fine-operator entry enumeration ---------------------------------------------------------
// level 0: the <= 7 entries of row (m, id) whose column component is mp (csrc/viscosity.cu k_visc_apply)
FLIP_D int gmg_entries_l0(int m, int mp, int id, int sy, int sz, size_t T, const float *__restrict__ coef,
                          const float *__restrict__ vol, int off[7][3], double val[7]) {
    const float *cc = coef, *cu = coef + T, *cv = coef + 2 * T, *cw = coef + 3 * T;
    float fR, fL, fT, fB, fF, fK;
    if (m == 0) { fR = cc[id]; fL = cc[id - 1]; fT = cw[id + sy]; fB = cw[id]; fF = cv[id + sz]; fK = cv[id]; }
    else if (m == 1) { fR = cw[id + 1]; fL = cw[id]; fT = cc[id]; fB = cc[id - sy]; fF = cu[id + sz]; fK = cu[id]; }
    else { fR = cv[id + 1]; fL = cv[id]; fT = cu[id + sy]; fB = cu[id]; fF = cc[id]; fK = cc[id - sz]; }
    int n = 0;
#define GMG_E(di, dj, dk, v) { off[n][0] = di; off[n][1] = dj; off[n][2] = dk; val[n] = v; n++; }
    if (m == mp) {
        GMG_E(0, 0, 0, (double)vol[(1 + m) * T + id] + (double)fR + (double)fL + (double)fT + (double)fB + (double)fF + (double)fK) GMG_E(1, 0, 0, -fR) GMG_E(-1, 0, 0, -fL) GMG_E(0, 1, 0, -fT) GMG_E(0, -1, 0, -fB)
        GMG_E(0, 0, 1, -fF) GMG_E(0, 0, -1, -fK)
    } else if (m == 0 && mp == 1) { GMG_E(0, 1, 0, -fT) GMG_E(-1, 1, 0, fT) GMG_E(0, 0, 0, fB) GMG_E(-1, 0, 0, -fB) }
    else if (m == 0 && mp == 2) { GMG_E(0, 0, 1, -fF) GMG_E(-1, 0, 1, fF) GMG_E(0, 0, 0, fK) GMG_E(-1, 0, 0, -fK) }
    else if (m == 1 && mp == 0) { GMG_E(1, 0, 0, -fR) GMG_E(1, -1, 0, fR) GMG_E(0, 0, 0, fL) GMG_E(0, -1, 0, -fL) }
    else if (m == 1 && mp == 2) { GMG_E(0, 0, 1, -fF) GMG_E(0, -1, 1, fF) GMG_E(0, 0, 0, fK) GMG_E(0, -1, 0, -fK) }
    else if (m == 2 && mp == 0) { GMG_E(1, 0, 0, -fR) GMG_E(1, 0, -1, fR) GMG_E(0, 0, 0, fL) GMG_E(0, 0, -1, -fL) }
    else { GMG_E(0, 1, 0, -fT) GMG_E(0, 1, -1, fT) GMG_E(0, 0, 0, fB) GMG_E(0, 0, -1, -fB) }
#undef GMG_E
    return n;
}

// A_c = P^T A P / 8.  One WARP per (coarse row I, column component mp); grid = ceil(3 * nrows / 8) CTAs of 8 warps.
// The warp walks the fine COLUMNS j that can reach a child of I, 32 at a time (one per lane):
//   t_j = sum_i P[i,I] A[i,j]   is gathered by the lane from row j of the (symmetric) fine operator, so the global
//                               loads of 32 columns are in flight together;
//   the <= 8 contributions t_j P[j,J] of every lane are then added to the warp's 80 fp64 accumulators in shared
//   memory in LANE ORDER (staged in shared memory; contribution c of lane l is applied by lane c in step l: the 8
//   parents of a column are distinct slots), so the sums are formed in a fixed order: no atomics, bit-reproducible.
// fp64 accumulation: the mass term of a coarse row (its row sum, ~1) is what remains of thousands of products of
// size ~1e4 that cancel.  FINE0: the fine level is level 0 (matrix-free coefficients), else an explicit level.
#define GMG_BUILD_WARPS 8
template <bool FINE0>
__global__ void __launch_bounds__(32 * GMG_BUILD_WARPS) k_gmg_build(Grid gc, Grid gf, const int *__restrict__ rows_c, const int *__restrict__ rng,
                                                                     const float *__restrict__ diag_c, float *__restrict__ S_c,
                                                                     const float *__restrict__ diag_f, const float *__restrict__ pn_f,
                                                                     const float *__restrict__ coef_f, const int *__restrict__ rowmap_f,
                                                                     const float *__restrict__ S_f, int nrows_f) {
    __shared__ double acc_s[GMG_BUILD_WARPS][80];
    __shared__ float wP_s[GMG_BUILD_WARPS][64];
    __shared__ int sslot_s[GMG_BUILD_WARPS][256];
    __shared__ double sval_s[GMG_BUILD_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int task = blockIdx.x * GMG_BUILD_WARPS + wid;        // (row of this rank's range, mp), mp fastest
    const int r = rng[0] + task / 3, mp = task % 3;
    if (r >= rng[1]) return;                                    // whole warp leaves together; no block-wide barrier below
    // A_c is symmetric: the blocks below the diagonal (row component > column component) are mirrored from the ones
    // above it by k_gmg_mirror instead of being computed a second time (a third of the tasks)
    if (rows_c[r] / gc.total > mp) return;
    double *acc = acc_s[wid];
    float *wP = wP_s[wid];
    int *sslot = sslot_s[wid];
    double *sval = sval_s[wid];
    const size_t Tc = (size_t)gc.total, Tf = (size_t)gf.total;
    const int enc = rows_c[r];
    const int m = enc / (int)Tc, idc = enc - m * (int)Tc;
    int I, J, K;
    gmg_unflatten(gc, idc, I, J, K);
    const GWin W = gmg_window(m, mp);
    for (int q = lane; q < 80; q += 32) acc[q] = 0.0;
    const int sy = SY(gf), sz = SZ(gf);
    // children of I (component m) and their prolongation weights P[i,I]: a 4 x 4 x 4 box from (2I-1, 2J-1, 2K-1)
    const int lo[3] = {2 * I - 1, 2 * J - 1, 2 * K - 1};
    const int cn[3] = {m == 0 ? 3 : 4, m == 1 ? 3 : 4, m == 2 ? 3 : 4};
    const int nf[3] = {gf.ni, gf.nj, gf.nk};
#pragma unroll
    for (int t = 0; t < 2; t++) {
        int q = lane + 32 * t;
        int a = q & 3, b2 = (q >> 2) & 3, c2 = q >> 4;
        int fi = lo[0] + a, fj = lo[1] + b2, fk = lo[2] + c2;
        float wi = m == 0 ? (a == 1 ? 1.0f : (a < 3 ? 0.5f : 0.0f)) : ((a == 1 || a == 2) ? 0.75f : 0.25f);
        float wj = m == 1 ? (b2 == 1 ? 1.0f : (b2 < 3 ? 0.5f : 0.0f)) : ((b2 == 1 || b2 == 2) ? 0.75f : 0.25f);
        float wk = m == 2 ? (c2 == 1 ? 1.0f : (c2 < 3 ? 0.5f : 0.0f)) : ((c2 == 1 || c2 == 2) ? 0.75f : 0.25f);
        float w = wi * wj * wk;
        if (w != 0.0f && fi >= 0 && fj >= 0 && fk >= 0 && fi <= nf[0] && fj <= nf[1] && fk <= nf[2]) {
            size_t o = m * Tf + gidx(gf, fi, fj, fk);
            float p = pn_f[o];
            w = (diag_f[o] != 0.0f && p > 0.0f) ? w / p : 0.0f;
        } else w = 0.0f;
        wP[q] = w;
    }
    __syncwarp();
    // offsets o = i - j of row j (component mp) towards component m, and the box of candidate columns
    const GWin R = gmg_window(mp, m);
    int olo[3], ohi[3], jlo[3], jn[3];
    for (int a = 0; a < 3; a++) {
        olo[a] = FINE0 ? -1 : R.lo[a];
        ohi[a] = FINE0 ? 1 : R.lo[a] + R.n[a] - 1;
        jlo[a] = lo[a] - ohi[a] < 0 ? 0 : lo[a] - ohi[a];
        int jhi = lo[a] + cn[a] - 1 - olo[a] > nf[a] ? nf[a] : lo[a] + cn[a] - 1 - olo[a];
        jn[a] = jhi - jlo[a] + 1;
        if (jn[a] < 0) jn[a] = 0;
    }
    const int ncol = jn[0] * jn[1] * jn[2];
    for (int base = 0; base < ncol; base += 32) {
        // ---- gather: this lane's column
        const int c = base + lane;
        double t = 0.0;
        int ji = 0, jj = 0, jk = 0;
        float pnj = 0.0f;
        if (c < ncol) {
            ji = jlo[0] + c % jn[0]; jj = jlo[1] + (c / jn[0]) % jn[1]; jk = jlo[2] + c / (jn[0] * jn[1]);
            const int idj = gidx(gf, ji, jj, jk);
            const size_t oj = mp * Tf + idj;
            if (diag_f[oj] != 0.0f) pnj = pn_f[oj];       // a column that is not an unknown drops out (pnj stays 0)
            if (pnj != 0.0f) {
                if (FINE0) {
                    int off[7][3];
                    double val[7];
                    int ne = gmg_entries_l0(mp, m, idj, sy, sz, Tf, coef_f, S_f /* level 0: the volume grids */, off, val);
                    for (int e = 0; e < ne; e++) {
                        int a = ji + off[e][0] - lo[0], b2 = jj + off[e][1] - lo[1], c2 = jk + off[e][2] - lo[2];
                        if (a < 0 || b2 < 0 || c2 < 0 || a >= cn[0] || b2 >= cn[1] || c2 >= cn[2]) continue;
                        t += (double)wP[(c2 * 4 + b2) * 4 + a] * val[e];
                    }
                } else {
                    const float *__restrict__ Sj = S_f + (size_t)rowmap_f[oj] * GMG_STRIDE + R.base;
                    int a0 = ji + olo[0] - lo[0], a1 = ji + ohi[0] - lo[0];
                    int b0 = jj + olo[1] - lo[1], b1 = jj + ohi[1] - lo[1];
                    int c0 = jk + olo[2] - lo[2], c1 = jk + ohi[2] - lo[2];
                    a0 = a0 < 0 ? 0 : a0; b0 = b0 < 0 ? 0 : b0; c0 = c0 < 0 ? 0 : c0;
                    a1 = a1 > cn[0] - 1 ? cn[0] - 1 : a1; b1 = b1 > cn[1] - 1 ? cn[1] - 1 : b1; c1 = c1 > cn[2] - 1 ? cn[2] - 1 : c1;
                    for (int c2 = c0; c2 <= c1; c2++)
                        for (int b2 = b0; b2 <= b1; b2++) {
                            const float *__restrict__ Sr = Sj + ((lo[2] + c2 - jk - R.lo[2]) * R.n[1] + (lo[1] + b2 - jj - R.lo[1])) * R.n[0] +
                                                           (lo[0] - ji - R.lo[0]);
                            for (int a = a0; a <= a1; a++) t += (double)wP[(c2 * 4 + b2) * 4 + a] * (double)Sr[a];
                        }
                }
                t /= (double)pnj;
            }
        }
        // ---- this lane's <= 8 contributions (slot, value); slot -1 = none
        int pi[2], pj[2], pk[2];
        float wi[2], wj[2], wk[2];
        gmg_parents(mp == 0, ji, pi[0], pi[1], wi[0], wi[1]);
        gmg_parents(mp == 1, jj, pj[0], pj[1], wj[0], wj[1]);
        gmg_parents(mp == 2, jk, pk[0], pk[1], wk[0], wk[1]);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int a = q & 1, b2 = (q >> 1) & 1, c2 = q >> 2;
            float w = wi[a] * wj[b2] * wk[c2];
            int PI = pi[a], PJ = pj[b2], PK = pk[c2];
            int slot = -1;
            if (t != 0.0 && w != 0.0f && PI >= 0 && PJ >= 0 && PK >= 0 && PI <= gc.ni && PJ <= gc.nj && PK <= gc.nk &&
                diag_c[(size_t)mp * Tc + gidx(gc, PI, PJ, PK)] != 0.0f) {
                int oi = PI - I - W.lo[0], oj2 = PJ - J - W.lo[1], ok = PK - K - W.lo[2];
                if (oi >= 0 && oj2 >= 0 && ok >= 0 && oi < W.n[0] && oj2 < W.n[1] && ok < W.n[2]) slot = (ok * W.n[1] + oj2) * W.n[0] + oi;
            }
            sslot[lane * 8 + q] = slot;
            sval[lane * 8 + q] = t * (double)w;
        }
        // ---- ordered scatter: step l applies lane l's contributions, contribution q by lane q
        const unsigned active = __ballot_sync(0xffffffffu, t != 0.0);
        __syncwarp();
        for (int l = 0; l < 32; l++) {
            if (!((active >> l) & 1u)) continue;      // uniform across the warp
            if (lane < 8) {
                int slot = sslot[l * 8 + lane];
                if (slot >= 0) acc[slot] += sval[l * 8 + lane];
            }
            __syncwarp();
        }
    }
    __syncwarp();
    for (int q = lane; q < W.size; q += 32) S_c[(size_t)r * GMG_STRIDE + W.base + q] = (float)(0.125 * acc[q]);
}

// The same product in GATHER form (mg_build = 1, default).  Phase 1 is the scatter kernel's: t_j for every fine column of
// the box, but kept in shared memory (fp64) instead of being scattered.  Phase 2: every lane owns output slots J and sums
// t_j P[j,J] over the <= 48 children j of J, walking them in ascending column order - the order in which the scatter
// kernel's lane-ordered steps reach a slot - so the two kernels are bit-identical (tests/test_gmg.py), without the
// 32 serialised steps per 32 columns.
#define GMG_TBOX_MAX 448   // columns of the box: children (<= 4 per axis) widened by the fine operator's reach (<= 2 each way)
FLIP_D float gmg_child_weight(bool own, int q) {   // child q = 0..3 of a coarse index along one axis (fine index 2I-1+q)
    return own ? (q == 1 ? 1.0f : (q < 3 ? 0.5f : 0.0f)) : ((q == 1 || q == 2) ? 0.75f : 0.25f);
}
template <bool FINE0>
__global__ void __launch_bounds__(32 * GMG_BUILD_WARPS) k_gmg_build_g(Grid gc, Grid gf, const int *__restrict__ rows_c, const int *__restrict__ rng,
                                                                       const float *__restrict__ diag_c, float *__restrict__ S_c,
                                                                       const float *__restrict__ diag_f, const float *__restrict__ pn_f,
                                                                       const float *__restrict__ coef_f, const int *__restrict__ rowmap_f,
                                                                       const float *__restrict__ S_f, int nrows_f) {
    __shared__ double t_s[GMG_BUILD_WARPS][GMG_TBOX_MAX];
    __shared__ float wP_s[GMG_BUILD_WARPS][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int task = blockIdx.x * GMG_BUILD_WARPS + wid;        // (row of this rank's range, mp), mp fastest
    const int r = rng[0] + task / 3, mp = task % 3;
    if (r >= rng[1]) return;                                    // whole warp leaves together; no block-wide barrier below
    if (rows_c[r] / gc.total > mp) return;                      // lower blocks are mirrored (k_gmg_mirror)
    double *tb = t_s[wid];
    float *wP = wP_s[wid];
    const size_t Tc = (size_t)gc.total, Tf = (size_t)gf.total;
    const int enc = rows_c[r];
    const int m = enc / (int)Tc, idc = enc - m * (int)Tc;
    int I, J, K;
    gmg_unflatten(gc, idc, I, J, K);
    const GWin W = gmg_window(m, mp);
    const int sy = SY(gf), sz = SZ(gf);
    const int lo[3] = {2 * I - 1, 2 * J - 1, 2 * K - 1};
    const int cn[3] = {m == 0 ? 3 : 4, m == 1 ? 3 : 4, m == 2 ? 3 : 4};
    const int nf[3] = {gf.ni, gf.nj, gf.nk};
#pragma unroll
    for (int t = 0; t < 2; t++) {
        int q = lane + 32 * t;
        int a = q & 3, b2 = (q >> 2) & 3, c2 = q >> 4;
        int fi = lo[0] + a, fj = lo[1] + b2, fk = lo[2] + c2;
        float w = gmg_child_weight(m == 0, a) * gmg_child_weight(m == 1, b2) * gmg_child_weight(m == 2, c2);
        if (w != 0.0f && fi >= 0 && fj >= 0 && fk >= 0 && fi <= nf[0] && fj <= nf[1] && fk <= nf[2]) {
            size_t o = m * Tf + gidx(gf, fi, fj, fk);
            float p = pn_f[o];
            w = (diag_f[o] != 0.0f && p > 0.0f) ? w / p : 0.0f;
        } else w = 0.0f;
        wP[q] = w;
    }
    __syncwarp();
    const GWin R = gmg_window(mp, m);
    int olo[3], ohi[3], jlo[3], jn[3];
    for (int a = 0; a < 3; a++) {
        olo[a] = FINE0 ? -1 : R.lo[a];
        ohi[a] = FINE0 ? 1 : R.lo[a] + R.n[a] - 1;
        jlo[a] = lo[a] - ohi[a] < 0 ? 0 : lo[a] - ohi[a];
        int jhi = lo[a] + cn[a] - 1 - olo[a] > nf[a] ? nf[a] : lo[a] + cn[a] - 1 - olo[a];
        jn[a] = jhi - jlo[a] + 1;
        if (jn[a] < 0) jn[a] = 0;
    }
    const int ncol = jn[0] * jn[1] * jn[2];
    // ---- phase 1: t_j = sum_i P[i,I] A[i,j] / pn_j for every column of the box
    for (int c = lane; c < ncol; c += 32) {
        double t = 0.0;
        const int ji = jlo[0] + c % jn[0], jj = jlo[1] + (c / jn[0]) % jn[1], jk = jlo[2] + c / (jn[0] * jn[1]);
        const int idj = gidx(gf, ji, jj, jk);
        const size_t oj = mp * Tf + idj;
        float pnj = 0.0f;
        if (diag_f[oj] != 0.0f) pnj = pn_f[oj];       // a column that is not an unknown drops out (pnj stays 0)
        if (pnj != 0.0f) {
            if (FINE0) {
                int off[7][3];
                double val[7];
                int ne = gmg_entries_l0(mp, m, idj, sy, sz, Tf, coef_f, S_f /* level 0: the volume grids */, off, val);
                for (int e = 0; e < ne; e++) {
                    int a = ji + off[e][0] - lo[0], b2 = jj + off[e][1] - lo[1], c2 = jk + off[e][2] - lo[2];
                    if (a < 0 || b2 < 0 || c2 < 0 || a >= cn[0] || b2 >= cn[1] || c2 >= cn[2]) continue;
                    t += (double)wP[(c2 * 4 + b2) * 4 + a] * val[e];
                }
            } else {
                const float *__restrict__ Sj = S_f + (size_t)rowmap_f[oj] * GMG_STRIDE + R.base;
                int a0 = ji + olo[0] - lo[0], a1 = ji + ohi[0] - lo[0];
                int b0 = jj + olo[1] - lo[1], b1 = jj + ohi[1] - lo[1];
                int c0 = jk + olo[2] - lo[2], c1 = jk + ohi[2] - lo[2];
                a0 = a0 < 0 ? 0 : a0; b0 = b0 < 0 ? 0 : b0; c0 = c0 < 0 ? 0 : c0;
                a1 = a1 > cn[0] - 1 ? cn[0] - 1 : a1; b1 = b1 > cn[1] - 1 ? cn[1] - 1 : b1; c1 = c1 > cn[2] - 1 ? cn[2] - 1 : c1;
                for (int c2 = c0; c2 <= c1; c2++)
                    for (int b2 = b0; b2 <= b1; b2++) {
                        const float *__restrict__ Sr = Sj + ((lo[2] + c2 - jk - R.lo[2]) * R.n[1] + (lo[1] + b2 - jj - R.lo[1])) * R.n[0] +
                                                       (lo[0] - ji - R.lo[0]);
                        for (int a = a0; a <= a1; a++) t += (double)wP[(c2 * 4 + b2) * 4 + a] * (double)Sr[a];
                    }
            }
            t /= (double)pnj;
        }
        tb[c] = t;
    }
    __syncwarp();
    // ---- phase 2: slot J = sum over the children j of J (component mp) of t_j P[j,J], ascending column order
    for (int q = lane; q < W.size; q += 32) {
        const int PI = I + W.lo[0] + q % W.n[0], PJ = J + W.lo[1] + (q / W.n[0]) % W.n[1], PK = K + W.lo[2] + q / (W.n[0] * W.n[1]);
        double acc = 0.0;
        if (PI >= 0 && PJ >= 0 && PK >= 0 && PI <= gc.ni && PJ <= gc.nj && PK <= gc.nk && diag_c[(size_t)mp * Tc + gidx(gc, PI, PJ, PK)] != 0.0f) {
            for (int c2 = 0; c2 < 4; c2++) {
                const float wk = gmg_child_weight(mp == 2, c2);
                const int kk = 2 * PK - 1 + c2 - jlo[2];
                if (wk == 0.0f || kk < 0 || kk >= jn[2]) continue;
                for (int b2 = 0; b2 < 4; b2++) {
                    const float wj = gmg_child_weight(mp == 1, b2);
                    const int jj = 2 * PJ - 1 + b2 - jlo[1];
                    if (wj == 0.0f || jj < 0 || jj >= jn[1]) continue;
                    const double *__restrict__ trow = tb + (kk * jn[1] + jj) * jn[0];
                    for (int a = 0; a < 4; a++) {
                        const float wi = gmg_child_weight(mp == 0, a);
                        const int ii = 2 * PI - 1 + a - jlo[0];
                        if (wi == 0.0f || ii < 0 || ii >= jn[0]) continue;
                        acc += trow[ii] * (double)(wi * wj * wk);
                    }
                }
            }
        }
        S_c[(size_t)r * GMG_STRIDE + W.base + q] = (float)(0.125 * acc);
    }
}

// lower blocks of the symmetric A_c: entry (row (m; J), column (mp; J + e)) with mp < m is entry
// (row (mp; J + e), column (m; J)) of an upper block; the windows of (m, mp) and (mp, m) are mirror images.
// One thread per (row, slot of a lower block).
__global__ void __launch_bounds__(256) k_gmg_mirror(Grid g, const int *__restrict__ rows, int nrows, const int *__restrict__ rowmap,
                                                     float *__restrict__ S) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int r = (int)(t / 160), q = (int)(t - (long long)r * 160);
    if (r >= nrows) return;
    const int enc = rows[r];
    const int m = enc / g.total, id = enc - m * g.total;
    if (m == 0) return;                       // rows of component 0 have no lower block
    const int mp = q / 80, ql = q - mp * 80;  // lower blocks: column components 0 .. m-1, 80 slots each
    if (mp >= m) return;
    const GWin W = gmg_window(m, mp);
    const int ei = W.lo[0] + ql % W.n[0], ej = W.lo[1] + (ql / W.n[0]) % W.n[1], ek = W.lo[2] + ql / (W.n[0] * W.n[1]);
    float v = 0.0f;
    const int ro = rowmap[mp * g.total + id + ei + ej * SY(g) + ek * SZ(g)];
    if (ro >= 0) {
        const GWin U = gmg_window(mp, m);     // the upper block's window holds the offset -e
        const int slot = U.base + ((-ek - U.lo[2]) * U.n[1] + (-ej - U.lo[1])) * U.n[0] + (-ei - U.lo[0]);
        v = S[(size_t)ro * GMG_STRIDE + slot];
    }
    S[(size_t)r * GMG_STRIDE + W.base + ql] = v;
}

// dense diagonal of an explicit level (also its unknown flag) = the (0,0,0) slot of the (m,m) window, and the
// per-row smoothing weight  w = min(omega / a_ii, GMG_L1_BOUND / sum_j |a_ij|).  Galerkin rows next to the free
// surface can have sum|a_ij| / a_ii >> 2, and lambda_max(D^-1 A_c) was measured at 3.3 - 4.3 on the coarse levels:
// plain damped Jacobi with omega = 0.5 then amplifies a few modes (omega * lambda > 2), the V-cycle turns
// indefinite and CG wanders for hundreds of iterations.  With the l1 cap, lambda_max(W A_c) <= GMG_L1_BOUND < 2
// by Gershgorin, for any geometry.  One warp per row.
#define GMG_L1_BOUND 1.6f
__global__ void __launch_bounds__(256) k_gmg_diag(Grid g, const int *__restrict__ rows, int nrows, const float *__restrict__ S,
                                                   float *__restrict__ diag, float *__restrict__ wj, float omega) {
    int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= nrows) return;
    int enc = rows[r];
    int m = enc / g.total;
    GWin W = gmg_window(m, m);
    int dslot = W.base + ((-W.lo[2]) * W.n[1] + (-W.lo[1])) * W.n[0] + (-W.lo[0]);
    const float *Sr = S + (size_t)r * GMG_STRIDE;
    float l1 = 0.0f;
    for (int q = lane; q < GMG_SLOTS; q += 32) l1 += fabsf(Sr[q]);
    for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    if (lane == 0) {
        float d = Sr[dslot];
        diag[enc] = d > 1e-30f ? d : 1e-30f;   // stays an unknown; a degenerate row is simply not smoothed
        float w = d > 1e-20f ? omega / d : 0.0f;
        if (l1 > 0.0f) w = fminf(w, GMG_L1_BOUND / l1);
        wj[r] = w;
    }
}

// offset (in elements, column component included) of slot `slot` of a row of component m; 0 for padding
FLIP_HD int gmg_slot_offset(const Grid &g, int m, int slot) {
    for (int mp = 0; mp < 3; mp++) {
        GWin W = gmg_window(m, mp);
        if (slot < W.base + W.size) {
            int q = slot - W.base;
            int di = W.lo[0] + q % W.n[0], dj = W.lo[1] + (q / W.n[0]) % W.n[1], dk = W.lo[2] + q / (W.n[0] * W.n[1]);
            return mp * g.total + di + dj * SY(g) + dk * SZ(g);
        }
    }
    return 0;
}

// Prologue loads of the multigrid kernels.  The coarse levels are latency bound (a few thousand rows: one kernel = a
// chain of dependent global loads, ~0.7 us each), so the values that do not depend on each other - the CG's done flag, the
// row range / cell count, the offset table - are requested together, pinned in program order by volatile asm (the
// compiler otherwise sinks them behind the early return), instead of one round trip after the other.
#ifdef FLIP_CPU_EMU
FLIP_D int ld_early(const int *p) { return *p; }
#else
FLIP_D int ld_early(const int *p) {
    int v;
    asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#endif
FLIP_D int ld_done(const CGState *st) { return st ? ld_early(&st->done) : 0; }

// one warp: row r of an explicit level.  MODE 1: out = xi + w (b - A xi)   2: out = (b - A xi) / pn.
// offs: the level's slot -> element offset table [3][GMG_STRIDE] (shared or global memory)
// L2: the vectors are read with L2-only loads (persistent kernels: other CTAs wrote them earlier in the same launch, and
// loads through const __restrict__ pointers may take the non-coherent path, which a fence does not invalidate)
template <int MODE, bool L2 = false>
FLIP_D void gmg_row_apply(const Grid &g, const int *__restrict__ rows, const float *__restrict__ S, int stride,
                          const float *__restrict__ wj, const float *b, const float *xi, float *out, const float *__restrict__ pn,
                          const int *offs, int r, int lane) {
    const int enc = rows[r];
    const int m = enc / g.total, id = enc - m * g.total;
    const float *__restrict__ Sr = S + (size_t)r * stride;   // stride: GMG_STRIDE, or GMG_CSTRIDE for compact rows
    const float *xc = xi + id;
    const int *om = offs + m * GMG_STRIDE;                   // the shared table keeps its pitch
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        int slot = lane + 32 * t;
        if (slot < stride) acc += Sr[slot] * (L2 ? ld_cg(xc + om[slot]) : xc[om[slot]]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float bv = L2 ? ld_cg(b + enc) : b[enc];
        if (MODE == 1) out[enc] = (L2 ? ld_cg(xi + enc) : xi[enc]) + wj[r] * (bv - acc);
        else { float p = pn[enc]; out[enc] = p > 0.0f ? (bv - acc) / p : 0.0f; }
    }
}

// One warp per row.  mode 0: xo = omega b / d (first sweep from zero)   1: xo = xi + omega (b - A xi) / d
// 2: ro = (b - A xi) / pn  (residual, pre-scaled for the restriction)
template <int MODE>
__global__ void __launch_bounds__(256) k_gmg_sweep(GLevelDev L, const float *__restrict__ b, const float *__restrict__ xi,
                                                    float *__restrict__ out, const float *__restrict__ pn, float omega,
                                                    const CGState *__restrict__ st) {
    __shared__ int offs[3][GMG_STRIDE];
    if (st && st->done) return;
    if (!xch_enter(L.X, L.X.nbr != 0)) return;
    const bool compact = L.cmeta != nullptr && L.cmeta[0] != 0;
    const int stride = compact ? GMG_CSTRIDE : GMG_STRIDE;
    const float *__restrict__ Sb = compact ? L.Sc : L.S;
    if (MODE != 0) {
        const int *__restrict__ ot = compact ? L.coffs : L.offs;
        for (int q = threadIdx.x; q < 3 * stride; q += blockDim.x) offs[q / stride][q % stride] = ot[q];
        __syncthreads();
    }
    const int r0 = L.rng[0], r1 = L.rng[1];
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int r = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < r1; r += nwarps) {
        if (MODE == 0) {
            if (lane == 0) { int enc = L.rows[r]; out[enc] = L.wj[r] * b[enc]; }
            continue;
        }
        gmg_row_apply<MODE == 2 ? 2 : 1>(L.g, L.rows, Sb, stride, L.wj, b, xi, out, pn, &offs[0][0], r, lane);
    }
    xch_leave(L.X, false);
}

#ifndef FLIP_CPU_EMU
// ---- the same sweep with the coefficient rows staged by TMA ---------------------------------------------------------
// k_gmg_sweep is bound by L1 (ncu: l1tex 79 %): every row gathers 240 vector values through it, and its own 960-byte
// coefficient row competes for the same cache.  Here lane 0 of every warp issues ONE bulk asynchronous copy
// (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) per row, two rows in flight per warp, so the
// coefficient stream goes L2 -> shared memory without touching L1 or the register file, and the gathers keep the cache.
// Same arithmetic, same order: bit-identical to k_gmg_sweep.
FLIP_D unsigned gmg_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
FLIP_D void gmg_bulk_load(float *dst_smem, const float *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gmg_smem(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gmg_smem(dst_smem)),
                 "l"(src), "r"(bytes), "r"(gmg_smem(bar))
                 : "memory");
}
FLIP_D void gmg_bar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GMG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GMG_DONE_%=;\n"
        "bra GMG_WAIT_%=;\n"
        "GMG_DONE_%=:\n"
        "}\n" ::"r"(gmg_smem(bar)),
        "r"(parity)
        : "memory");
}

template <int MODE>   // MODE 1 or 2
__global__ void __launch_bounds__(256) k_gmg_sweep_tma(GLevelDev L, const float *__restrict__ b, const float *__restrict__ xi,
                                                        float *__restrict__ out, const float *__restrict__ pn, float omega,
                                                        const CGState *__restrict__ st) {
    __shared__ int offs[3][GMG_STRIDE];
    __shared__ __align__(128) float Srow[8][2][GMG_STRIDE];
    __shared__ __align__(8) unsigned long long bars[8][2];
    const int done = ld_done(st), r0 = ld_early(L.rng), r1 = ld_early(L.rng + 1);
    const bool compact = L.cmeta != nullptr && ld_early(L.cmeta) != 0;
    const int stride = compact ? GMG_CSTRIDE : GMG_STRIDE;
    const float *__restrict__ Sb = compact ? L.Sc : L.S;
    const unsigned row_bytes = (unsigned)stride * 4u;
    {
        const int *__restrict__ ot = compact ? L.coffs : L.offs;
        for (int q = threadIdx.x; q < 3 * stride; q += blockDim.x) offs[q / stride][q % stride] = ot[q];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gmg_smem(&bars[wid][0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gmg_smem(&bars[wid][1])) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the async proxy sees the initialised barriers
    }
    if (done) return;
    if (!xch_enter(L.X, L.X.nbr != 0)) return;
    __syncthreads();
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    int r = r0 + blockIdx.x * (blockDim.x >> 5) + wid;
    unsigned phases = 0u;   // bit q: the parity the next wait on barrier q expects (a bit mask: an indexed array would live in local memory)
    int buf = 0;
    if (r < r1 && lane == 0) gmg_bulk_load(Srow[wid][0], Sb + (size_t)r * stride, row_bytes, &bars[wid][0]);
    int enc = r < r1 ? L.rows[r] : 0;
    for (; r < r1; r += nwarps) {
        const int rn = r + nwarps;
        if (rn < r1 && lane == 0) gmg_bulk_load(Srow[wid][buf ^ 1], Sb + (size_t)rn * stride, row_bytes, &bars[wid][buf ^ 1]);
        // A warp works on one row at a time, so a row costs the SUM of its dependent memory round trips.  Everything that
        // does not depend on the product is therefore requested together, before the wait for the coefficient row: the
        // gathers, the operands of the row update (lane 0) and the index of the NEXT row.
        const int enc_next = rn < r1 ? L.rows[rn] : 0;
        const int m = enc / L.g.total, id = enc - m * L.g.total;
        const float *__restrict__ xc = xi + id;
        float xv[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int slot = lane + 32 * t;
            xv[t] = slot < stride ? xc[offs[m][slot]] : 0.0f;
        }
        float bv = 0.0f, u0 = 0.0f, u1 = 0.0f;
        if (lane == 0) {
            bv = b[enc];
            if (MODE == 1) { u0 = xi[enc]; u1 = L.wj[r]; }
            else u0 = pn[enc];
        }
        gmg_bar_wait(&bars[wid][buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
        const float *Sr = Srow[wid][buf];
        float acc = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int slot = lane + 32 * t;
            if (slot < stride) acc += Sr[slot] * xv[t];
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (MODE == 1) out[enc] = u0 + u1 * (bv - acc);
            else out[enc] = u0 > 0.0f ? (bv - acc) / u0 : 0.0f;
        }
        __syncwarp();   // every lane is done with this buffer before the copy after next overwrites it
        buf ^= 1;
        enc = enc_next;
    }
    xch_leave(L.X, false);
}
#endif

// ---- compact rows of the first explicit level ------------------------------------------------------------------------
// Level 1 is the Galerkin product of the level-0 STENCIL, which is much sparser than a general row of the hierarchy: at
// most 159 of its 235 slots are ever non-zero, the same ones for every row of a component (measured with
// dev/gmg_fill.py; deeper levels fill all 235).  After the set-up the rows of that level are therefore copied into a second
// array with GMG_CSTRIDE = 160 slots per row (640 bytes: one bulk copy) and their own slot -> offset table, and the sweeps
// read those: a third fewer coefficient bytes and - what bounds the kernel - a third fewer gathers per row.  The kept slots
// are found on the device for every solve (union of the non-zero slots over the rows of a component); should a component
// ever need more than GMG_CSTRIDE, cmeta[0] stays 0 and the sweeps keep reading the full rows.  The full rows stay what the
// set-up of the next level, the mirror pass and the diagnostics read.
__global__ void __launch_bounds__(256) k_gmg_slot_mask(const int *__restrict__ rows, const int *__restrict__ nrows, int T,
                                                        const float *__restrict__ S, unsigned *__restrict__ mask) {
    const int n = *nrows;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    unsigned bits0 = 0u, bits1 = 0u, bits2 = 0u;   // bit t: slot lane + 32 t is non-zero in some row of component 0 / 1 / 2
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
        const int m = rows[r] / T;
        const float *__restrict__ Sr = S + (size_t)r * GMG_STRIDE;
        unsigned b = 0u;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int slot = lane + 32 * t;
            if (slot < GMG_SLOTS && Sr[slot] != 0.0f) b |= 1u << t;
        }
        if (m == 0) bits0 |= b; else if (m == 1) bits1 |= b; else bits2 |= b;
    }
    for (int t = 0; t < 8; t++) {
        unsigned w0 = __ballot_sync(0xffffffffu, (bits0 >> t) & 1u), w1 = __ballot_sync(0xffffffffu, (bits1 >> t) & 1u),
                 w2 = __ballot_sync(0xffffffffu, (bits2 >> t) & 1u);
        if (lane == 0) {
            if (w0) atomicOr(&mask[t], w0);
            if (w1) atomicOr(&mask[8 + t], w1);
            if (w2) atomicOr(&mask[16 + t], w2);
        }
    }
}

// one CTA of 96 threads: warp m turns the mask of component m into the compact slot list (ascending slot order)
__global__ void __launch_bounds__(96) k_gmg_compact_table(const unsigned *__restrict__ mask, const int *__restrict__ offs,
                                                           int *__restrict__ coffs, int *__restrict__ cslot, int *__restrict__ cmeta) {
    __shared__ int kept[3];
    const int m = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int base = 0;
    for (int t = 0; t < 8; t++) {
        const unsigned w = mask[m * 8 + t];
        const int pos = base + __popc(w & ((1u << lane) - 1u));
        if (((w >> lane) & 1u) && pos < GMG_CSTRIDE) {
            const int slot = lane + 32 * t;
            cslot[m * GMG_CSTRIDE + pos] = slot;
            coffs[m * GMG_CSTRIDE + pos] = offs[m * GMG_STRIDE + slot];
        }
        base += __popc(w);
    }
    for (int j = base + lane; j < GMG_CSTRIDE; j += 32) { cslot[m * GMG_CSTRIDE + j] = -1; coffs[m * GMG_CSTRIDE + j] = 0; }
    if (lane == 0) { kept[m] = base; cmeta[1 + m] = base; }
    __syncthreads();
    if (threadIdx.x == 0) cmeta[0] = (kept[0] <= GMG_CSTRIDE && kept[1] <= GMG_CSTRIDE && kept[2] <= GMG_CSTRIDE) ? 1 : 0;
}

// one warp per row: Sc[r][j] = S[r][cslot[m][j]]
__global__ void __launch_bounds__(256) k_gmg_compact_rows(const int *__restrict__ rows, const int *__restrict__ nrows, int T,
                                                           const float *__restrict__ S, const int *__restrict__ cslot,
                                                           const int *__restrict__ cmeta, float *__restrict__ Sc) {
    if (cmeta[0] == 0) return;
    const int n = *nrows;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
        const int m = rows[r] / T;
        const float *__restrict__ Sr = S + (size_t)r * GMG_STRIDE;
        for (int j = lane; j < GMG_CSTRIDE; j += 32) {
            const int sl = cslot[m * GMG_CSTRIDE + j];
            Sc[(size_t)r * GMG_CSTRIDE + j] = sl >= 0 ? Sr[sl] : 0.0f;
        }
    }
}

// ---- x-groups of the first explicit level ---------------------------------------------------------------------------
// A group = the rows of one 8-cell x-run of a tile (same component, j, k): consecutive rows of the k-plane-major row
// list, {first row, 8-bit mask of the cells that are unknowns}.  The restriction from level 0 runs over groups: lane
// (xr, a) = (coarse row of the run, child offset along x), so the 32 lanes of a load read 18 consecutive fine values
// (one or two cache lines) where the warp-per-row kernel touches 8 scattered lines: 85 -> 29 us at 256^3 on B200.
// The same lane mapping was tried for the Jacobi sweeps (a warp walks the 65-82 x-lines of a group's window, coefficient
// blocks staged by TMA): 116 us against 80 us for the warp-per-row sweep below - with one 7.6 KB coefficient block per
// warp the SM holds 24 warps, too few to hide the latency of ~10 dependent load batches per group.  Not kept.
// coarse b = P^T r / 8 and the first sweep from zero (k_gmg_restrict_first) over the x-groups of the coarse level: lane
// (xr, a) reads fine x = 2 (i0 + xr) - 1 + a of every (fj, fk) line of the 4 x 4 children box, so a load covers 18
// consecutive fine values instead of 8 scattered lines
__global__ void __launch_bounds__(256) k_gmg_restrict_x(GLevelDev C, Grid gf, const float *__restrict__ rf, float *__restrict__ bc,
                                                         float *__restrict__ x0, const CGState *__restrict__ st) {
    const int done = ld_done(st), g0 = ld_early(C.grng), g1 = ld_early(C.grng + 1);
    if (done) return;
    if (!xch_enter(C.X, C.X.nbr != 0)) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int T = C.g.total;
    const int xr = lane & 7, a = lane >> 3;
    for (int g = g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < g1; g += nwarps) {
        const int2 G = C.groups[g];
        const int r0 = G.x;
        const unsigned mask = (unsigned)G.y;
        const int enc0 = C.rows[r0];
        const int m = enc0 / T;
        const int encb = enc0 - (__ffs(mask) - 1);
        int I0, J, K;
        gmg_unflatten(C.g, encb - m * T, I0, J, K);
        const bool has = (mask >> xr) & 1u;
        const int fi = 2 * (I0 + xr) - 1 + a;
        const float wi = gmg_child_weight(m == 0, a);
        const float *rr = rf + (size_t)m * gf.total;
        const bool xin = has && wi != 0.0f && fi >= 0 && fi <= gf.ni;
        float acc = 0.0f;
        for (int c2 = 0; c2 < 4; c2++) {
            const int fk = 2 * K - 1 + c2;
            const float wk = gmg_child_weight(m == 2, c2);
            if (wk == 0.0f || fk < 0 || fk > gf.nk) continue;
            for (int b2 = 0; b2 < 4; b2++) {
                const int fj = 2 * J - 1 + b2;
                const float wj = gmg_child_weight(m == 1, b2);
                if (wj == 0.0f || fj < 0 || fj > gf.nj) continue;
                if (xin) acc += wi * wj * wk * rr[gidx(gf, fi, fj, fk)];
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (has && a == 0) {
            const int enc = encb + xr, r = r0 + __popc(mask & ((1u << xr) - 1u));
            const float bv = 0.125f * acc;
            bc[enc] = bv;
            x0[enc] = C.wj[r] * bv;
        }
    }
    xch_leave(C.X, false);
}

// ---- level 0 on the solver's compact cell list ------------------------------------------------
// Same loads as k_visc_apply (viscosity.cu), in fp32: one thread per cell index that holds an unknown,
// all three face rows at once.  MODE 0: x = omega b / d.  1: x = xi + omega (b - A xi) / d.
// 2: r = (b - A xi) / pn (residual, pre-scaled for the restriction).  3: like 1, and the result is the
// preconditioned residual z handed back to the CG in fp64.
struct G0Params {
    Grid g;
    const int *cell_list, *cell_count;
    const float *coef, *diag, *pn;
    const float *vol;     // the solver's 7 volume grids (vvol); faces U,V,W are grids 1,2,3
    Xch X;                // exchange context (the cell list holds this rank's slab only)
};

template <int MODE>
__global__ void __launch_bounds__(256) k_gmg0_sweep(G0Params L, const double *__restrict__ b, const float *__restrict__ xi,
                                                     float *__restrict__ out, double *__restrict__ zout, float omega,
                                                     const CGState *__restrict__ st) {
    const int done = ld_done(st), nc = ld_early(L.cell_count);
    if (done) return;
    if (MODE != 0 && !xch_enter(L.X, L.X.nbr != 0)) return;   // mode 0 reads and writes this rank's own cells only
    const Grid &g = L.g;
    const int sy = SY(g), sz = SZ(g);
    const size_t T = (size_t)g.total;
    const float *__restrict__ cc = L.coef, *__restrict__ cu = L.coef + T, *__restrict__ cv = L.coef + 2 * T, *__restrict__ cw = L.coef + 3 * T;
    for (int qq = blockIdx.x * blockDim.x + threadIdx.x; qq < nc; qq += gridDim.x * blockDim.x) {
        const int id = L.cell_list[qq];
        const float dU = L.diag[id], dV = L.diag[T + id], dW = L.diag[2 * T + id];
        const float bU = (float)b[id], bV = (float)b[T + id], bW = (float)b[2 * T + id];
        if (MODE == 0) {
            if (dU != 0.0f) out[id] = omega * bU / dU;
            if (dV != 0.0f) out[T + id] = omega * bV / dV;
            if (dW != 0.0f) out[2 * T + id] = omega * bW / dW;
            continue;
        }
        const float *__restrict__ su = xi, *__restrict__ sv = xi + T, *__restrict__ sw = xi + 2 * T;
        const float u0 = su[id], u_xp = su[id + 1], u_xm = su[id - 1], u_yp = su[id + sy], u_ym = su[id - sy],
                    u_zp = su[id + sz], u_zm = su[id - sz], u_xp_ym = su[id + 1 - sy], u_xp_zm = su[id + 1 - sz];
        const float v0 = sv[id], v_xp = sv[id + 1], v_xm = sv[id - 1], v_yp = sv[id + sy], v_ym = sv[id - sy],
                    v_zp = sv[id + sz], v_zm = sv[id - sz], v_xm_yp = sv[id - 1 + sy], v_yp_zm = sv[id + sy - sz];
        const float w0 = sw[id], w_xp = sw[id + 1], w_xm = sw[id - 1], w_yp = sw[id + sy], w_ym = sw[id - sy],
                    w_zp = sw[id + sz], w_zm = sw[id - sz], w_xm_zp = sw[id - 1 + sz], w_ym_zp = sw[id - sy + sz];
        const float c0 = cc[id], c_xm = cc[id - 1], c_ym = cc[id - sy], c_zm = cc[id - sz];
        const float eu0 = cu[id], eu_yp = cu[id + sy], eu_zp = cu[id + sz];
        const float ev0 = cv[id], ev_xp = cv[id + 1], ev_zp = cv[id + sz];
        const float ew0 = cw[id], ew_xp = cw[id + 1], ew_yp = cw[id + sy];
        float aU = 0.0f, aV = 0.0f, aW = 0.0f;
        if (dU != 0.0f) {
            const float fR = c0, fL = c_xm, fT = ew_yp, fB = ew0, fF = ev_zp, fK = ev0;
            aU = L.vol[T + id] * u0 + fR * (u0 - u_xp) + fL * (u0 - u_xm) + fT * (u0 - u_yp) + fB * (u0 - u_ym) + fF * (u0 - u_zp) +
                 fK * (u0 - u_zm) - fT * (v_yp - v_xm_yp) + fB * (v0 - v_xm) - fF * (w_zp - w_xm_zp) + fK * (w0 - w_xm);
        }
        if (dV != 0.0f) {
            const float fR = ew_xp, fL = ew0, fT = c0, fB = c_ym, fF = eu_zp, fK = eu0;
            aV = L.vol[2 * T + id] * v0 + fR * (v0 - v_xp) + fL * (v0 - v_xm) + fT * (v0 - v_yp) + fB * (v0 - v_ym) + fF * (v0 - v_zp) +
                 fK * (v0 - v_zm) - fR * (u_xp - u_xp_ym) + fL * (u0 - u_ym) - fF * (w_zp - w_ym_zp) + fK * (w0 - w_ym);
        }
        if (dW != 0.0f) {
            const float fR = ev_xp, fL = ev0, fT = eu_yp, fB = eu0, fF = c0, fK = c_zm;
            aW = L.vol[3 * T + id] * w0 + fR * (w0 - w_xp) + fL * (w0 - w_xm) + fT * (w0 - w_yp) + fB * (w0 - w_ym) + fF * (w0 - w_zp) +
                 fK * (w0 - w_zm) - fR * (u_xp - u_xp_zm) + fL * (u0 - u_zm) - fT * (v_yp - v_yp_zm) + fB * (v0 - v_zm);
        }
        if (MODE == 2) {
            const float pU = L.pn[id], pV = L.pn[T + id], pW = L.pn[2 * T + id];
            if (dU != 0.0f) out[id] = pU > 0.0f ? (bU - aU) / pU : 0.0f;
            if (dV != 0.0f) out[T + id] = pV > 0.0f ? (bV - aV) / pV : 0.0f;
            if (dW != 0.0f) out[2 * T + id] = pW > 0.0f ? (bW - aW) / pW : 0.0f;
        } else {
            const float xU = dU != 0.0f ? u0 + omega * (bU - aU) / dU : 0.0f;
            const float xV = dV != 0.0f ? v0 + omega * (bV - aV) / dV : 0.0f;
            const float xW = dW != 0.0f ? w0 + omega * (bW - aW) / dW : 0.0f;
            if (MODE == 3) { zout[id] = (double)xU; zout[T + id] = (double)xV; zout[2 * T + id] = (double)xW; }
            else {
                if (dU != 0.0f) out[id] = xU;
                if (dV != 0.0f) out[T + id] = xV;
                if (dW != 0.0f) out[2 * T + id] = xW;
            }
        }
    }
    if (MODE != 0) xch_leave(L.X, false);
}

// value of the coarse correction P x_c at fine face (m; i,j,k), not yet divided by pn
template <bool L2 = false>
FLIP_D float gmg_interp(int m, int i, int j, int k, const Grid &gc, const float *xc) {
    int pi[2], pj[2], pk[2];
    float wi[2], wj[2], wk[2];
    gmg_parents(m == 0, i, pi[0], pi[1], wi[0], wi[1]);
    gmg_parents(m == 1, j, pj[0], pj[1], wj[0], wj[1]);
    gmg_parents(m == 2, k, pk[0], pk[1], wk[0], wk[1]);
    const float *x = xc + (size_t)m * gc.total;
    float v = 0.0f;
    for (int c2 = 0; c2 < 2; c2++)
        for (int b2 = 0; b2 < 2; b2++)
            for (int a = 0; a < 2; a++) {
                float w = wi[a] * wj[b2] * wk[c2];
                if (w == 0.0f) continue;
                int I = pi[a], J = pj[b2], K = pk[c2];
                if (I < 0 || J < 0 || K < 0 || I > gc.ni || J > gc.nj || K > gc.nk) continue;
                v += w * (L2 ? ld_cg(x + gidx(gc, I, J, K)) : x[gidx(gc, I, J, K)]);   // x_c is exactly 0 where the coarse face is not an unknown
            }
    return v;
}

// level 0: x += P x_c on the cell list
__global__ void __launch_bounds__(256) k_gmg0_prolong(G0Params F, Grid gc, const float *__restrict__ xc, float *__restrict__ xf,
                                                       const CGState *__restrict__ st) {
    const int done = ld_done(st), nc = ld_early(F.cell_count);
    if (done) return;
    if (!xch_enter(F.X, F.X.nbr != 0)) return;
    const Grid &g = F.g;
    const size_t T = (size_t)g.total;
    for (int qq = blockIdx.x * blockDim.x + threadIdx.x; qq < nc; qq += gridDim.x * blockDim.x) {
        const int id = F.cell_list[qq];
        int i, j, k;
        gmg_unflatten(g, id, i, j, k);
        for (int m = 0; m < 3; m++) {
            float p = F.pn[m * T + id];
            if (p > 0.0f) xf[m * T + id] += gmg_interp(m, i, j, k, gc, xc) / p;
        }
    }
    xch_leave(F.X, false);
}

// explicit levels: x += P x_c, one thread per fine row
__global__ void __launch_bounds__(256) k_gmg_prolong(GLevelDev F, const float *__restrict__ pn, Grid gc, const float *__restrict__ xc,
                                                      float *__restrict__ xf, const CGState *__restrict__ st) {
    const int done = ld_done(st), r0 = ld_early(F.rng), r1 = ld_early(F.rng + 1);
    if (done) return;
    for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x) {
        int enc = F.rows[r];
        int m = enc / F.g.total, id = enc - m * F.g.total;
        float p = pn[enc];
        if (p <= 0.0f) continue;
        int i, j, k;
        gmg_unflatten(F.g, id, i, j, k);
        xf[enc] += gmg_interp(m, i, j, k, gc, xc) / p;
    }
}

// coarse b = P^T r / 8 (r already divided by pn) and the first sweep from zero, x0 = w b.  One warp per coarse
// row: the <= 48 fine children are spread over the lanes (two per lane), so a row costs one memory round trip
// instead of 48 serialised ones.
template <bool L2 = false>
FLIP_D void gmg_row_restrict(const Grid &gc, const int *__restrict__ rows, const float *__restrict__ wjc, const Grid &gf,
                             const float *rf, float *bc, float *x0, int r, int lane) {
    const int enc = rows[r];
    const int m = enc / gc.total, id = enc - m * gc.total;
    int I, J, K;
    gmg_unflatten(gc, id, I, J, K);
    const float *rr = rf + (size_t)m * gf.total;
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < 2; t++) {
        int q = lane + 32 * t;              // child (a, b2, c2) in a 4 x 4 x 4 box, a fastest
        int a = q & 3, b2 = (q >> 2) & 3, c2 = q >> 4;
        // along a component's own axis: fine 2I-1, 2I, 2I+1 with 1/2, 1, 1/2; across it: 2J-1 .. 2J+2 with 1/4, 3/4, 3/4, 1/4
        int fi = 2 * I - 1 + a, fj = 2 * J - 1 + b2, fk = 2 * K - 1 + c2;
        float wi = m == 0 ? (a == 1 ? 1.0f : (a < 3 ? 0.5f : 0.0f)) : ((a == 1 || a == 2) ? 0.75f : 0.25f);
        float wj = m == 1 ? (b2 == 1 ? 1.0f : (b2 < 3 ? 0.5f : 0.0f)) : ((b2 == 1 || b2 == 2) ? 0.75f : 0.25f);
        float wk = m == 2 ? (c2 == 1 ? 1.0f : (c2 < 3 ? 0.5f : 0.0f)) : ((c2 == 1 || c2 == 2) ? 0.75f : 0.25f);
        float w = wi * wj * wk;
        if (w != 0.0f && fi >= 0 && fj >= 0 && fk >= 0 && fi <= gf.ni && fj <= gf.nj && fk <= gf.nk)
            acc += w * (L2 ? ld_cg(rr + gidx(gf, fi, fj, fk)) : rr[gidx(gf, fi, fj, fk)]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        float bv = 0.125f * acc;
        bc[enc] = bv;
        x0[enc] = wjc[r] * bv;
    }
}

__global__ void __launch_bounds__(256) k_gmg_restrict_first(GLevelDev C, Grid gf, const float *__restrict__ rf, float *__restrict__ bc,
                                                             float *__restrict__ x0, const CGState *__restrict__ st) {
    const int done = ld_done(st), r0 = ld_early(C.rng), r1 = ld_early(C.rng + 1);
    if (done) return;
    if (!xch_enter(C.X, C.X.nbr != 0)) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int r = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < r1; r += nwarps)
        gmg_row_restrict(C.g, C.rows, C.wj, gf, rf, bc, x0, r, lane);
    xch_leave(C.X, false);
}

// ---- coarsest level: dense inverse -------------------------------------------------------------
// The hierarchy stops at the first level with <= mg_dense_rows (128) rows - the 4^3 level, 50 rows at 256^3 - and
// solves it exactly with a dense inverse computed once per solve: one launch per V-cycle instead of 25
// latency-bound sweeps, and an exact (hence symmetric) coarse solve.
#define GMG_DENSE_MAX 512

// single CTA: dense A (fp64, symmetrised) from the stored rows, in-place Gauss-Jordan inversion (A is SPD: no
// pivoting), result as fp32
__global__ void __launch_bounds__(1024) k_gmg_dense_inverse(GLevelDev L, const int *__restrict__ rowmap, double *__restrict__ D,
                                                             float *__restrict__ Ainv) {
    const int n = *L.nrows;
    const int T = L.g.total;
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) D[q] = 0.0;
    __syncthreads();
    for (int q = threadIdx.x; q < n * GMG_SLOTS; q += blockDim.x) {
        int r = q / GMG_SLOTS, slot = q - r * GMG_SLOTS;
        float v = L.S[(size_t)r * GMG_STRIDE + slot];
        if (v == 0.0f) continue;
        int enc = L.rows[r];
        int m = enc / T, id = enc - m * T;
        int c = rowmap[id + L.offs[m * GMG_STRIDE + slot]];
        if (c >= 0) D[r * n + c] = (double)v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) {       // symmetrise (P^T A P is symmetric up to rounding)
        int i = q / n, j = q - i * n;
        if (i < j) { double a = 0.5 * (D[i * n + j] + D[j * n + i]); D[i * n + j] = a; D[j * n + i] = a; }
    }
    __syncthreads();
    __shared__ double prow[GMG_DENSE_MAX], pcol[GMG_DENSE_MAX];
    for (int k = 0; k < n; k++) {
        const double p = 1.0 / D[k * n + k];
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            prow[j] = j == k ? p : D[k * n + j] * p;
            pcol[j] = j == k ? 0.0 : D[j * n + k];
        }
        __syncthreads();
        for (int q = threadIdx.x; q < n * n; q += blockDim.x) {
            int i = q / n, j = q - i * n;
            double v;
            if (i == k) v = prow[j];
            else if (j == k) v = -pcol[i] * p;
            else v = D[q] - pcol[i] * prow[j];
            D[q] = v;
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) Ainv[q] = (float)D[q];
}

// x = A^-1 b on the coarsest level, one warp per row
__global__ void __launch_bounds__(256) k_gmg_dense_apply(GLevelDev L, const float *__restrict__ Ainv, const float *__restrict__ b,
                                                          float *__restrict__ x, const CGState *__restrict__ st) {
    const int done = ld_done(st), n = ld_early(L.nrows);
    if (done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        float acc = 0.0f;
        for (int j = lane; j < n; j += 32) acc += Ainv[i * n + j] * b[L.rows[j]];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) x[L.rows[i]] = acc;
    }
}

