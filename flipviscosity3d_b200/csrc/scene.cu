// Scene construction on the device: triangle mesh -> nodal signed distance field, boundary union, liquid seeding.
//
// Reference behaviour being reproduced (relative to /root/reference):
//   MeshLevelSet::calculateSignedDistanceField   src/meshlevelset.cpp:138-347 (distance :350-451, crossing parity :395-473)
//   FluidSimulation::addBoundary / resetBoundary src/fluidsimulation.cpp:45-62, domain box :198-239
//   FluidSimulation::addLiquid                   src/fluidsimulation.cpp:64-97 (8 jittered candidates per cell, libc rand())
//
// The reference computes exact point-triangle distances only inside a 3-cell box around every triangle and spreads the
// closest triangle to the rest of the grid with one sequential breadth-first pass (order dependent, approximate).  Here
// EVERY node gets the exact distance to the closest triangle (one thread per node, the triangles streamed through shared
// memory): identical, bit for bit, wherever the true distance is <= 3 cells (there the closest triangle's box contains the
// node and the reference's min over boxes is the true min; ties go to the lowest triangle index on both sides), and exact
// instead of approximate farther out.  Signs come from the same x-ray crossing parity in double precision.  Nothing in
// the substep reads the far field except through its sign (face fractions, cell-centre sign tests, seeding test), and the
// push-out of a particle that has entered the solid only ever happens within a cell or two of the surface, so a scene
// built here evolves exactly like one built by the host layer (tests/test_gpu_scene.py).
//
// Seeding keeps the reference's particle ORDER and its random sequence: glibc's rand() (TYPE_3 additive feedback
// generator, unseeded = srand(1)) is restated on the host as a tight loop (FlipRand, checked against libc in the tests),
// streamed to the device in chunks, and each chunk is compacted in (cell, candidate) order.
#include "sim.h"
#include "scan.h"
#include "../../include/flip_b200.h"
#include <cmath>
#include <cstring>
#include <vector>

// ---- glibc random(): r[i] = r[i-3] + r[i-31], output r[i] >> 1 (stdlib/random_r.c, TYPE_3) ------------------------
struct FlipRand {
    int32_t r[31];
    int f, b;
    void seed(unsigned int s) {
        if (s == 0) s = 1;
        r[0] = (int32_t)s;
        for (int i = 1; i < 31; i++) {
            long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            long word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; i++) next();
    }
    inline int32_t next() {
        uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
        r[f] = (int32_t)v;
        if (++f >= 31) f = 0;
        if (++b >= 31) b = 0;
        return (int32_t)(v >> 1);
    }
};
static FlipRand g_rand;
static bool g_rand_seeded = false;
static FlipRand &the_rand() {
    if (!g_rand_seeded) { g_rand.seed(1); g_rand_seeded = true; }
    return g_rand;
}

// ---- exact point-triangle distance, the reference's float expressions (src/meshlevelset.cpp:350-391, 438-451) ------
struct V3 { float x, y, z; };
FLIP_D V3 vsub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
FLIP_D V3 vadd(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
FLIP_D V3 vmul(float s, V3 v) { return V3{v.x * s, v.y * s, v.z * s}; }
FLIP_D float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FLIP_D float vlen2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
FLIP_D float vlen(V3 v) { return sqrtf(vlen2(v)); }

FLIP_D float segment_distance(V3 x0, V3 x1, V3 x2) {
    V3 e = vsub(x2, x1);
    double m2 = (double)vlen2(e);
    float s = (float)((double)vdot(vsub(x2, x0), e) / m2);
    if (s < 0) s = 0;
    else if (s > 1) s = 1;
    return vlen(vsub(x0, vadd(vmul(s, x1), vmul(1 - s, x2))));
}

FLIP_D float triangle_distance(V3 x0, V3 x1, V3 x2, V3 x3) {
    V3 x13 = vsub(x1, x3), x23 = vsub(x2, x3), x03 = vsub(x0, x3);
    float m13 = vlen2(x13), m23 = vlen2(x23);
    float d = vdot(x13, x23);
    float invdet = 1.0f / fmaxf(m13 * m23 - d * d, 1e-30f);
    float a = vdot(x13, x03), b = vdot(x23, x03);
    float w23 = invdet * (m23 * a - d * b);
    float w31 = invdet * (m13 * b - d * a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) return vlen(vsub(x0, vadd(vadd(vmul(w23, x1), vmul(w31, x2)), vmul(w12, x3))));
    if (w23 > 0) return fminf(segment_distance(x0, x1, x2), segment_distance(x0, x1, x3));
    if (w31 > 0) return fminf(segment_distance(x0, x1, x2), segment_distance(x0, x2, x3));
    return fminf(segment_distance(x0, x1, x3), segment_distance(x0, x2, x3));
}

// one thread per node; triangles staged through shared memory 128 at a time.  out = flat (W, H, D) array, x fastest.
#define SDF_TILE 128
__global__ void __launch_bounds__(256) k_sdf_distance(int W, int H, int D, double dx, const float *__restrict__ tri9, int nt,
                                                      float *__restrict__ out) {
    __shared__ float tv[SDF_TILE * 9];
    const long long n = (long long)W * H * D;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n;
    int i = 0, j = 0, k = 0;
    if (live) { i = (int)(t % W); j = (int)((t / W) % H); k = (int)(t / ((long long)W * H)); }
    const V3 gp = V3{(float)(i * dx), (float)(j * dx), (float)(k * dx)};
    float best = (float)((W + H + D) * dx);
    for (int base = 0; base < nt; base += SDF_TILE) {
        const int cnt = min(SDF_TILE, nt - base);
        __syncthreads();
        for (int q = threadIdx.x; q < cnt * 9; q += blockDim.x) tv[q] = tri9[(size_t)base * 9 + q];
        __syncthreads();
        if (live) {
            for (int c = 0; c < cnt; c++) {
                const float *p = tv + c * 9;
                float dist = triangle_distance(gp, V3{p[0], p[1], p[2]}, V3{p[3], p[4], p[5]}, V3{p[6], p[7], p[8]});
                if (dist < best) best = dist;
            }
        }
    }
    if (live) out[t] = best;
}

// x-ray crossing counts (src/meshlevelset.cpp:249-268, 395-473): one warp per triangle over its (j, k) box
FLIP_D int sdf_orientation(double x1, double y1, double x2, double y2, double *area2) {
    *area2 = y1 * x2 - x1 * y2;
    if (*area2 > 0) return 1;
    if (*area2 < 0) return -1;
    if (y2 > y1) return 1;
    if (y2 < y1) return -1;
    if (x1 > x2) return 1;
    if (x1 < x2) return -1;
    return 0;
}
FLIP_D bool sdf_barycentric(double x0, double y0, double x1, double y1, double x2, double y2, double x3, double y3, double *a,
                            double *b, double *c) {
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    double oa, ob, oc;
    int sa = sdf_orientation(x2, y2, x3, y3, &oa);
    if (sa == 0) return false;
    if (sdf_orientation(x3, y3, x1, y1, &ob) != sa) return false;
    if (sdf_orientation(x1, y1, x2, y2, &oc) != sa) return false;
    double inv = 1.0 / (oa + ob + oc);
    *a = oa * inv; *b = ob * inv; *c = oc * inv;
    return true;
}
FLIP_D int sdf_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256) k_sdf_crossings(int W, int H, int D, double dx, const float *__restrict__ tri9, int nt,
                                                       int *__restrict__ crossings) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nt) return;
    const float *p = tri9 + (size_t)warp * 9;
    const double invdx = 1.0 / dx;
    const double fip = (double)p[0] * invdx, fjp = (double)p[1] * invdx, fkp = (double)p[2] * invdx;
    const double fiq = (double)p[3] * invdx, fjq = (double)p[4] * invdx, fkq = (double)p[5] * invdx;
    const double fir = (double)p[6] * invdx, fjr = (double)p[7] * invdx, fkr = (double)p[8] * invdx;
    const int j0 = sdf_clampi((int)ceil(fmin(fjp, fmin(fjq, fjr))), 0, H - 1), j1 = sdf_clampi((int)floor(fmax(fjp, fmax(fjq, fjr))), 0, H - 1);
    const int k0 = sdf_clampi((int)ceil(fmin(fkp, fmin(fkq, fkr))), 0, D - 1), k1 = sdf_clampi((int)floor(fmax(fkp, fmax(fkq, fkr))), 0, D - 1);
    const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
    if (nj <= 0 || nk <= 0) return;
    for (long long q = lane; q < (long long)nj * nk; q += 32) {
        const int j = j0 + (int)(q % nj), k = k0 + (int)(q / nj);
        double a, b, c;
        if (!sdf_barycentric(j, k, fjp, fkp, fjq, fkq, fjr, fkr, &a, &b, &c)) continue;
        const double fi = a * fip + b * fiq + c * fir;
        const int cell = (int)ceil(fi);
        if (cell < 0) atomicAdd(&crossings[(size_t)W * (j + (size_t)H * k)], 1);
        else if (cell < W) atomicAdd(&crossings[cell + (size_t)W * (j + (size_t)H * k)], 1);
    }
}

// parity along +x: one thread per (j, k) row; optional negation of the whole field (inverted boundaries)
__global__ void __launch_bounds__(256) k_sdf_sign(int W, int H, int D, const int *__restrict__ crossings, float *__restrict__ phi, int negate) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= H * D) return;
    int total = 0;
    for (int i = 0; i < W; i++) {
        const size_t id = i + (size_t)W * row;
        total += crossings[id];
        float v = phi[id];
        if (total % 2 == 1) v = -v;
        phi[id] = negate ? -v : v;
    }
}

// solid SDF (padded layout) <- sdf (flat nodal array): replace, or union by min (MeshLevelSet::calculateUnion)
__global__ void __launch_bounds__(256) k_sdf_merge(Grid g, const float *__restrict__ sdf, float *__restrict__ phi_sol, int replace) {
    int i, j, k;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!unflatten(t, g.ni + 1, g.nj + 1, g.nk + 1, i, j, k)) return;
    const int id = gidx(g, i, j, k);
    const float v = sdf[t];
    if (replace || v < phi_sol[id]) phi_sol[id] = v;
}

// Interpolation::trilinearInterpolate on a FLAT nodal array (same expressions as grid.h trilinear_grid)
FLIP_D float sdf_trilinear_flat(const float *__restrict__ f, int w, int h, int d, double dxd, double invdx, float px, float py, float pz) {
    int gi = pos_to_index(px, invdx), gj = pos_to_index(py, invdx), gk = pos_to_index(pz, invdx);
    float gx = index_to_pos(gi, dxd), gy = index_to_pos(gj, dxd), gz = index_to_pos(gk, dxd);
    double x = (double)(px - gx) * invdx, y = (double)(py - gy) * invdx, z = (double)(pz - gz) * invdx;
    double p[8];
#define SDF_CORNER(n, a, b, c)                                                                            \
    {                                                                                                     \
        int ii = gi + a, jj = gj + b, kk = gk + c;                                                        \
        p[n] = (ii >= 0 && jj >= 0 && kk >= 0 && ii < w && jj < h && kk < d)                              \
                   ? (double)f[(size_t)ii + (size_t)w * ((size_t)jj + (size_t)h * (size_t)kk)] : 0.0;     \
    }
    SDF_CORNER(0, 0, 0, 0) SDF_CORNER(1, 1, 0, 0) SDF_CORNER(2, 0, 1, 0) SDF_CORNER(3, 0, 0, 1)
    SDF_CORNER(4, 1, 0, 1) SDF_CORNER(5, 0, 1, 1) SDF_CORNER(6, 1, 1, 0) SDF_CORNER(7, 1, 1, 1)
#undef SDF_CORNER
    return (float)(p[0] * (1 - x) * (1 - y) * (1 - z) + p[1] * x * (1 - y) * (1 - z) + p[2] * (1 - x) * y * (1 - z) +
                   p[3] * (1 - x) * (1 - y) * z + p[4] * x * (1 - y) * z + p[5] * (1 - x) * y * z +
                   p[6] * x * y * (1 - z) + p[7] * x * y * z);
}

// candidate positions of src/fluidsimulation.cpp:79-92: corner of the cell + three jitters in [0, dx]
FLIP_D void seed_candidate(const Grid &g, long long cell, const int *__restrict__ rnd, float &x, float &y, float &z) {
    const int i = (int)(cell % g.ni), j = (int)((cell / g.ni) % g.nj), k = (int)(cell / ((long long)g.ni * g.nj));
    const double hi = (double)g.dx;                       // _randomDouble(0, _dx): min + rand() / (RAND_MAX / (max - min))
    const double div = 2147483647.0 / hi;
    x = (float)((double)i * g.dxd) + (float)((double)rnd[0] / div);
    y = (float)((double)j * g.dxd) + (float)((double)rnd[1] / div);
    z = (float)((double)k * g.dxd) + (float)((double)rnd[2] / div);
}

// flags[c] = candidate c of the chunk is kept (inside the liquid mesh, outside the solid)
__global__ void __launch_bounds__(256) k_seed_flags(Grid g, long long cell0, long long ncand, const int *__restrict__ rnd,
                                                    const float *__restrict__ liq_flat, const float *__restrict__ phi_sol,
                                                    int *__restrict__ flags) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    float x, y, z;
    seed_candidate(g, cell0 + c / 8, rnd + 3 * c, x, y, z);
    int keep = 0;
    if ((double)sdf_trilinear_flat(liq_flat, g.ni + 1, g.nj + 1, g.nk + 1, g.dxd, g.invdx, x, y, z) < 0.0)
        keep = (float)trilinear_grid(g, phi_sol, g.ni + 1, g.nj + 1, g.nk + 1, x, y, z) >= 0 ? 1 : 0;
    flags[c] = keep;
}

__global__ void __launch_bounds__(256) k_seed_emit(Grid g, long long cell0, long long ncand, const int *__restrict__ rnd,
                                                   const int *__restrict__ flags, const int *__restrict__ offs, long long base,
                                                   float *px, float *py, float *pz, float *vx, float *vy, float *vz, unsigned *pid) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand || !flags[c]) return;
    float x, y, z;
    seed_candidate(g, cell0 + c / 8, rnd + 3 * c, x, y, z);
    const long long o = base + offs[c];
    px[o] = x; py[o] = y; pz[o] = z;
    vx[o] = 0.0f; vy[o] = 0.0f; vz[o] = 0.0f;
    pid[o] = (unsigned)o;
}

// ------------------------------------------------------------------------------------------------------------------
static void mesh_bounds_check(const Sim &s, const float *verts, int nv, const char *what) {
    // AABB(points) as the reference computes it (src/aabb.cpp:71-108, 205-211) must lie inside the domain
    // (FLUIDSIM_ASSERT at src/fluidsimulation.cpp:46-49, 65-68)
    double lo[3] = {verts[0], verts[1], verts[2]}, hi[3] = {verts[0], verts[1], verts[2]};
    for (int i = 0; i < nv; i++)
        for (int a = 0; a < 3; a++) { lo[a] = std::fmin((double)verts[3 * i + a], lo[a]); hi[a] = std::fmax((double)verts[3 * i + a], hi[a]); }
    const int n[3] = {s.g.ni, s.g.nj, s.g.nk};
    for (int a = 0; a < 3; a++) {
        float mn = (float)lo[a], mx = mn + (float)(hi[a] - lo[a] + 1e-9);
        double w = n[a] * s.g.dx;
        if (!(mn >= 0.0f && mx >= 0.0f && mn < 0.0f + w && mx < 0.0f + w)) throw FlipError(std::string(what) + ": the mesh must lie inside the simulation domain");
    }
}

// nodal SDF of a mesh into a flat device array (W*H*D floats, from the handle's scratch)
static float *mesh_sdf_device(Sim &s, const float *verts, int nv, const int *tris, int nt, bool negate) {
    const Grid &g = s.g;
    const int W = g.ni + 1, H = g.nj + 1, D = g.nk + 1;
    const long long n = (long long)W * H * D;
    std::vector<float> tri9((size_t)nt * 9);
    for (int t = 0; t < nt; t++)
        for (int c = 0; c < 3; c++) {
            int v = tris[3 * t + c];
            if (v < 0 || v >= nv) throw FlipError("mesh: triangle index out of range");
            for (int a = 0; a < 3; a++) tri9[(size_t)t * 9 + 3 * c + a] = verts[3 * v + a];
        }
    float *tri_dev = nullptr;
    CUDA_CHECK(cudaMalloc((void **)&tri_dev, tri9.size() * sizeof(float) + 16));
    CUDA_CHECK(cudaMemcpyAsync(tri_dev, tri9.data(), tri9.size() * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    float *sdf = s.vnode;                       // scratch outside the viscosity stage: 7 * total floats
    int *crossings = (int *)(s.vnode + (size_t)g.total);
    CUDA_CHECK(cudaMemsetAsync(crossings, 0, (size_t)n * sizeof(int), s.stream));
    FLIP_LAUNCH_SYNC(k_sdf_distance, cdiv(n, 256), 256, s.stream, W, H, D, g.dxd, (const float *)tri_dev, nt, sdf);
    FLIP_LAUNCH(k_sdf_crossings, cdiv(32LL * nt, 256), 256, s.stream, W, H, D, g.dxd, (const float *)tri_dev, nt, crossings);
    FLIP_LAUNCH(k_sdf_sign, cdiv((long long)H * D, 256), 256, s.stream, W, H, D, (const int *)crossings, sdf, negate ? 1 : 0);
    s.kernel_launches += 3;
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    cudaFree(tri_dev);
    return sdf;
}

static void merge_boundary(Sim &s, const float *sdf, bool replace) {
    const Grid &g = s.g;
    long long n = (long long)(g.ni + 1) * (g.nj + 1) * (g.nk + 1);
    FLIP_LAUNCH(k_sdf_merge, cdiv(n, 256), 256, s.stream, g, sdf, s.phi_sol, replace ? 1 : 0);
    s.kernel_launches++;
    KERNEL_CHECK();
    solid_precompute(s);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

// domain box inset by 3 dx + 1e-6 as a closed mesh, turned inside out (src/fluidsimulation.cpp:198-239)
void scene_reset_boundary(Sim &s) {
    const Grid &g = s.g;
    const float dx = g.dx;
    double v = -3 * dx - 1e-6;
    double half = 0.5 * v;
    float px = 0.0f - (float)half, py = 0.0f - (float)half, pz = 0.0f - (float)half;
    double w = (double)(g.ni * dx) + v, h = (double)(g.nj * dx) + v, d = (double)(g.nk * dx) + v;
    float x1 = px + (float)w, y1 = py + (float)h, z1 = pz + (float)d;
    const float vx[8][3] = {{px, py, pz}, {x1, py, pz}, {x1, py, z1}, {px, py, z1}, {px, y1, pz}, {x1, y1, pz}, {x1, y1, z1}, {px, y1, z1}};
    const int tr[12][3] = {{0, 1, 2}, {0, 2, 3}, {4, 7, 6}, {4, 6, 5}, {0, 3, 7}, {0, 7, 4}, {1, 5, 6}, {1, 6, 2}, {0, 4, 5}, {0, 5, 1}, {3, 2, 6}, {3, 6, 7}};
    float *sdf = mesh_sdf_device(s, &vx[0][0], 8, &tr[0][0], 12, true);
    merge_boundary(s, sdf, true);
}

void scene_add_boundary(Sim &s, const float *verts, int nv, const int *tris, int nt, bool inverted) {
    if (nv <= 0 || nt <= 0) throw FlipError("addBoundary: the mesh is empty");
    mesh_bounds_check(s, verts, nv, "addBoundary");
    float *sdf = mesh_sdf_device(s, verts, nv, tris, nt, inverted);
    merge_boundary(s, sdf, false);
}

static void grow_particles(Sim &s, long long need) {
    if (need <= s.cap) return;
    long long cap = need + need / 4 + 1024;
    float *np_[2][6];
    unsigned *nid[2];
    int *ncell = nullptr;
    for (int b = 0; b < 2; b++) {
        for (int f = 0; f < 6; f++) heap_alloc(s, np_[b][f], (size_t)cap);
        heap_alloc(s, nid[b], (size_t)cap);
    }
    heap_alloc(s, ncell, 2 * (size_t)cap);
    const int c = s.cur;
    if (s.np > 0) {
        for (int f = 0; f < 6; f++) CUDA_CHECK(cudaMemcpyAsync(np_[c][f], s.p[c][f], (size_t)s.np * sizeof(float), cudaMemcpyDeviceToDevice, s.stream));
        CUDA_CHECK(cudaMemcpyAsync(nid[c], s.pid[c], (size_t)s.np * sizeof(unsigned), cudaMemcpyDeviceToDevice, s.stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    for (int b = 0; b < 2; b++) {
        for (int f = 0; f < 6; f++) { heap_free(s, s.p[b][f]); s.p[b][f] = np_[b][f]; }
        heap_free(s, s.pid[b]); s.pid[b] = nid[b];
    }
    heap_free(s, s.cell_of); s.cell_of = ncell;
    s.cap = cap;
}

long long scene_add_liquid(Sim &s, const float *verts, int nv, const int *tris, int nt) {
    if (nv <= 0 || nt <= 0) throw FlipError("addLiquid: the mesh is empty");
    mesh_bounds_check(s, verts, nv, "addLiquid");
    const Grid &g = s.g;
    // particle ids are positions in the caller's order; after a substep the device order is the binned one, so new
    // particles can only be appended to a set that has not been stepped since it was loaded
    if (s.np > 0 && s.binned) throw FlipError("addLiquid: particles were already binned by a substep; read them back and reload first");
    float *liq = mesh_sdf_device(s, verts, nv, tris, nt, false);
    const long long ncells = (long long)g.ni * g.nj * g.nk;
    const long long chunk_cells = 1 << 20;
    const long long chunk_cand = chunk_cells * 8;
    int *rnd_host = nullptr, *rnd_dev = nullptr, *flags = nullptr, *offs = nullptr, *tmp = nullptr;
    CUDA_CHECK(cudaMallocHost((void **)&rnd_host, (size_t)chunk_cand * 3 * sizeof(int)));
    CUDA_CHECK(cudaMalloc((void **)&rnd_dev, (size_t)chunk_cand * 3 * sizeof(int)));
    CUDA_CHECK(cudaMalloc((void **)&flags, (size_t)chunk_cand * sizeof(int)));
    CUDA_CHECK(cudaMalloc((void **)&offs, ((size_t)chunk_cand + 1) * sizeof(int)));
    CUDA_CHECK(cudaMalloc((void **)&tmp, ((size_t)chunk_cand / 2048 + 2) * sizeof(int)));
    FlipRand &R = the_rand();
    long long added = 0;
    try {
        for (long long cell0 = 0; cell0 < ncells; cell0 += chunk_cells) {
            const long long nc = std::min(chunk_cells, ncells - cell0), ncand = nc * 8;
            for (long long q = 0; q < ncand * 3; q++) rnd_host[q] = R.next();
            CUDA_CHECK(cudaMemcpyAsync(rnd_dev, rnd_host, (size_t)ncand * 3 * sizeof(int), cudaMemcpyHostToDevice, s.stream));
            FLIP_LAUNCH(k_seed_flags, cdiv(ncand, 256), 256, s.stream, g, cell0, ncand, (const int *)rnd_dev, (const float *)liq,
                        (const float *)s.phi_sol, flags);
            exclusive_scan(s, flags, offs, tmp, (int)ncand);
            int count = 0;
            CUDA_CHECK(cudaMemcpyAsync(&count, offs + ncand, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
            s.kernel_launches += 1;
            if (count == 0) continue;
            if (s.np + count >= ((long long)1 << 31)) throw FlipError("addLiquid: more than 2^31-1 particles");
            grow_particles(s, s.np + count);
            const int c = s.cur;
            FLIP_LAUNCH(k_seed_emit, cdiv(ncand, 256), 256, s.stream, g, cell0, ncand, (const int *)rnd_dev, (const int *)flags,
                        (const int *)offs, (long long)s.np, s.p[c][0], s.p[c][1], s.p[c][2], s.p[c][3], s.p[c][4], s.p[c][5], s.pid[c]);
            s.kernel_launches += 1;
            KERNEL_CHECK();
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
            s.np += count;
            added += count;
        }
    } catch (...) {
        cudaFreeHost(rnd_host); cudaFree(rnd_dev); cudaFree(flags); cudaFree(offs); cudaFree(tmp);
        throw;
    }
    cudaFreeHost(rnd_host); cudaFree(rnd_dev); cudaFree(flags); cudaFree(offs); cudaFree(tmp);
    s.binned = false;
    return added;
}

void scene_srand(unsigned int seed) { g_rand.seed(seed); g_rand_seeded = true; }
int scene_rand_next() { return (int)the_rand().next(); }

void scene_mesh_sdf(Sim &s, const float *verts, int nv, const int *tris, int nt, float *out_host) {
    const Grid &g = s.g;
    float *sdf = mesh_sdf_device(s, verts, nv, tris, nt, false);
    size_t n = (size_t)(g.ni + 1) * (g.nj + 1) * (g.nk + 1);
    CUDA_CHECK(cudaMemcpy(out_host, sdf, n * sizeof(float), cudaMemcpyDeviceToHost));
}
