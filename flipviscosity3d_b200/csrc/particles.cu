// Particle kernels: cell binning, liquid SDF (gather), P2G (gather), G2P + FLIP blend + RK2 +
// solid push-out + clamp.
//
// Reference behaviour being reproduced (all citations relative to /root/reference):
//   liquid SDF      src/particlelevelset.cpp:98-139
//   P2G             src/fluidsimulation.cpp:364-438 (+ masking :440-498)
//   G2P / advect    src/fluidsimulation.cpp:315-352, 535-546; src/macvelocityfield.cpp:455-578
//
// Design: particles are kept cell-binned (counting sort, deterministic order inside a cell), so
// both scatters of the reference (min-scatter for the SDF, Wyvill-weighted add-scatter for P2G)
// become GATHERS over contiguous particle runs: no atomics, bit-reproducible run to run, and a
// k-slab decomposition only needs one ghost layer of particles instead of a halo-add.
#include "cg.h"
#include "scan.h"

// ------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------
__global__ void k_cell_count(Grid g, const float *__restrict__ px, const float *__restrict__ py,
                             const float *__restrict__ pz, int *__restrict__ cell_of,
                             int *__restrict__ counts, long long np) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    // the cell a particle belongs to: Grid3d::positionToGridIndex (src/grid3d.h:60-65)
    int i = pos_to_index(px[t], g.invdx), j = pos_to_index(py[t], g.invdx), k = pos_to_index(pz[t], g.invdx);
    i = min(max(i, 0), g.ni - 1); j = min(max(j, 0), g.nj - 1); k = min(max(k, 0), g.nk - 1);
    int c = gidx(g, i, j, k);
    cell_of[t] = c;
    atomicAdd(&counts[c], 1);
}

__global__ void k_cell_scatter(const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                               int *__restrict__ cursor, int *__restrict__ order, long long np) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    int c = cell_of[t];
    int slot = cell_start[c] + atomicAdd(&cursor[c], 1);
    order[slot] = (int)t;
}

// Final placement: inside a cell particles are ordered by their persistent id, which makes the
// binned order (and every float sum over it) independent of the atomics above.
struct PArrays { float *a[6]; unsigned *id; };
__global__ void k_cell_place(const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                             const int *__restrict__ order, PArrays src, PArrays dst, long long np) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    int me = order[t];
    int c = cell_of[me];
    int b = cell_start[c], e = cell_start[c + 1];
    unsigned myid = src.id[me];
    int rank = 0;
    for (int q = b; q < e; q++) rank += (src.id[order[q]] < myid) ? 1 : 0;
    int d = b + rank;
#pragma unroll
    for (int f = 0; f < 6; f++) dst.a[f][d] = src.a[f][me];
    dst.id[d] = myid;
}

void bin_particles(Sim &s) {
    if (s.binned) return;
    const Grid &g = s.g;
    int n = g.total;
    CUDA_CHECK(cudaMemsetAsync(s.cell_cursor, 0, sizeof(int) * (size_t)n, s.stream));
    if (s.np > 0) {
        int c = s.cur;
        FLIP_LAUNCH(k_cell_count, cdiv(s.np, 256), 256, s.stream, g, (const float *)s.p[c][0], (const float *)s.p[c][1],
                    (const float *)s.p[c][2], s.cell_of, s.cell_cursor, s.np);
        s.kernel_launches++;
    }
    exclusive_scan(s, s.cell_cursor, s.cell_start, s.scan_tmp, n);
    if (s.np > 0) {
        int c = s.cur;
        int *order = s.cell_of + s.cap;  // second half of the cell_of allocation
        CUDA_CHECK(cudaMemsetAsync(s.cell_cursor, 0, sizeof(int) * (size_t)n, s.stream));
        FLIP_LAUNCH(k_cell_scatter, cdiv(s.np, 256), 256, s.stream, (const int *)s.cell_of, (const int *)s.cell_start,
                    s.cell_cursor, order, s.np);
        PArrays src, dst;
        for (int f = 0; f < 6; f++) { src.a[f] = s.p[c][f]; dst.a[f] = s.p[c ^ 1][f]; }
        src.id = s.pid[c]; dst.id = s.pid[c ^ 1];
        FLIP_LAUNCH(k_cell_place, cdiv(s.np, 256), 256, s.stream, (const int *)s.cell_of, (const int *)s.cell_start,
                    (const int *)order, src, dst, s.np);
        s.kernel_launches += 2;
        s.cur = c ^ 1;
    }
    KERNEL_CHECK();
    s.binned = true;
    s.world_epoch++;   // the cell occupancy the block list is built from is new
}

// ------------------------------------------------------------------------------------------
// liquid SDF: phi(cell) = min over particles in the 3x3x3 cells around it of |centre - p| - r,
// then phi = -dx/2 where phi < dx/2 inside the solid.   (src/particlelevelset.cpp:98-139)
// min is exact and order independent, so this gather is bit-identical to the reference scatter.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CG_THREADS) k_liquid_sdf(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                             const float *__restrict__ px, const float *__restrict__ py,
                             const float *__restrict__ pz, const int *__restrict__ cell_start,
                             const float *__restrict__ sol_center, float *__restrict__ phi, float radius) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    if (i >= g.ni || j >= g.nj || k >= g.nk) continue;
    float cx = index_to_center(i, g.dxd), cy = index_to_center(j, g.dxd), cz = index_to_center(k, g.dxd);
    float best = 3.0f * g.dx;  // ParticleLevelSet::_getMaxDistance (src/particlelevelset.cpp:94-96)
    int i0 = max(i - 1, 0), i1 = min(i + 1, g.ni - 1);
    for (int kk = max(k - 1, 0); kk <= min(k + 1, g.nk - 1); kk++) {
        for (int jj = max(j - 1, 0); jj <= min(j + 1, g.nj - 1); jj++) {
            int b = cell_start[gidx(g, i0, jj, kk)], e = cell_start[gidx(g, i1, jj, kk) + 1];
            for (int q = b; q < e; q++) {
                float vx = cx - px[q], vy = cy - py[q], vz = cz - pz[q];
                float dist = sqrtf(vx * vx + vy * vy + vz * vz) - radius;
                if (dist < best) best = dist;
            }
        }
    }
    int id = gidx(g, i, j, k);
    // _extrapolateSignedDistanceIntoSolids (src/particlelevelset.cpp:127-139)
    if ((double)best < 0.5 * g.dxd && sol_center[id] < 0.0f) best = -0.5f * g.dx;
    phi[id] = best;
  }
}

void stage_update_liquid_sdf(Sim &s) {
    bin_particles(s);
    const Grid &g = s.g;
    int c = s.cur;
    grid_list_ensure(s);   // blocks within one block of a particle (or of last substep's liquid): elsewhere phi stays 3 dx
    FLIP_LAUNCH(k_liquid_sdf, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.p[c][0], (const float *)s.p[c][1], (const float *)s.p[c][2], (const int *)s.cell_start,
                (const float *)s.sol_center, s.phi_liq, s.particle_radius);
    s.kernel_launches++;
    KERNEL_CHECK();
    s.world_epoch++;       // the liquid SDF changed
}

// ------------------------------------------------------------------------------------------
// P2G: Wyvill-kernel (radius dx) normalised transfer of one velocity component to its faces,
// then the masking of _advectVelocityField{U,V,W}: a face keeps its value only if it borders a
// fluid cell and received weight >= 1e-9.    (src/fluidsimulation.cpp:364-498)
// Gather form: a face sums over the particles of the 2x3x3 (U), 3x2x3 (V), 3x3x2 (W) cells that
// can be within dx of it, with the reference's own acceptance test distsq < rsq.
// ------------------------------------------------------------------------------------------
struct Wyvill { float rsq, c1, c2, c3; };

template <int DIR>
FLIP_D void p2g_face(const Grid &g, const float *__restrict__ px, const float *__restrict__ py,
                     const float *__restrict__ pz, const float *__restrict__ pv,
                     const int *__restrict__ cell_start, const float *__restrict__ phi, const Wyvill wy,
                     int i, int j, int k, float *__restrict__ out, unsigned char *__restrict__ valid) {
    int id = gidx(g, i, j, k);
    float val = 0.0f;
    unsigned char ok = 0;
    if (face_borders_fluid(g, phi, DIR, i, j, k)) {
        // offsets of src/fluidsimulation.cpp:369-379
        float ox = DIR == 0 ? 0.0f : g.hdx, oy = DIR == 1 ? 0.0f : g.hdx, oz = DIR == 2 ? 0.0f : g.hdx;
        float gx = index_to_pos(i, g.dxd), gy = index_to_pos(j, g.dxd), gz = index_to_pos(k, g.dxd);
        // candidate cells: along DIR the face sits on the cell boundary (cells c-1, c); across it
        // the shifted sample sits at the cell centre (cells c-1 .. c+1)
        int i0 = max(i - 1, 0), i1 = min(DIR == 0 ? i : i + 1, g.ni - 1);
        int j0 = max(j - 1, 0), j1 = min(DIR == 1 ? j : j + 1, g.nj - 1);
        int k0 = max(k - 1, 0), k1 = min(DIR == 2 ? k : k + 1, g.nk - 1);
        float sum = 0.0f, wsum = 0.0f;
        for (int kk = k0; kk <= k1; kk++) {
            for (int jj = j0; jj <= j1; jj++) {
                int b = cell_start[gidx(g, i0, jj, kk)], e = cell_start[gidx(g, i1, jj, kk) + 1];
                for (int q = b; q < e; q++) {
                    float vx = gx - (px[q] - ox), vy = gy - (py[q] - oy), vz = gz - (pz[q] - oz);
                    float d2 = vx * vx + vy * vy + vz * vz;
                    if (d2 < wy.rsq) {
                        float w = 1.0f - wy.c1 * d2 * d2 * d2 + wy.c2 * d2 * d2 - wy.c3 * d2;
                        sum += w * pv[q];
                        wsum += w;
                    }
                }
            }
        }
        if (!((double)wsum < 1e-9)) {
            val = sum / wsum;
            ok = 1;
        }
    }
    out[id] = val;
    valid[id] = ok;
}

__global__ void __launch_bounds__(CG_THREADS) k_p2g(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                             const float *__restrict__ px, const float *__restrict__ py,
                                             const float *__restrict__ pz, const float *__restrict__ vx,
                                             const float *__restrict__ vy, const float *__restrict__ vz,
                                             const int *__restrict__ cell_start, const float *__restrict__ phi,
                                             Wyvill wy, float *__restrict__ vel, unsigned char *__restrict__ valid) {
    FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
        size_t T = (size_t)g.total;
        if (j < g.nj && k < g.nk) p2g_face<0>(g, px, py, pz, vx, cell_start, phi, wy, i, j, k, vel, valid);
        if (i < g.ni && k < g.nk) p2g_face<1>(g, px, py, pz, vy, cell_start, phi, wy, i, j, k, vel + T, valid + T);
        if (i < g.ni && j < g.nj) p2g_face<2>(g, px, py, pz, vz, cell_start, phi, wy, i, j, k, vel + 2 * T, valid + 2 * T);
    }
}

void stage_advect_velocity_field(Sim &s) {
    bin_particles(s);
    const Grid &g = s.g;
    int c = s.cur;
    // Wyvill coefficients exactly as src/fluidsimulation.cpp:383-388 (float arithmetic, same order)
    Wyvill wy;
    float r = g.dx;
    wy.rsq = r * r;
    wy.c1 = (4.0f / 9.0f) * (1.0f / (r * r * r * r * r * r));
    wy.c2 = (17.0f / 9.0f) * (1.0f / (r * r * r * r));
    wy.c3 = (22.0f / 9.0f) * (1.0f / (r * r));
    grid_list_ensure(s);
    FLIP_LAUNCH(k_p2g, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.p[c][0], (const float *)s.p[c][1],
                (const float *)s.p[c][2], (const float *)s.p[c][3], (const float *)s.p[c][4], (const float *)s.p[c][5],
                (const int *)s.cell_start, (const float *)s.phi_liq, wy, s.vel, s.valid);
    s.kernel_launches++;
    KERNEL_CHECK();
    extrapolate_velocity(s);
    // _savedVelocityField = _MACVelocity (src/fluidsimulation.cpp:518)
    CUDA_CHECK(cudaMemcpyAsync(s.saved, s.vel, sizeof(float) * 3 * (size_t)g.total, cudaMemcpyDeviceToDevice, s.stream));
}

// ------------------------------------------------------------------------------------------
// G2P: trilinear MAC sampling in double, corners outside the component grid read 0
// (src/macvelocityfield.cpp:455-578, src/interpolation.cpp:57-66)
// ------------------------------------------------------------------------------------------
template <int DIR>
FLIP_D double mac_sample(const Grid &g, const float *__restrict__ f, double x, double y, double z) {
    // isPositionInGrid (src/grid3d.h:130-132)
    if (!(x >= 0 && y >= 0 && z >= 0 && x < g.dxd * g.ni && y < g.dxd * g.nj && z < g.dxd * g.nk)) return 0.0;
    if (DIR != 0) x -= 0.5 * g.dxd;
    if (DIR != 1) y -= 0.5 * g.dxd;
    if (DIR != 2) z -= 0.5 * g.dxd;
    int i = (int)floor(x * g.invdx), j = (int)floor(y * g.invdx), k = (int)floor(z * g.invdx);
    double gx = (double)i * g.dxd, gy = (double)j * g.dxd, gz = (double)k * g.dxd;
    double inv = 1 / g.dxd;
    double ix = (x - gx) * inv, iy = (y - gy) * inv, iz = (z - gz) * inv;
    const int w = g.ni + (DIR == 0), h = g.nj + (DIR == 1), d = g.nk + (DIR == 2);
    double p[8];
#define FLIP_CORNER(n, a, b, c)                                                            \
    {                                                                                      \
        int ii = i + a, jj = j + b, kk = k + c;                                            \
        p[n] = (ii >= 0 && jj >= 0 && kk >= 0 && ii < w && jj < h && kk < d)               \
                   ? (double)f[gidx(g, ii, jj, kk)] : 0.0;                                 \
    }
    FLIP_CORNER(0, 0, 0, 0) FLIP_CORNER(1, 1, 0, 0) FLIP_CORNER(2, 0, 1, 0) FLIP_CORNER(3, 0, 0, 1)
    FLIP_CORNER(4, 1, 0, 1) FLIP_CORNER(5, 0, 1, 1) FLIP_CORNER(6, 1, 1, 0) FLIP_CORNER(7, 1, 1, 1)
#undef FLIP_CORNER
    return p[0] * (1 - ix) * (1 - iy) * (1 - iz) + p[1] * ix * (1 - iy) * (1 - iz) + p[2] * (1 - ix) * iy * (1 - iz) +
           p[3] * (1 - ix) * (1 - iy) * iz + p[4] * ix * (1 - iy) * iz + p[5] * (1 - ix) * iy * iz +
           p[6] * ix * iy * (1 - iz) + p[7] * ix * iy * iz;
}

FLIP_D void mac_velocity(const Grid &g, const float *__restrict__ vel, float px, float py, float pz,
                         float &ox, float &oy, float &oz) {
    double x = px, y = py, z = pz;
    size_t T = (size_t)g.total;
    ox = (float)mac_sample<0>(g, vel, x, y, z);
    oy = (float)mac_sample<1>(g, vel + T, x, y, z);
    oz = (float)mac_sample<2>(g, vel + 2 * T, x, y, z);
}

struct AdvectBox {  // AABB(0,0,0,ni*dx,..).expand(-2dx-1e-4) and its clamp  (src/aabb.cpp:112-129, 213-233)
    float minp[3];
    double maxp_excl[3];
    float maxclamp[3];
};

__global__ void __launch_bounds__(256) k_advect_particles(Grid g, float *__restrict__ px, float *__restrict__ py,
                                                          float *__restrict__ pz, float *__restrict__ vx,
                                                          float *__restrict__ vy, float *__restrict__ vz,
                                                          const float *__restrict__ vel, const float *__restrict__ saved,
                                                          const float *__restrict__ phi_sol, float dt, float pic,
                                                          AdvectBox box, long long np) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    float x = px[t], y = py[t], z = pz[t];
    // _updateFluidParticleVelocities (src/fluidsimulation.cpp:341-352)
    float nx, ny, nz, sx, sy, sz;
    mac_velocity(g, vel, x, y, z, nx, ny, nz);
    mac_velocity(g, saved, x, y, z, sx, sy, sz);
    float ux = vx[t], uy = vy[t], uz = vz[t];
    float fx = ux + nx - sx, fy = uy + ny - sy, fz = uz + nz - sz;
    float flip = 1.0f - pic;
    vx[t] = pic * nx + flip * fx;
    vy[t] = pic * ny + flip * fy;
    vz[t] = pic * nz + flip * fz;
    // _traceRK2 (src/fluidsimulation.cpp:535-541); the first evaluation is (nx,ny,nz) again
    float h = 0.5f * dt;
    float mx, my, mz;
    mac_velocity(g, vel, x + h * nx, y + h * ny, z + h * nz, mx, my, mz);
    x += dt * mx; y += dt * my; z += dt * mz;
    // push out of the solid along the gradient (src/fluidsimulation.cpp:325-333)
    float phi = (float)trilinear_grid(g, phi_sol, g.ni + 1, g.nj + 1, g.nk + 1, x, y, z);
    if (phi < 0) {
        // Interpolation::trilinearInterpolateGradient (src/interpolation.cpp:122-184)
        int gi = pos_to_index(x, g.invdx), gj = pos_to_index(y, g.invdx), gk = pos_to_index(z, g.invdx);
        float gx = index_to_pos(gi, g.dxd), gy = index_to_pos(gj, g.dxd), gz = index_to_pos(gk, g.dxd);
        double ix = (double)(x - gx) * g.invdx, iy = (double)(y - gy) * g.invdx, iz = (double)(z - gz) * g.invdx;
        float v[2][2][2];
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++)
                for (int c = 0; c < 2; c++) {
                    int ii = gi + a, jj = gj + b, kk = gk + c;
                    v[a][b][c] = (ii >= 0 && jj >= 0 && kk >= 0 && ii <= g.ni && jj <= g.nj && kk <= g.nk)
                                     ? phi_sol[gidx(g, ii, jj, kk)] : 0.0f;
                }
#define FLIP_BILERP(v00, v10, v01, v11, s, t2) \
    ((1 - (t2)) * ((1 - (s)) * (double)(v00) + (s) * (double)(v10)) + (t2) * ((1 - (s)) * (double)(v01) + (s) * (double)(v11)))
        float dxx = (float)FLIP_BILERP(v[1][0][0] - v[0][0][0], v[1][1][0] - v[0][1][0], v[1][0][1] - v[0][0][1],
                                       v[1][1][1] - v[0][1][1], iy, iz);
        float dyy = (float)FLIP_BILERP(v[0][1][0] - v[0][0][0], v[1][1][0] - v[1][0][0], v[0][1][1] - v[0][0][1],
                                       v[1][1][1] - v[1][0][1], ix, iz);
        float dzz = (float)FLIP_BILERP(v[0][0][1] - v[0][0][0], v[1][0][1] - v[1][0][0], v[0][1][1] - v[0][1][0],
                                       v[1][1][1] - v[1][1][0], ix, iy);
#undef FLIP_BILERP
        float l2 = dxx * dxx + dyy * dyy + dzz * dzz;
        if (l2 > 0) {
            float inv = 1.0f / sqrtf(l2);  // vmath::normalize: v / len with operator/ = v * (1.0f/len)
            dxx *= inv; dyy *= inv; dzz *= inv;
        }
        x -= phi * dxx; y -= phi * dyy; z -= phi * dzz;
    }
    // clamp into the inset box (src/fluidsimulation.cpp:335-337)
    bool inside = x >= box.minp[0] && y >= box.minp[1] && z >= box.minp[2] && (double)x < box.maxp_excl[0] &&
                  (double)y < box.maxp_excl[1] && (double)z < box.maxp_excl[2];
    if (!inside) {
        x = fminf(fmaxf(x, box.minp[0]), box.maxclamp[0]);
        y = fminf(fmaxf(y, box.minp[1]), box.maxclamp[1]);
        z = fminf(fmaxf(z, box.minp[2]), box.maxclamp[2]);
    }
    px[t] = x; py[t] = y; pz[t] = z;
}

void stage_advect_particles(Sim &s, float dt) {
    if (s.np == 0) return;
    const Grid &g = s.g;
    AdvectBox box;
    int n[3] = {g.ni, g.nj, g.nk};
    for (int a = 0; a < 3; a++) {
        // AABB boundary(0,0,0, n*dx ...) with float n*dx widened to double; expand(-2*dx - 1e-4)
        double width = (double)((float)n[a] * g.dx);
        double v = (double)(-2 * g.dx) - 1e-4;
        double hh = 0.5 * v;
        float pos = 0.0f - (float)hh;
        width += v;
        box.minp[a] = pos;
        box.maxp_excl[a] = (double)pos + width;
        box.maxclamp[a] = (pos + (float)width) - (float)1e-6;
    }
    int c = s.cur;
    FLIP_LAUNCH(k_advect_particles, cdiv(s.np, 256), 256, s.stream, g, s.p[c][0], s.p[c][1], s.p[c][2], s.p[c][3],
                s.p[c][4], s.p[c][5], (const float *)s.vel, (const float *)s.saved, (const float *)s.phi_sol, dt,
                s.pic_ratio, box, s.np);
    s.kernel_launches++;
    KERNEL_CHECK();
    s.binned = false;
}
