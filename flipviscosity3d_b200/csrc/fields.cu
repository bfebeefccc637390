// Grid-only stages of the substep: static solid-derived fields, body force, layered velocity
// extrapolation, pressure gradient update, constraint, CFL max-reduce.
//
// Reference behaviour being reproduced (relative to /root/reference):
//   weights            src/fluidsimulation.cpp:549-582, src/meshlevelset.cpp:92-126
//   body force         src/fluidsimulation.cpp:271-312
//   extrapolation      src/macvelocityfield.cpp:580-694 (7 layers, src/fluidsimulation.cpp:690-694)
//   apply pressure     src/fluidsimulation.cpp:598-688
//   constrain          src/fluidsimulation.cpp:696-729
//   CFL                src/fluidsimulation.cpp:241-269
#include "cg.h"
#include "levelset_math.h"

// ------------------------------------------------------------------------------------------
// static fields derived from the solid SDF (computed once per boundary change; the reference
// recomputes the weights every substep although the solid never moves)
// ------------------------------------------------------------------------------------------
__global__ void k_solid_center(Grid g, const float *__restrict__ ps, float *__restrict__ sc) {
    int i, j, k;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, g.ni, g.nj, g.nk, i, j, k)) return;
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    // MeshLevelSet::getDistanceAtCellCenter (src/meshlevelset.cpp:66-76), same summation order
    sc[id] = 0.125f * (ps[id] + ps[id + 1] + ps[id + sy] + ps[id + 1 + sy] + ps[id + sz] + ps[id + 1 + sz] +
                       ps[id + sy + sz] + ps[id + 1 + sy + sz]);
}

__global__ void k_solid_faces(Grid g, const float *__restrict__ ps, const float *__restrict__ sc,
                              float *__restrict__ weight, unsigned char *__restrict__ fstate) {
    int i, j, k;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, g.ni + 1, g.nj + 1, g.nk + 1, i, j, k)) return;
    int id = gidx(g, i, j, k), sy = SY(g), sz = SZ(g);
    size_t T = (size_t)g.total;
    if (j < g.nj && k < g.nk) {  // U face
        float w = 1.0f - frac_inside4(ps[id], ps[id + sy], ps[id + sz], ps[id + sy + sz]);
        weight[id] = fmaxf(0.0f, fminf(w, 1.0f));
        // ViscositySolver::_computeFaceStateGrid (src/viscositysolver.cpp:80-123)
        bool edge = i == 0 || i == g.ni;
        fstate[id] = (edge || sc[id - 1] + sc[id] <= 0) ? 1 : 0;
    }
    if (i < g.ni && k < g.nk) {  // V face
        float w = 1.0f - frac_inside4(ps[id], ps[id + sz], ps[id + 1], ps[id + 1 + sz]);
        weight[T + id] = fmaxf(0.0f, fminf(w, 1.0f));
        bool edge = j == 0 || j == g.nj;
        fstate[T + id] = (edge || sc[id - sy] + sc[id] <= 0) ? 1 : 0;
    }
    if (i < g.ni && j < g.nj) {  // W face
        float w = 1.0f - frac_inside4(ps[id], ps[id + sy], ps[id + 1], ps[id + 1 + sy]);
        weight[2 * T + id] = fmaxf(0.0f, fminf(w, 1.0f));
        bool edge = k == 0 || k == g.nk;
        fstate[2 * T + id] = (edge || sc[id - sz] + sc[id] <= 0) ? 1 : 0;
    }
}

void solid_precompute(Sim &s) {
    const Grid &g = s.g;
    CUDA_CHECK(cudaMemsetAsync(s.sol_center, 0, sizeof(float) * (size_t)g.total, s.stream));
    CUDA_CHECK(cudaMemsetAsync(s.weight, 0, sizeof(float) * 3 * (size_t)g.total, s.stream));
    // faces outside the component range count as solid for the viscosity stencil
    CUDA_CHECK(cudaMemsetAsync(s.fstate, 1, 3 * (size_t)g.total, s.stream));
    long long nc = (long long)g.ni * g.nj * g.nk;
    FLIP_LAUNCH(k_solid_center, cdiv(nc, 256), 256, s.stream, g, (const float *)s.phi_sol, s.sol_center);
    long long nf = (long long)(g.ni + 1) * (g.nj + 1) * (g.nk + 1);
    FLIP_LAUNCH(k_solid_faces, cdiv(nf, 256), 256, s.stream, g, (const float *)s.phi_sol, (const float *)s.sol_center,
                s.weight, s.fstate);
    s.kernel_launches += 2;
    KERNEL_CHECK();
}

// ------------------------------------------------------------------------------------------
// near-liquid block list.  Everything a substep computes on the grid is non-default only within 8 cells of a liquid cell
// (valid faces border liquid cells, extrapolation reaches 7 layers, viscosity volumes 2 cells), i.e. inside the
// 26-neighbourhood of a block that holds a particle or a liquid cell.  The grid stages therefore run over
//     list = dilate1(blocks with a particle or phi < 0)  U  blocks listed during the previous substep
// (the second term rewrites the defaults where the liquid has left), ~10 % of the 256^3 bunny grid, instead of 17 M cells.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CG_THREADS) k_grid_flag0(Grid g, const float *__restrict__ phi, const int *__restrict__ cell_start,
                                                            int *__restrict__ flag0) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    BlockCell c = block_cell(g, blockIdx.x, threadIdx.x);
    if (c.inside && c.i < g.ni && c.j < g.nj && c.k < g.nk) {
        int id = gidx(g, c.i, c.j, c.k);
        if (phi[id] < 0.0f || (cell_start && cell_start[id + 1] > cell_start[id])) any = 1;   // benign same-value race
    }
    __syncthreads();
    if (threadIdx.x == 0) flag0[blockIdx.x] = any;
}

__global__ void __launch_bounds__(256) k_grid_dilate(Grid g, const int *__restrict__ flag0, const int *__restrict__ dirty,
                                                     int *__restrict__ dirty_next, int *__restrict__ flag, int all) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.nblocks) return;
    int bi = b % g.nbx, r = b / g.nbx;
    int bj = r % g.nby, bk = r / g.nby;
    int act = 0;
    for (int c = -1; c <= 1; c++)
        for (int bb = -1; bb <= 1; bb++)
            for (int a = -1; a <= 1; a++) {
                int x = bi + a, y = bj + bb, z = bk + c;
                if (x < 0 || y < 0 || z < 0 || x >= g.nbx || y >= g.nby || z >= g.nbz) continue;
                act |= flag0[x + g.nbx * (y + g.nby * z)];
            }
    if (act) dirty_next[b] = 1;
    flag[b] = (act || dirty[b] || all) ? 1 : 0;
}

void grid_list_ensure(Sim &s) {
    if (s.grid_list_epoch == s.world_epoch) return;
    const Grid &g = s.g;
    FLIP_LAUNCH_SYNC(k_grid_flag0, g.nblocks, CG_THREADS, s.stream, g, (const float *)s.phi_liq,
                     s.binned ? (const int *)s.cell_start : (const int *)nullptr, s.grid_flag0);
    FLIP_LAUNCH(k_grid_dilate, cdiv(g.nblocks, 256), 256, s.stream, g, (const int *)s.grid_flag0, (const int *)s.grid_dirty,
                s.grid_dirty_next, s.grid_flag, s.use_block_lists ? 0 : 1);
    FLIP_LAUNCH_SYNC(k_compact_blocks, 1, 1024, s.stream, (const int *)s.grid_flag, g.nblocks, s.grid_list, s.grid_count);
    s.kernel_launches += 3;
    KERNEL_CHECK();
    s.grid_list_epoch = s.world_epoch;
}

void grid_list_mark_all_dirty(Sim &s) {
    CUDA_CHECK(cudaMemsetAsync(s.grid_dirty, 1, sizeof(int) * (size_t)s.g.nblocks, s.stream));
    s.world_epoch++;
}

// After a substep in which every grid stage ran, the blocks that were only listed because they had been listed before
// hold defaults again: from now on "previously listed" means "listed during this substep".
void grid_list_end_substep(Sim &s, bool all_stages_ran) {
    if (!all_stages_ran) return;
    int *t = s.grid_dirty; s.grid_dirty = s.grid_dirty_next; s.grid_dirty_next = t;
    CUDA_CHECK(cudaMemsetAsync(s.grid_dirty_next, 0, sizeof(int) * (size_t)s.g.nblocks, s.stream));
    s.world_epoch++;   // the next list is built from the new "previously listed" set
}

// ------------------------------------------------------------------------------------------
// body force
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CG_THREADS) k_body_force(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                            const float *__restrict__ phi, float *__restrict__ vel, float ax, float ay, float az) {
    FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
        int id = gidx(g, i, j, k);
        size_t T = (size_t)g.total;
        if (j < g.nj && k < g.nk && face_borders_fluid(g, phi, 0, i, j, k)) vel[id] += ax;
        if (i < g.ni && k < g.nk && face_borders_fluid(g, phi, 1, i, j, k)) vel[T + id] += ay;
        if (i < g.ni && j < g.nj && face_borders_fluid(g, phi, 2, i, j, k)) vel[2 * T + id] += az;
    }
}

void stage_add_body_force(Sim &s, float dt) {
    const Grid &g = s.g;
    grid_list_ensure(s);
    // _gravity.x * dt in float (src/fluidsimulation.cpp:286)
    FLIP_LAUNCH(k_body_force, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.phi_liq, s.vel, s.gravity[0] * dt, s.gravity[1] * dt, s.gravity[2] * dt);
    s.kernel_launches++;
    KERNEL_CHECK();
}

// ------------------------------------------------------------------------------------------
// layered extrapolation.  The reference's sweep is order independent inside a layer (a cell found
// in layer L only averages neighbours that were KNOWN before layer L), so one Jacobi kernel per
// layer over all three components reproduces it exactly, including the float summation order
// -i,+i,-j,+j,-k,+k.  layer[] holds: 0 = valid input, L = became known in layer L, 254 = frozen
// border unknown (DONE), 255 = unknown.
// ------------------------------------------------------------------------------------------
#define LAYER_DONE 254
#define LAYER_UNKNOWN 255

// Block-shaped (one CTA per 8x8x8 block) so that the same pass also flags the blocks that hold a valid face:
// the layers only ever reach faces within `layers` <= 8 cells of a valid one, i.e. inside the 26-neighbourhood
// of a flagged block, and the layer kernels then run over that block list (~10 % of the 256^3 grid) instead of
// sweeping the whole grid 14 times per substep.
__global__ void __launch_bounds__(CG_THREADS) k_extrap_init(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                             const unsigned char *__restrict__ valid, unsigned char *__restrict__ layer,
                                                             int *__restrict__ flag) {
    __shared__ int any;
  for (int lb = blockIdx.x, nlb = *gcount; lb < nlb; lb += gridDim.x) {
    const int blk = glist[lb];
    __syncthreads();
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    BlockCell bc = block_cell(g, blk, threadIdx.x);
    bool has = false;
    if (bc.inside) {
        const int i = bc.i, j = bc.j, k = bc.k;
        int id = gidx(g, i, j, k);
        size_t T = (size_t)g.total;
        for (int c = 0; c < 3; c++) {
            int w = g.ni + (c == 0), h = g.nj + (c == 1), d = g.nk + (c == 2);
            if (i >= w || j >= h || k >= d) continue;
            bool border = i == 0 || j == 0 || k == 0 || i == w - 1 || j == h - 1 || k == d - 1;
            bool v = valid[c * T + id] != 0;
            has = has || v;
            layer[c * T + id] = v ? 0 : (border ? LAYER_DONE : LAYER_UNKNOWN);
        }
    }
    if (has) any = 1;  // benign same-value race
    __syncthreads();
    if (threadIdx.x == 0) flag[blk] = any;
  }
}

// flag2[b] = any flagged block in the 26-neighbourhood of b
__global__ void __launch_bounds__(256) k_extrap_dilate(Grid g, const int *__restrict__ flag, int *__restrict__ flag2) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.nblocks) return;
    int bi = b % g.nbx, r = b / g.nbx;
    int bj = r % g.nby, bk = r / g.nby;
    int any = 0;
    for (int c = -1; c <= 1; c++)
        for (int bb = -1; bb <= 1; bb++)
            for (int a = -1; a <= 1; a++) {
                int x = bi + a, y = bj + bb, z = bk + c;
                if (x < 0 || y < 0 || z < 0 || x >= g.nbx || y >= g.nby || z >= g.nbz) continue;
                any |= flag[x + g.nbx * (y + g.nby * z)];
            }
    flag2[b] = any;
}

__global__ void __launch_bounds__(CG_THREADS) k_extrap_layer(Grid g, const int *__restrict__ list, const int *__restrict__ count,
                                                              float *__restrict__ vel, unsigned char *__restrict__ layer, int L) {
    const int nb = *count;
    const int sy = SY(g), sz = SZ(g);
    const size_t T = (size_t)g.total;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        BlockCell bc = block_cell(g, list[b], threadIdx.x);
        if (!bc.inside) continue;
        const int i = bc.i, j = bc.j, k = bc.k;
        int id = gidx(g, i, j, k);
        for (int c = 0; c < 3; c++) {
            int w = g.ni + (c == 0), h = g.nj + (c == 1), d = g.nk + (c == 2);
            if (i >= w || j >= h || k >= d) continue;
            unsigned char *ly = layer + c * T;
            if (ly[id] != LAYER_UNKNOWN) continue;
            // an UNKNOWN cell is never on the border (those are DONE), so all six neighbours exist
            float *f = vel + c * T;
            const int off[6] = {-1, 1, -sy, sy, -sz, sz};
            const int ii[6] = {i - 1, i + 1, i, i, i, i};
            const int jj[6] = {j, j, j - 1, j + 1, j, j};
            const int kk[6] = {k, k, k, k, k - 1, k + 1};
            float sum = 0.0f;
            int cnt = 0;
            bool reached = false;
            for (int n = 0; n < 6; n++) {
                if (ly[id + off[n]] < L) {  // KNOWN at the start of this layer
                    sum += f[id + off[n]];
                    cnt++;
                    // only KNOWN cells with interior indices push the front (loop bounds 1..dim-2,
                    // src/macvelocityfield.cpp:604-606)
                    if (ii[n] >= 1 && ii[n] <= w - 2 && jj[n] >= 1 && jj[n] <= h - 2 && kk[n] >= 1 && kk[n] <= d - 2)
                        reached = true;
                }
            }
            if (reached) {
                f[id] = sum / (float)cnt;
                ly[id] = (unsigned char)L;
            }
        }
    }
}

void extrapolate_velocity(Sim &s) {
    const Grid &g = s.g;
    grid_list_ensure(s);
    // valid faces only exist inside the near-liquid list: blocks outside it keep flag 0 and layer "unknown"
    CUDA_CHECK(cudaMemsetAsync(s.ext_flag, 0, sizeof(int) * (size_t)g.nblocks, s.stream));
    FLIP_LAUNCH_SYNC(k_extrap_init, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                     (const unsigned char *)s.valid, s.layer, s.ext_flag);
    FLIP_LAUNCH(k_extrap_dilate, cdiv(g.nblocks, 256), 256, s.stream, g, (const int *)s.ext_flag, s.ext_flag2);
    FLIP_LAUNCH_SYNC(k_compact_blocks, 1, 1024, s.stream, (const int *)s.ext_flag2, g.nblocks, s.ext_list, s.ext_count);
    int G = s.num_sms * 4 < g.nblocks ? s.num_sms * 4 : g.nblocks;
    for (int L = 1; L <= s.extrap_layers; L++) {
        FLIP_LAUNCH(k_extrap_layer, G, CG_THREADS, s.stream, g, (const int *)s.ext_list, (const int *)s.ext_count, s.vel, s.layer, L);
    }
    s.kernel_launches += 3 + s.extrap_layers;
    KERNEL_CHECK();
}

// ------------------------------------------------------------------------------------------
// pressure gradient update + validity
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CG_THREADS) k_apply_pressure(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                 const float *__restrict__ phi, const float *__restrict__ pr,
                                 const float *__restrict__ weight, float *__restrict__ vel,
                                 unsigned char *__restrict__ valid, float dt, float minfrac) {
  FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
    int id = gidx(g, i, j, k);
    size_t T = (size_t)g.total;
    const int st[3] = {1, SY(g), SZ(g)};
    for (int c = 0; c < 3; c++) {
        int w = g.ni + (c == 0), h = g.nj + (c == 1), d = g.nk + (c == 2);
        if (i >= w || j >= h || k >= d) continue;
        int n = c == 0 ? g.ni : (c == 1 ? g.nj : g.nk);
        int x = c == 0 ? i : (c == 1 ? j : k);
        float v = 0.0f;
        unsigned char ok = 0;
        // interior faces only: loops run 1..n-1 (src/fluidsimulation.cpp:612, 628, 643)
        if (x >= 1 && x < n && weight[c * T + id] > 0 && face_borders_fluid(g, phi, c, i, j, k)) {
            float p0 = pr[id - st[c]], p1 = pr[id];
            // ParticleLevelSet::getFaceWeight{U,V,W} = fractionInside(phi(x-1), phi(x))
            float theta = fmaxf(frac_inside2(phi[id - st[c]], phi[id]), minfrac);
            v = vel[c * T + id] + (-dt * (p1 - p0) / (g.dx * theta));
            ok = 1;
        }
        vel[c * T + id] = v;   // non-valid faces are zeroed (src/fluidsimulation.cpp:658-687)
        valid[c * T + id] = ok;
    }
  }
}

void apply_pressure(Sim &s, float dt) {
    const Grid &g = s.g;
    grid_list_ensure(s);
    FLIP_LAUNCH(k_apply_pressure, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.phi_liq, (const float *)s.pressure, (const float *)s.weight, s.vel, s.valid, dt, s.minfrac);
    s.kernel_launches++;
    KERNEL_CHECK();
}

// ------------------------------------------------------------------------------------------
// constrain + max|u| (the CFL reduce of the NEXT substep is fused here: nothing touches the grid
// velocities between _constrainVelocityField and the next _cfl())
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CG_THREADS) k_constrain(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                           const float *__restrict__ weight, float *__restrict__ vel, float *__restrict__ saved) {
    FOR_LIST_CELLS(g, glist, gcount, i, j, k) {
        int id = gidx(g, i, j, k);
        size_t T = (size_t)g.total;
        for (int c = 0; c < 3; c++) {
            int w = g.ni + (c == 0), h = g.nj + (c == 1), d = g.nk + (c == 2);
            if (i >= w || j >= h || k >= d) continue;
            if (weight[c * T + id] == 0) {
                vel[c * T + id] = 0.0f;
                saved[c * T + id] = 0.0f;
            }
        }
    }
}

void stage_constrain(Sim &s) {
    const Grid &g = s.g;
    grid_list_ensure(s);
    FLIP_LAUNCH(k_constrain, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                (const float *)s.weight, s.vel, s.saved);
    s.kernel_launches++;
    KERNEL_CHECK();
}

// max |u| over all faces.  |u| >= 0, so the float bit pattern orders like an unsigned integer and
// atomicMax on it is exact and order independent.  Pads are zero and never written.
__global__ void __launch_bounds__(CG_THREADS) k_max_abs(Grid g, const int *__restrict__ glist, const int *__restrict__ gcount,
                                                         const float *__restrict__ f, unsigned *__restrict__ out) {
    __shared__ unsigned wmax[CG_THREADS / 32];
    const size_t T = (size_t)g.total;
    float m = 0.0f;
    FOR_LIST_CELLS(g, glist, gcount, i, j, k) {   // faces outside the list are zero; pads are zero and never written
        int id = gidx(g, i, j, k);
        m = fmaxf(m, fmaxf(fabsf(f[id]), fmaxf(fabsf(f[T + id]), fabsf(f[2 * T + id]))));
    }
    unsigned u = __float_as_uint(m);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned v = __shfl_xor_sync(0xffffffffu, u, o);
        u = v > u ? v : u;
    }
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < CG_THREADS / 32; w++) u = wmax[w] > u ? wmax[w] : u;
        atomicMax(out, u);
    }
}

float compute_max_velocity(Sim &s) {
    const Grid &g = s.g;
    CUDA_CHECK(cudaMemsetAsync(s.maxvel_dev, 0, sizeof(float), s.stream));
    grid_list_ensure(s);
    FLIP_LAUNCH_SYNC(k_max_abs, list_grid(s), CG_THREADS, s.stream, g, (const int *)s.grid_list, (const int *)s.grid_count,
                     (const float *)s.vel, (unsigned *)s.maxvel_dev);
    s.kernel_launches++;
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(s.maxvel_host, s.maxvel_dev, sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    return *s.maxvel_host;
}
