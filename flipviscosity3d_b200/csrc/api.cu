// C ABI (include/flip_b200.h), memory management, wire-format conversion, substep driver.
#include "sim.h"
#include "../../include/flip_b200.h"
#include <cmath>
#include <cstring>
#include <limits>

struct flip_sim { Sim s; };

static std::string g_create_error;

#ifdef FLIP_CPU_EMU
#define FLIP_BUILD_KIND "cpu-emu"
#else
#define FLIP_BUILD_KIND "sm_100a"
#endif

// every device buffer comes from the handle's symmetric heap (heap.h), zero filled
template <class T>
static void dev_alloc(Sim &s, T *&p, size_t n) { heap_alloc(s, p, n); }

void sim_alloc(Sim &s, int ni, int nj, int nk, float dx) {
    s.g = make_grid(ni, nj, nk, dx);
    const Grid &g = s.g;
    size_t T = (size_t)g.total;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    s.num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    // Heap sizing: the fixed fields below take ~330 B per padded cell, the multigrid vectors ~70 B more; particles and
    // the explicit multigrid rows come on top.  Later needs grow the heap by further chunks (until the peers map it).
    s.heap.first_chunk = (size_t)540 * T + ((size_t)64 << 20);
    s.heap.grow_chunk = (size_t)130 * T + ((size_t)64 << 20);
    // _particleRadius = (float)(_dx * 1.01*sqrt(3.0)/2.0)  (src/fluidsimulation.cpp:36)
    s.particle_radius = (float)((double)dx * 1.01 * sqrt(3.0) / 2.0);
    s.extrap_layers = (int)ceil(s.cfl_number) + 2;

    dev_alloc(s, s.cell_start, T + 1);
    dev_alloc(s, s.cell_cursor, T);
    dev_alloc(s, s.scan_tmp, T / 2048 + 2);
    dev_alloc(s, s.phi_liq, T);
    dev_alloc(s, s.phi_sol, T);
    dev_alloc(s, s.sol_center, T);
    dev_alloc(s, s.vel, 3 * T);
    dev_alloc(s, s.saved, 3 * T);
    dev_alloc(s, s.weight, 3 * T);
    dev_alloc(s, s.valid, 3 * T);
    dev_alloc(s, s.layer, 3 * T);
    dev_alloc(s, s.fstate, 3 * T);
    dev_alloc(s, s.viscosity, T);
    dev_alloc(s, s.pressure, T);
    dev_alloc(s, s.maxvel_dev, 1);
    dev_alloc(s, s.pcoef, T);
    dev_alloc(s, s.cg_x, 3 * T);
    dev_alloc(s, s.cg_r, 3 * T);
    dev_alloc(s, s.cg_s, 3 * T);
    dev_alloc(s, s.cg_q, 3 * T);
    dev_alloc(s, s.cg_z, 3 * T);
    dev_alloc(s, s.cg_w, 3 * T);
    dev_alloc(s, s.vvol, 7 * T);
    dev_alloc(s, s.vnode, 7 * T);
    dev_alloc(s, s.vvalid, T);
    dev_alloc(s, s.vcoef, 4 * T);
    dev_alloc(s, s.vdiag, 3 * T);
    dev_alloc(s, s.vmass, 3 * T);
    dev_alloc(s, s.grid_flag0, (size_t)g.nblocks); dev_alloc(s, s.grid_flag, (size_t)g.nblocks);
    dev_alloc(s, s.grid_dirty, (size_t)g.nblocks); dev_alloc(s, s.grid_dirty_next, (size_t)g.nblocks);
    dev_alloc(s, s.grid_list, (size_t)g.nblocks); dev_alloc(s, s.grid_count, 1);
    dev_alloc(s, s.ext_flag, (size_t)g.nblocks); dev_alloc(s, s.ext_flag2, (size_t)g.nblocks);
    dev_alloc(s, s.ext_list, (size_t)g.nblocks); dev_alloc(s, s.ext_count, 1);
    dev_alloc(s, s.blk_flag, (size_t)g.nblocks);
    dev_alloc(s, s.blk_list, (size_t)g.nblocks);
    dev_alloc(s, s.blk_count, 1);
    dev_alloc(s, s.unk_count, 1);
    dev_alloc(s, s.cell_list, T);
    dev_alloc(s, s.cell_count, 1);
    dev_alloc(s, s.part, 6 * (size_t)FLIP_MAX_RANKS * FLIP_CG_MAXGRID);
    dev_alloc(s, s.part_peers, FLIP_MAX_RANKS);
    dev_alloc(s, s.link, 1);
    dev_alloc(s, s.link_peers, FLIP_MAX_RANKS);
    dev_alloc(s, s.cuts, 1);
    dev_alloc(s, s.plane_count, (size_t)nk + 2);
    dev_alloc(s, s.cgst, 2);
    dev_alloc(s, s.grid_bar, 1);
    dev_alloc(s, s.grid_slots, (size_t)2 * (1024 + 1) * 4);   // resident.h: [2][GRID_MAX_CTAS + 1][GRID_SLOT_WORDS]
    CUDA_CHECK(cudaMallocHost((void **)&s.cgst_host, sizeof(CGState)));
    CUDA_CHECK(cudaMallocHost((void **)&s.count_host, 2 * sizeof(int)));
    CUDA_CHECK(cudaMallocHost((void **)&s.maxvel_host, sizeof(float)));
    CUDA_CHECK(cudaMallocHost((void **)&s.xch_status_host, sizeof(int)));
    *s.xch_status_host = 0;
    *s.count_host = 0;
    *s.maxvel_host = 0;
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    // initial solid: none (phi = +large everywhere) until flip_set_solid_sdf; viscosity 1.0
    // like the reference's initialize (src/fluidsimulation.cpp:39)
}

static void free_particles(Sim &s) {
    for (int b = 0; b < 2; b++) {
        for (int f = 0; f < 6; f++) heap_free(s, s.p[b][f]);
        heap_free(s, s.pid[b]);
    }
    heap_free(s, s.cell_of);
    s.cap = 0;
}

void sim_reserve_particles(Sim &s, long long n) {
    if (n <= s.cap) return;
    free_particles(s);
    long long cap = n + n / 8 + 1024;
    for (int b = 0; b < 2; b++) {
        for (int f = 0; f < 6; f++) dev_alloc(s, s.p[b][f], (size_t)cap);
        dev_alloc(s, s.pid[b], (size_t)cap);
    }
    dev_alloc(s, s.cell_of, 2 * (size_t)cap);
    s.cap = cap;
}

void sim_free(Sim &s) {
    cudaStreamSynchronize(s.stream);
#ifndef FLIP_CPU_EMU
    // graphs may hold captured NCCL kernels: destroy them before the communicator
    for (int q = 0; q < 2; q++) if (s.cg_graph[q]) { cudaGraphExecDestroy((cudaGraphExec_t)s.cg_graph[q]); s.cg_graph[q] = 0; }
#endif
    if (s.out_stream) {
        cudaStreamSynchronize((cudaStream_t)s.out_stream);
        for (int q = 0; q < 2; q++) { if (s.out_buf[q]) cudaFree(s.out_buf[q]); cudaEventDestroy((cudaEvent_t)s.out_ready[q]); cudaEventDestroy((cudaEvent_t)s.out_done[q]); }
        cudaStreamDestroy((cudaStream_t)s.out_stream);
        s.out_stream = 0;
    }
    for (int q = 0; q < 4; q++) if (s.user_ev[q]) { cudaEventDestroy((cudaEvent_t)s.user_ev[q]); s.user_ev[q] = 0; }
    try { dist_shutdown(s); } catch (...) {}
    viscosity_free(s);
    s.heap.destroy();   // every device buffer of the handle
    if (s.cgst_host) cudaFreeHost(s.cgst_host);
    if (s.count_host) cudaFreeHost(s.count_host);
    if (s.maxvel_host) cudaFreeHost(s.maxvel_host);
    if (s.xch_status_host) cudaFreeHost(s.xch_status_host);
    if (s.stream) cudaStreamDestroy(s.stream);
}

// ------------------------------------------------------------------------------------------
// wire format <-> padded layout
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void k_unpack(Grid g, const T *__restrict__ in, T *__restrict__ field, int w, int h, int d) {
    int i, j, k;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!unflatten(t, w, h, d, i, j, k)) return;
    field[gidx(g, i, j, k)] = in[t];
}
template <class T>
__global__ void k_pack(Grid g, const T *__restrict__ field, T *__restrict__ out, int w, int h, int d) {
    int i, j, k;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!unflatten(t, w, h, d, i, j, k)) return;
    out[t] = field[gidx(g, i, j, k)];
}

__global__ void k_fill_box(Grid g, float *__restrict__ field, int w, int h, int d, float v) {
    int i, j, k;
    if (!unflatten((long long)blockIdx.x * blockDim.x + threadIdx.x, w, h, d, i, j, k)) return;
    field[gidx(g, i, j, k)] = v;
}

__global__ void k_particles_in(const float *__restrict__ aos, float *px, float *py, float *pz, float *vx, float *vy,
                               float *vz, unsigned *pid, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    px[t] = aos[6 * t]; py[t] = aos[6 * t + 1]; pz[t] = aos[6 * t + 2];
    vx[t] = aos[6 * t + 3]; vy[t] = aos[6 * t + 4]; vz[t] = aos[6 * t + 5];
    pid[t] = (unsigned)t;
}
__global__ void k_particles_out(float *__restrict__ aos, const float *px, const float *py, const float *pz,
                                const float *vx, const float *vy, const float *vz, const unsigned *pid, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t o = 6 * (size_t)pid[t];
    aos[o] = px[t]; aos[o + 1] = py[t]; aos[o + 2] = pz[t];
    aos[o + 3] = vx[t]; aos[o + 4] = vy[t]; aos[o + 5] = vz[t];
}

struct FieldRef { float *ptr; int w, h, d; };

static bool field_ref(Sim &s, int f, FieldRef &r) {
    const Grid &g = s.g;
    size_t T = (size_t)g.total;
    int ni = g.ni, nj = g.nj, nk = g.nk;
    switch (f) {
        case FLIP_F_LIQUID_SDF: r = {s.phi_liq, ni, nj, nk}; return true;
        case FLIP_F_SOLID_SDF: r = {s.phi_sol, ni + 1, nj + 1, nk + 1}; return true;
        case FLIP_F_U: r = {s.vel, ni + 1, nj, nk}; return true;
        case FLIP_F_V: r = {s.vel + T, ni, nj + 1, nk}; return true;
        case FLIP_F_W: r = {s.vel + 2 * T, ni, nj, nk + 1}; return true;
        case FLIP_F_SAVED_U: r = {s.saved, ni + 1, nj, nk}; return true;
        case FLIP_F_SAVED_V: r = {s.saved + T, ni, nj + 1, nk}; return true;
        case FLIP_F_SAVED_W: r = {s.saved + 2 * T, ni, nj, nk + 1}; return true;
        case FLIP_F_WEIGHT_U: r = {s.weight, ni + 1, nj, nk}; return true;
        case FLIP_F_WEIGHT_V: r = {s.weight + T, ni, nj + 1, nk}; return true;
        case FLIP_F_WEIGHT_W: r = {s.weight + 2 * T, ni, nj, nk + 1}; return true;
        case FLIP_F_PRESSURE: r = {s.pressure, ni, nj, nk}; return true;
        case FLIP_F_VISCOSITY: r = {s.viscosity, ni + 1, nj + 1, nk + 1}; return true;
        case FLIP_F_VOL_CENTER: r = {s.vvol, ni, nj, nk}; return true;
        case FLIP_F_VOL_U: r = {s.vvol + T, ni + 1, nj, nk}; return true;
        case FLIP_F_VOL_V: r = {s.vvol + 2 * T, ni, nj + 1, nk}; return true;
        case FLIP_F_VOL_W: r = {s.vvol + 3 * T, ni, nj, nk + 1}; return true;
        case FLIP_F_VOL_EDGE_U: r = {s.vvol + 4 * T, ni, nj + 1, nk + 1}; return true;
        case FLIP_F_VOL_EDGE_V: r = {s.vvol + 5 * T, ni + 1, nj, nk + 1}; return true;
        case FLIP_F_VOL_EDGE_W: r = {s.vvol + 6 * T, ni + 1, nj + 1, nk}; return true;
    }
    return false;
}

// staging buffer: vnode (7*total floats) is scratch outside the viscosity stage
static float *staging(Sim &s) { return s.vnode; }

static void upload_field(Sim &s, const FieldRef &r, const float *in) {
    size_t n = (size_t)r.w * r.h * r.d;
    CUDA_CHECK(cudaMemcpyAsync(staging(s), in, n * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    auto kern = &k_unpack<float>;
    FLIP_LAUNCH(kern, cdiv((long long)n, 256), 256, s.stream, s.g, (const float *)staging(s), r.ptr, r.w, r.h, r.d);
    KERNEL_CHECK();
    grid_list_mark_all_dirty(s);   // the caller may have put data anywhere: the next block lists cover the whole grid once
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}
static void download_field(Sim &s, const FieldRef &r, float *out) {
    size_t n = (size_t)r.w * r.h * r.d;
    auto kern = &k_pack<float>;
    FLIP_LAUNCH(kern, cdiv((long long)n, 256), 256, s.stream, s.g, (const float *)r.ptr, staging(s), r.w, r.h, r.d);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(out, staging(s), n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

// ------------------------------------------------------------------------------------------
// substep driver
// ------------------------------------------------------------------------------------------
void sim_substep(Sim &s, float dt) {
    cudaEvent_t ev[8];
    for (int i = 0; i < 8; i++) CUDA_CHECK(cudaEventCreate(&ev[i]));
    CUDA_CHECK(cudaEventRecord(ev[0], s.stream));
    stage_update_liquid_sdf(s);
    CUDA_CHECK(cudaEventRecord(ev[1], s.stream));
    stage_advect_velocity_field(s);
    CUDA_CHECK(cudaEventRecord(ev[2], s.stream));
    stage_add_body_force(s, dt);
    CUDA_CHECK(cudaEventRecord(ev[3], s.stream));
    stage_apply_viscosity(s, dt);
    CUDA_CHECK(cudaEventRecord(ev[4], s.stream));
    stage_project(s, dt);
    CUDA_CHECK(cudaEventRecord(ev[5], s.stream));
    stage_constrain(s);
    CUDA_CHECK(cudaEventRecord(ev[6], s.stream));
    stage_advect_particles(s, dt);
    grid_list_end_substep(s, true);
    CUDA_CHECK(cudaEventRecord(ev[7], s.stream));
    CUDA_CHECK(cudaEventSynchronize(ev[7]));
    for (int i = 0; i < 7; i++) CUDA_CHECK(cudaEventElapsedTime(&s.stage_ms[i], ev[i], ev[i + 1]));
    CUDA_CHECK(cudaEventElapsedTime(&s.stage_ms[7], ev[0], ev[7]));
    for (int i = 0; i < 8; i++) CUDA_CHECK(cudaEventDestroy(ev[i]));
    s.substeps++;
}

static float sim_cfl(Sim &s) {
    float maxvel = compute_max_velocity(s);
    // (float)((_CFLConditionNumber * _dx) / maxvel): float arithmetic, +inf when maxvel == 0
    return (s.cfl_number * s.g.dx) / maxvel;
}

int sim_advance(Sim &s, float dt) {
    // FluidSimulation::advance (src/fluidsimulation.cpp:135-168)
    float t = 0;
    int n = 0;
    while (t < dt) {
        float substep = sim_cfl(s);
        if (t + substep > dt) substep = dt - t;
        if (s.verbose) printf("Taking substep of size %f (to %0.3f%% of the frame)\n", substep, 100 * (t + substep) / dt);
        sim_substep(s, substep);
        t += substep;
        n++;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
#define API_BEGIN(h)                                          \
    if (!(h)) return FLIP_EINVAL;                             \
    Sim &s = (h)->s;                                          \
    try {
#define API_END()                                             \
    }                                                         \
    catch (const std::bad_alloc &) { s.last_error = "out of host memory"; return FLIP_ENOMEM; } \
    catch (const std::exception &e) { s.last_error = e.what(); return FLIP_ECUDA; }             \
    return FLIP_OK;

static int fail_inval(Sim &s, const char *msg) { s.last_error = msg; return FLIP_EINVAL; }

extern "C" {

const char *flip_version(void) { return "flip_b200 0.1 " FLIP_BUILD_KIND; }

int flip_create(int ni, int nj, int nk, float dx, flip_sim **out) {
    if (!out) return FLIP_EINVAL;
    *out = nullptr;
    if (ni < 4 || nj < 4 || nk < 4 || !(dx > 0)) { g_create_error = "flip_create: need ni,nj,nk >= 4 and dx > 0"; return FLIP_EINVAL; }
    flip_sim *h = nullptr;
    try {
        h = new flip_sim();
        sim_alloc(h->s, ni, nj, nk, dx);
        Sim &s = h->s;
        const Grid &g = s.g;
        // reference defaults: viscosity 1.0 on the (ni+1)(nj+1)(nk+1) grid; no solid until uploaded
        FLIP_LAUNCH(k_fill_box, cdiv((long long)(ni + 1) * (nj + 1) * (nk + 1), 256), 256, s.stream, g, s.viscosity, ni + 1,
                    nj + 1, nk + 1, 1.0f);
        FLIP_LAUNCH(k_fill_box, cdiv((long long)(ni + 1) * (nj + 1) * (nk + 1), 256), 256, s.stream, g, s.phi_sol, ni + 1,
                    nj + 1, nk + 1, (float)(ni + nj + nk) * dx);
        // defaults of the fields the grid stages only rewrite near the liquid (fields.cu, near-liquid block list):
        // liquid SDF = its "no particle in reach" value 3 dx (src/particlelevelset.cpp:94-96), extrapolation layer = unknown
        FLIP_LAUNCH(k_fill_box, cdiv((long long)ni * nj * nk, 256), 256, s.stream, g, s.phi_liq, ni, nj, nk, 3.0f * dx);
        CUDA_CHECK(cudaMemsetAsync(s.layer, 0xFF, 3 * (size_t)g.total, s.stream));
        KERNEL_CHECK();
        solid_precompute(s);
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
    } catch (const std::exception &e) {
        g_create_error = e.what();
        if (h) { try { sim_free(h->s); } catch (...) {} delete h; }
        return FLIP_ECUDA;
    }
    *out = h;
    return FLIP_OK;
}

int flip_destroy(flip_sim *h) {
    if (!h) return FLIP_EINVAL;
    try { sim_free(h->s); } catch (...) {}
    delete h;
    return FLIP_OK;
}

const char *flip_last_error(flip_sim *h) { return h ? h->s.last_error.c_str() : g_create_error.c_str(); }

int flip_synchronize(flip_sim *h) {
    API_BEGIN(h)
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    API_END()
}

int flip_set_solid_sdf(flip_sim *h, const float *phi) {
    API_BEGIN(h)
    if (!phi) return fail_inval(s, "flip_set_solid_sdf: null pointer");
    FieldRef r;
    field_ref(s, FLIP_F_SOLID_SDF, r);
    upload_field(s, r, phi);
    solid_precompute(s);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    API_END()
}

int flip_set_particles(flip_sim *h, const float *aos, int64_t n) {
    API_BEGIN(h)
    if (n < 0 || (n > 0 && !aos)) return fail_inval(s, "flip_set_particles: bad arguments");
    if (n >= (int64_t)1 << 31) return fail_inval(s, "flip_set_particles: more than 2^31-1 particles");
    sim_reserve_particles(s, n);
    s.np = n;
    s.cur = 0;
    s.binned = false;
    if (n > 0) {
        // AoS staging in the spare particle buffer (6 floats * cap contiguous is not guaranteed, so
        // stage through the cg_q scratch: 3*total doubles >= 6*n floats is checked)
        size_t need = 6 * (size_t)n * sizeof(float);
        float *stage = nullptr;
        bool own = false;
        if (need <= 3 * (size_t)s.g.total * sizeof(double)) stage = (float *)s.cg_q;
        else { CUDA_CHECK(cudaMalloc((void **)&stage, need)); own = true; }
        CUDA_CHECK(cudaMemcpyAsync(stage, aos, need, cudaMemcpyHostToDevice, s.stream));
        FLIP_LAUNCH(k_particles_in, cdiv(n, 256), 256, s.stream, (const float *)stage, s.p[0][0], s.p[0][1], s.p[0][2],
                    s.p[0][3], s.p[0][4], s.p[0][5], s.pid[0], (long long)n);
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (own) cudaFree(stage);
    }
    API_END()
}

int flip_get_particles(flip_sim *h, float *aos, int64_t capacity, int64_t *n_out) {
    API_BEGIN(h)
    if (n_out) *n_out = s.np;
    if (!aos) return FLIP_OK;
    if (capacity < s.np) return fail_inval(s, "flip_get_particles: capacity too small");
    if (s.np > 0) {
        size_t need = 6 * (size_t)s.np * sizeof(float);
        float *stage = nullptr;
        bool own = false;
        if (need <= 3 * (size_t)s.g.total * sizeof(double)) stage = (float *)s.cg_q;
        else { CUDA_CHECK(cudaMalloc((void **)&stage, need)); own = true; }
        int c = s.cur;
        FLIP_LAUNCH(k_particles_out, cdiv(s.np, 256), 256, s.stream, stage, (const float *)s.p[c][0], (const float *)s.p[c][1],
                    (const float *)s.p[c][2], (const float *)s.p[c][3], (const float *)s.p[c][4], (const float *)s.p[c][5],
                    (const unsigned *)s.pid[c], s.np);
        KERNEL_CHECK();
        CUDA_CHECK(cudaMemcpyAsync(aos, stage, need, cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (own) cudaFree(stage);
    }
    API_END()
}

int flip_reset_boundary(flip_sim *h) {
    API_BEGIN(h)
    scene_reset_boundary(s);
    API_END()
}

int flip_add_boundary_mesh(flip_sim *h, const float *verts, int nv, const int *tris, int nt, int inverted) {
    API_BEGIN(h)
    if (!verts || !tris) return fail_inval(s, "flip_add_boundary_mesh: null pointer");
    scene_add_boundary(s, verts, nv, tris, nt, inverted != 0);
    API_END()
}

int flip_add_liquid_mesh(flip_sim *h, const float *verts, int nv, const int *tris, int nt, int64_t *added_out) {
    API_BEGIN(h)
    if (!verts || !tris) return fail_inval(s, "flip_add_liquid_mesh: null pointer");
    long long n = scene_add_liquid(s, verts, nv, tris, nt);
    if (added_out) *added_out = n;
    API_END()
}

int flip_mesh_sdf(flip_sim *h, const float *verts, int nv, const int *tris, int nt, float *out_nodal) {
    API_BEGIN(h)
    if (!verts || !tris || !out_nodal || nv <= 0 || nt <= 0) return fail_inval(s, "flip_mesh_sdf: bad arguments");
    scene_mesh_sdf(s, verts, nv, tris, nt, out_nodal);
    API_END()
}

int flip_srand(unsigned int seed) { scene_srand(seed); return FLIP_OK; }
int flip_rand(void) { return scene_rand_next(); }

// positions only, caller's order, 12 bytes per particle (what the reference's exporters write)
__global__ void k_positions_out(float *__restrict__ xyz, const float *px, const float *py, const float *pz, const unsigned *pid, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t o = 3 * (size_t)pid[t];
    xyz[o] = px[t]; xyz[o + 1] = py[t]; xyz[o + 2] = pz[t];
}

int flip_get_positions_async(flip_sim *h, float *xyz_pinned, int64_t capacity, int64_t *n_out) {
    API_BEGIN(h)
    if (n_out) *n_out = s.np;
    if (!xyz_pinned || capacity < s.np) return fail_inval(s, "flip_get_positions_async: null pointer or capacity too small");
    if (s.np == 0) return FLIP_OK;
    if (!s.out_stream) {
        CUDA_CHECK(cudaStreamCreateWithFlags((cudaStream_t *)&s.out_stream, cudaStreamNonBlocking));
        for (int q = 0; q < 2; q++) { cudaEvent_t e; CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); s.out_ready[q] = (void *)e;
                                      CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); s.out_done[q] = (void *)e; }
    }
    const int b = s.out_next;
    s.out_next ^= 1;
    if ((size_t)s.np > s.out_cap[b]) {
        if (s.out_buf[b]) { CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)s.out_stream)); cudaFree(s.out_buf[b]); }
        s.out_cap[b] = (size_t)s.np + (size_t)s.np / 8 + 1024;
        CUDA_CHECK(cudaMalloc((void **)&s.out_buf[b], s.out_cap[b] * 3 * sizeof(float)));
        s.out_used[b] = false;
    }
    // the staging buffer may still be draining from the export before last
    if (s.out_used[b]) CUDA_CHECK(cudaStreamWaitEvent(s.stream, (cudaEvent_t)s.out_done[b], 0));
    const int c = s.cur;
    FLIP_LAUNCH(k_positions_out, cdiv(s.np, 256), 256, s.stream, s.out_buf[b], (const float *)s.p[c][0], (const float *)s.p[c][1],
                (const float *)s.p[c][2], (const unsigned *)s.pid[c], s.np);
    s.kernel_launches++;
    KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord((cudaEvent_t)s.out_ready[b], s.stream));
    CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)s.out_stream, (cudaEvent_t)s.out_ready[b], 0));
    CUDA_CHECK(cudaMemcpyAsync(xyz_pinned, s.out_buf[b], (size_t)s.np * 3 * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)s.out_stream));
    CUDA_CHECK(cudaEventRecord((cudaEvent_t)s.out_done[b], (cudaStream_t)s.out_stream));
    s.out_used[b] = true;
    API_END()
}

int flip_output_wait(flip_sim *h) {
    API_BEGIN(h)
    if (s.out_stream) CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)s.out_stream));
    API_END()
}

int flip_num_particles(flip_sim *h, int64_t *n_out) {
    if (!h || !n_out) return FLIP_EINVAL;
    *n_out = h->s.np;
    return FLIP_OK;
}

int flip_set_viscosity_uniform(flip_sim *h, float v) {
    API_BEGIN(h)
    if (!(v >= 0.0f)) return fail_inval(s, "setViscosity: value must be >= 0 (FLUIDSIM_ASSERT, src/fluidsimulation.cpp:100)");
    const Grid &g = s.g;
    FLIP_LAUNCH(k_fill_box, cdiv((long long)(g.ni + 1) * (g.nj + 1) * (g.nk + 1), 256), 256, s.stream, g, s.viscosity,
                g.ni + 1, g.nj + 1, g.nk + 1, v);
    KERNEL_CHECK();
    s.viscosity_nonzero = v > 0.0f;
    grid_list_mark_all_dirty(s);   // the viscosity-derived fields are only maintained while there is viscosity
    API_END()
}

int flip_set_viscosity_grid(flip_sim *h, const float *v) {
    API_BEGIN(h)
    if (!v) return fail_inval(s, "flip_set_viscosity_grid: null pointer");
    const Grid &g = s.g;
    size_t n = (size_t)(g.ni + 1) * (g.nj + 1) * (g.nk + 1);
    bool nonzero = false;
    for (size_t q = 0; q < n; q++) {
        if (!(v[q] >= 0.0f)) return fail_inval(s, "setViscosity: values must be >= 0 (FLUIDSIM_ASSERT, src/fluidsimulation.cpp:119)");
        if (v[q] > 0.0f) nonzero = true;
    }
    FieldRef r;
    field_ref(s, FLIP_F_VISCOSITY, r);
    upload_field(s, r, v);   // (marks every block dirty)
    s.viscosity_nonzero = nonzero;
    API_END()
}

int flip_set_gravity(flip_sim *h, float gx, float gy, float gz) {
    if (!h) return FLIP_EINVAL;
    h->s.gravity[0] = gx; h->s.gravity[1] = gy; h->s.gravity[2] = gz;
    return FLIP_OK;
}

int flip_advance(flip_sim *h, float dt, int *substeps_out) {
    API_BEGIN(h)
    int n = sim_advance(s, dt);
    if (substeps_out) *substeps_out = n;
    API_END()
}

int flip_substep(flip_sim *h, float substep) {
    API_BEGIN(h)
    sim_substep(s, substep);
    API_END()
}

int flip_cfl(flip_sim *h, float *out) {
    API_BEGIN(h)
    if (!out) return fail_inval(s, "flip_cfl: null pointer");
    *out = sim_cfl(s);
    API_END()
}

#define STAGE0(name, call)            \
    int name(flip_sim *h) {           \
        API_BEGIN(h)                  \
        call;                         \
        CUDA_CHECK(cudaStreamSynchronize(s.stream)); \
        API_END()                     \
    }
#define STAGE1(name, call)            \
    int name(flip_sim *h, float dt) { \
        API_BEGIN(h)                  \
        call;                         \
        CUDA_CHECK(cudaStreamSynchronize(s.stream)); \
        API_END()                     \
    }

STAGE0(flip_stage_update_liquid_sdf, stage_update_liquid_sdf(s))
STAGE0(flip_stage_advect_velocity_field, stage_advect_velocity_field(s))
STAGE1(flip_stage_add_body_force, stage_add_body_force(s, dt))
STAGE1(flip_stage_apply_viscosity, stage_apply_viscosity(s, dt))
STAGE1(flip_stage_project, stage_project(s, dt))
STAGE0(flip_stage_constrain, stage_constrain(s))
STAGE1(flip_stage_advect_particles, stage_advect_particles(s, dt))
STAGE1(flip_solve_pressure, solve_pressure(s, dt))
STAGE1(flip_apply_pressure, apply_pressure(s, dt))
STAGE0(flip_extrapolate, extrapolate_velocity(s))
STAGE0(flip_viscosity_volumes, viscosity_volumes(s))

int flip_get_field(flip_sim *h, int field, float *out) {
    API_BEGIN(h)
    FieldRef r;
    if (!out || !field_ref(s, field, r)) return fail_inval(s, "flip_get_field: bad field id or null pointer");
    download_field(s, r, out);
    API_END()
}

int flip_set_field(flip_sim *h, int field, const float *in) {
    API_BEGIN(h)
    FieldRef r;
    if (!in || !field_ref(s, field, r)) return fail_inval(s, "flip_set_field: bad field id or null pointer");
    if (field == FLIP_F_SOLID_SDF) return flip_set_solid_sdf(h, in);
    if (field == FLIP_F_VISCOSITY) return flip_set_viscosity_grid(h, in);
    upload_field(s, r, in);
    API_END()
}

int flip_get_valid(flip_sim *h, int comp, uint8_t *out) {
    API_BEGIN(h)
    if (!out || comp < 0 || comp > 2) return fail_inval(s, "flip_get_valid: bad arguments");
    const Grid &g = s.g;
    int w = g.ni + (comp == 0), hh = g.nj + (comp == 1), d = g.nk + (comp == 2);
    size_t n = (size_t)w * hh * d;
    unsigned char *stage = (unsigned char *)staging(s);
    auto kern = &k_pack<unsigned char>;
    FLIP_LAUNCH(kern, cdiv((long long)n, 256), 256, s.stream, g, (const unsigned char *)(s.valid + (size_t)comp * g.total), stage, w, hh, d);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(out, stage, n, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    API_END()
}

int flip_set_valid(flip_sim *h, int comp, const uint8_t *in) {
    API_BEGIN(h)
    if (!in || comp < 0 || comp > 2) return fail_inval(s, "flip_set_valid: bad arguments");
    const Grid &g = s.g;
    int w = g.ni + (comp == 0), hh = g.nj + (comp == 1), d = g.nk + (comp == 2);
    size_t n = (size_t)w * hh * d;
    unsigned char *stage = (unsigned char *)staging(s);
    CUDA_CHECK(cudaMemcpyAsync(stage, in, n, cudaMemcpyHostToDevice, s.stream));
    auto kern = &k_unpack<unsigned char>;
    FLIP_LAUNCH(kern, cdiv((long long)n, 256), 256, s.stream, g, (const unsigned char *)stage, s.valid + (size_t)comp * g.total, w, hh, d);
    KERNEL_CHECK();
    grid_list_mark_all_dirty(s);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    API_END()
}

int flip_set_param(flip_sim *h, const char *name, double value) {
    if (!h || !name) return FLIP_EINVAL;
    Sim &s = h->s;
    std::string n(name);
    if (n == "pressure_tol") s.pressure_tol = value;
    else if (n == "pressure_maxit") s.pressure_maxit = (int)value;
    else if (n == "viscosity_tol") s.visc_tol = value;
    else if (n == "viscosity_maxit") s.visc_maxit = (int)value;
    else if (n == "viscosity_accept") s.visc_accept = value;
    else if (n == "maxit_scale") { s.pressure_maxit_scale = (int)value; s.visc_maxit_scale = (int)value; }
    else if (n == "cg_chunk") s.cg_chunk = (int)value;
    else if (n == "use_graphs") s.use_graphs = (int)value;
    else if (n == "cg_variant") { s.cg_variant_pressure = s.cg_variant_viscosity = (int)value; }
    else if (n == "cg_variant_pressure") { s.cg_variant_pressure = (int)value; }
    else if (n == "cg_variant_viscosity") { s.cg_variant_viscosity = (int)value; }
    else if (n == "cg_grid_mult") { s.cg_grid_mult = (int)value < 1 ? 1 : (int)value; s.xch_epoch++; }
    else if (n == "viscosity_precond") s.visc_precond = (int)value;
    else if (n == "viscosity_operator") s.visc_operator = (int)value;
    else if (n == "viscosity_warm_start") s.visc_warm_start = (int)value;
    else if (n == "mg_sweeps") s.mg_sweeps = (int)value;
    else if (n == "mg_coarse_sweeps") s.mg_coarse_sweeps = (int)value;
    else if (n == "mg_omega") s.mg_omega = (float)value;
    else if (n == "mg_alpha") s.mg_alpha = (float)value;
    else if (n == "mg_minvol") s.mg_minvol = (float)value;
    else if (n == "mg_prune") s.mg_prune = (int)value;
    else if (n == "mg_levels") s.mg_levels = (int)value;
    else if (n == "mg_flexible") s.mg_flexible = (int)value;
    else if (n == "mg_chunk") s.mg_chunk = (int)value;
    else if (n == "dist_p2p") { if ((int)value == 0) dist_p2p_shutdown(s); }   // unmap the peers: plain replicas again
    else if (n == "shard_min_unknowns") s.shard_min_unknowns = (long long)value;
    else if (n == "xch_nbr_wait") { s.xch_nbr_wait = (int)value; s.xch_epoch++; }
    else if (n == "xch_timeout_s") s.xch_timeout_s = value;                    // takes effect at the next flip_dist_p2p_import
    else if (n == "mg_tma") s.mg_tma = (int)value;
    else if (n == "mg_dense") s.mg_dense = (int)value;
    else if (n == "mg_build") s.mg_build = (int)value;
    else if (n == "mg_xgroup") s.mg_xgroup = (int)value;
    else if (n == "mg_compact") s.mg_compact = (int)value;
    else if (n == "pressure_resident") s.pres_resident = (int)value;
    else if (n == "mg_dense_rows") s.mg_dense_rows = (int)value;
    else if (n == "mg_sweeps_l0") s.mg_sweeps_l0 = (int)value;
    else if (n == "mg_sweeps_l1") s.mg_sweeps_l1 = (int)value;
    else if (n == "pic_ratio") s.pic_ratio = (float)value;
    else if (n == "cfl") { s.cfl_number = (float)value; s.extrap_layers = (int)ceil(s.cfl_number) + 2; }
    else if (n == "use_block_lists") { s.use_block_lists = (int)value; s.world_epoch++; }
    else if (n == "verbose") s.verbose = (int)value;
    else return fail_inval(s, "flip_set_param: unknown parameter name");
    return FLIP_OK;
}

int flip_get_stats(flip_sim *h, flip_stats *out) {
    if (!h || !out) return FLIP_EINVAL;
    Sim &s = h->s;
    memset(out, 0, sizeof(*out));
    out->substeps = s.substeps;
    out->particles = s.np;
    out->kernel_launches = s.kernel_launches;
    out->pressure_iterations = s.pres_stats.iters;
    out->pressure_converged = s.pres_stats.converged;
    out->pressure_active_blocks = s.pres_stats.blocks;
    out->pressure_unknowns = s.pres_stats.unknowns;
    out->viscosity_iterations = s.visc_stats.iters;
    out->viscosity_converged = s.visc_stats.converged;
    out->viscosity_active_blocks = s.visc_stats.blocks;
    out->viscosity_unknowns = s.visc_stats.unknowns;
    out->viscosity_applied = s.visc_stats.skipped == 0 ? 1 : 0;
    out->pressure_residual = s.pres_stats.resid;
    out->viscosity_residual = s.visc_stats.resid;
    out->pressure_rhs_max = s.pres_stats.bmax;
    out->viscosity_rhs_max = s.visc_stats.bmax;
    for (int i = 0; i < 8; i++) out->stage_ms[i] = s.stage_ms[i];
    out->pressure_solve_ms = s.pres_stats.ms;
    out->viscosity_solve_ms = s.visc_stats.ms;
    out->viscosity_setup_ms = s.visc_setup_ms;
    return FLIP_OK;
}

int flip_time_kernel(flip_sim *h, const char *name, int reps, float *ms_per_launch, uint64_t *algorithmic_bytes) {
    if (!h || !name || !ms_per_launch || !algorithmic_bytes) return FLIP_EINVAL;
    Sim &s = h->s;
    try {
        unsigned long long b = 0;
        int rc = viscosity_time_kernel(s, name, reps, ms_per_launch, &b);
        *algorithmic_bytes = b;
        if (rc != 0) return fail_inval(s, rc == -2 ? "flip_time_kernel: unknown kernel name" : "flip_time_kernel: no viscosity solve to time yet");
    } catch (const std::exception &e) { s.last_error = e.what(); return FLIP_ECUDA; }
    return FLIP_OK;
}

int flip_event_record(flip_sim *h, int slot) {
    API_BEGIN(h)
    if (slot < 0 || slot >= 4) return fail_inval(s, "flip_event_record: slot must be 0..3");
    if (!s.user_ev[slot]) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); s.user_ev[slot] = (void *)e; }
    CUDA_CHECK(cudaEventRecord((cudaEvent_t)s.user_ev[slot], s.stream));
    API_END()
}

int flip_event_elapsed_ms(flip_sim *h, int slot_from, int slot_to, float *ms) {
    API_BEGIN(h)
    if (!ms || slot_from < 0 || slot_from >= 4 || slot_to < 0 || slot_to >= 4 || !s.user_ev[slot_from] || !s.user_ev[slot_to])
        return fail_inval(s, "flip_event_elapsed_ms: bad slots (record both first)");
    CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)s.user_ev[slot_to]));
    CUDA_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)s.user_ev[slot_from], (cudaEvent_t)s.user_ev[slot_to]));
    API_END()
}

int flip_dist_unique_id(void *out128) {
    if (!out128) return FLIP_EINVAL;
    try { dist_get_unique_id(out128); } catch (const std::exception &e) { g_create_error = e.what(); return FLIP_ENCCL; }
    return FLIP_OK;
}

int flip_dist_init(flip_sim *h, int rank, int nranks, const void *unique_id128) {
    if (!h) return FLIP_EINVAL;
    Sim &s = h->s;
    if (nranks > 1 && !unique_id128) return fail_inval(s, "flip_dist_init: unique id required");
    try { dist_init(s, rank, nranks, unique_id128); } catch (const std::exception &e) { s.last_error = e.what(); return FLIP_ENCCL; }
    return FLIP_OK;
}

int flip_dist_p2p_blob_size(void) { return dist_p2p_blob_size(); }

int flip_dist_p2p_export(flip_sim *h, void *out) {
    API_BEGIN(h)
    if (!out) return fail_inval(s, "flip_dist_p2p_export: null pointer");
    dist_p2p_export(s, out);
    API_END()
}

int flip_dist_p2p_import(flip_sim *h, const void *all_blobs) {
    API_BEGIN(h)
    if (!all_blobs) return fail_inval(s, "flip_dist_p2p_import: null pointer");
    dist_p2p_import(s, all_blobs);
    API_END()
}

int flip_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return FLIP_EINVAL;
    return cudaMallocHost(ptr, (size_t)bytes) == cudaSuccess ? FLIP_OK : FLIP_ENOMEM;
}
int flip_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? FLIP_OK : FLIP_ECUDA; }

}  // extern "C"
