// Device-resident simulation state and the stage entry points (host side of the kernels).
#pragma once
#include "grid.h"
#include "heap.h"
#include "xch.h"
#include <string>
#include <vector>

struct CGState {
    double rho;      // r.z of the current iterate
    double resid;    // max|r|
    double tol;      // absolute tolerance on max|r|
    double bmax;     // max|b|
    int iter;
    int done;        // 1 = stop iterating
    int converged;   // 1 = resid under tol
    int maxit;
    int fail;        // 1 = rho was 0/NaN at start (pcgsolver.h:264-267)
    int first;       // single-reduction variant: no previous direction yet
    double alpha;    // single-reduction variant: previous step length
};

struct SolveStats {
    int iters; int converged; int unknowns; int skipped; double resid; double bmax; float ms; int blocks;
};

#define FLIP_CG_MAXGRID 1024  // upper bound on the persistent CG grid (partials per reduction)

struct GridBar { unsigned gen; int broken; };   // grid-wide exchange state of the persistent kernels (resident.h)

struct Sim {
    Grid g;
    cudaStream_t stream = 0;
    int num_sms = 1;
    std::string last_error;

    // --- parameters (reference defaults: src/fluidsimulation.h:121-130) ---
    float gravity[3] = {0.0f, -9.81f, 0.0f};
    float cfl_number = 5.0f;
    float minfrac = 0.01f;
    float pic_ratio = 0.05f;
    float particle_radius = 0.0f;       // (float)(dx*1.01*sqrt(3)/2), src/fluidsimulation.cpp:36
    int extrap_layers = 7;              // ceil(CFL)+2, src/fluidsimulation.cpp:692
    double pressure_tol = 1e-9;         // absolute, src/pressuresolver.h:224
    int pressure_maxit = 200;           // the reference's MICCG(0) cap (pressuresolver.h:225)
    int pressure_maxit_scale = 40;      // this library's Jacobi-PCG is allowed maxit*scale iterations
    double visc_tol = 1e-6;             // relative to max|rhs|, src/viscositysolver.h:200
    double visc_accept = 10.0;          // src/viscositysolver.h:201
    int visc_maxit = 700;               // src/viscositysolver.h:202
    int visc_maxit_scale = 40;
    int visc_warm_start = 0;            // start the viscosity CG from the current velocity instead of 0
    int visc_operator = 0;              // 0 = exact mass term (default), 1 = the reference's fp32-rounded diagonal (strict parity)
    int visc_precond = 2;               // 2 = Galerkin multigrid (gmg.h, default), 0 = diagonal
    int mg_sweeps = 2;                  // damped-Jacobi sweeps before = after the coarse correction
    int mg_coarse_sweeps = 24;
    float mg_omega = 0.5f;
    float mg_minvol = 0.02f;            // coarse faces need at least this liquid volume fraction
    int mg_prune = 0;                   // 1: drop coarse strain terms that touch air faces (made it worse: off)
    float mg_alpha = 1.0f;              // scale of the coarse-grid correction
    int mg_levels = 8;                  // cap on the number of levels
    // Galerkin multigrid: sweeps on level 0 / level 1 (0 = mg_sweeps).  Level 1 is the expensive one (235 stored
    // coefficients per row, 8x fewer rows than level 0's matrix-free stencil): ONE sweep there and two everywhere
    // else costs no iterations (scipy prototype: 32 vs 31 at 128^3) and halves the level-1 traffic.
    int mg_sweeps_l0 = 3, mg_sweeps_l1 = 1;
    int mg_tma = 1;                     // explicit-level sweeps stage their coefficient rows with TMA bulk copies (gmg.h)
    int mg_xgroup = 1;                  // first explicit level: restriction over groups of 8 x-neighbour rows (gmg.h k_gmg_restrict_x)
    int mg_compact = 1;                 // first explicit level: sweeps read compact rows (160 of 240 slots, gmg.h k_gmg_compact_rows)
    int mg_build = 1;                   // Galerkin products: 1 = gather form (gmg.h k_gmg_build_g), 0 = lane-ordered scatter (bit-identical, slower)
    int mg_dense = 1;                   // exact dense solve on the first level with <= mg_dense_rows rows (else Jacobi sweeps there)
    int mg_dense_rows = 128;            // single-CTA Gauss-Jordan: 0.3 ms at 128 rows, 7 ms at 304 (measured) - keep it small
    int mg_chunk = 8;                   // multigrid CG iterations per graph replay / host convergence poll
    int mg_flexible = 1;                // Polak-Ribiere beta in the multigrid-preconditioned CG
    int cg_chunk = 32;
    // 0 = textbook PCG (3 kernels, 2 reduction points per iteration), 1 = single-reduction
    // Chronopoulos-Gear recurrences (2 kernels).  Measured on B200 at 256^3: the fused update kernel wins
    // for the scalar pressure system (9.3 vs 10.5 ms) and loses for the 3-component viscosity system
    // (86 vs 70 us/iteration: 18 fp64 values in flight per thread cost occupancy).
    int pres_resident = 1;              // pressure CG as ONE persistent launch with its state on chip when it fits (resident.h)
    int cg_variant_pressure = 1;
    int cg_variant_viscosity = 0;
    int cg_grid_mult = 2;               // persistent CG grid = SMs x this (CTAs of 512 threads)                  // CG iterations launched between host convergence polls
    int verbose = 0;
    int use_graphs = 1;                 // replay the CG iteration chunk from a CUDA graph
    void *cg_graph[2] = {0, 0};         // cudaGraphExec_t: [0] pressure, [1] viscosity
    unsigned long long cg_graph_tag[2] = {0, 0};   // chunk, variant and exchange epoch the graph was captured under
    long long cg_graph_launches[2] = {0, 0};
    bool viscosity_nonzero = true;      // reference initial viscosity is 1.0 everywhere

    // --- particles (SoA, cell-binned every substep) ---
    long long np = 0, cap = 0;
    float *p[2][6] = {{0}};   // [buffer][px,py,pz,vx,vy,vz]
    unsigned *pid[2] = {0, 0};
    int cur = 0;              // active particle buffer
    int *cell_of = 0;         // padded cell index per particle
    int *cell_start = 0;      // [total+1] exclusive scan of counts (padded layout)
    int *cell_cursor = 0;     // [total]
    int *scan_tmp = 0;
    bool binned = false;

    // --- fields, all in the padded layout, 3-component arrays are [3*total] ---
    float *phi_liq = 0;       // cell-centred liquid SDF
    float *phi_sol = 0;       // nodal solid SDF
    float *sol_center = 0;    // solid SDF at cell centres (mean of 8 nodes)
    float *vel = 0;           // MAC velocity U,V,W
    float *saved = 0;         // copy taken after P2G+extrapolation
    float *weight = 0;        // solid face weights (static)
    unsigned char *valid = 0; // per face validity
    unsigned char *layer = 0; // extrapolation layer id per face
    unsigned char *fstate = 0;// viscosity face state (1 = solid, 0 = fluid), static
    float *viscosity = 0;     // (ni+1)(nj+1)(nk+1) grid
    float *pressure = 0;      // float pressure of the last projection
    float *maxvel_dev = 0;    // [1] max |u| over faces
    float *maxvel_host = 0;   // pinned

    // --- solver workspaces ---
    float4 *pcoef = 0;        // pressure stencil {diag, +i, +j, +k}
    double *cg_x = 0, *cg_r = 0, *cg_s = 0, *cg_q = 0;   // [3*total] (pressure uses component 0)
    double *cg_z = 0;         // [3*total] preconditioned residual (multigrid mode; u = M^-1 r in the single-reduction CG)
    double *cg_w = 0;         // [3*total] w = A u (single-reduction CG)
    void *gmg = 0;            // Galerkin multigrid hierarchy (GMG*, gmg.h / viscosity.cu)
    float *vvol = 0;          // 7 volume grids [7*total]: center,U,V,W,edgeU,edgeV,edgeW
    float *vnode = 0;         // 7 nodal phi grids [7*total]
    unsigned char *vvalid = 0;// dilated liquid mask
    float *vcoef = 0;         // 4 coefficient grids [4*total]: center, edgeU, edgeV, edgeW
    float *vdiag = 0;         // [3*total] row diagonals (0 = not an unknown)
    float *vmass = 0;         // [3*total] mass term of every row as the CG operator applies it (viscosity.cu k_visc_rows)
    // near-liquid block list of the grid stages (fields.cu grid_list_ensure): blocks within one block of a particle or a
    // liquid cell, plus the blocks that were in the list during the previous substep (their fields go back to defaults)
    int *grid_flag0 = 0, *grid_flag = 0, *grid_dirty = 0, *grid_dirty_next = 0, *grid_list = 0, *grid_count = 0;   // [nblocks] x5, [1]
    unsigned long long world_epoch = 1, grid_list_epoch = 0;   // particles / liquid SDF changed <-> list built for
    int use_block_lists = 1;   // 0: every grid stage sweeps all blocks (debug / A-B timing)
    int *ext_flag = 0, *ext_flag2 = 0, *ext_list = 0, *ext_count = 0;   // extrapolation block list [nblocks] x3, [1]
    int *blk_flag = 0;        // [nblocks]
    int *blk_list = 0;        // [nblocks]
    int *blk_count = 0;       // [1]
    int *cell_list = 0;       // [total] compact list of cells with >= 1 unknown (this solve)
    int *cell_count = 0;      // [1]
    double *part = 0;         // [6 * FLIP_MAX_RANKS * FLIP_CG_MAXGRID] reduction partials (kind-major, cg.h)
    GridBar *grid_bar = 0; unsigned long long *grid_slots = 0;      // barrier state of the persistent kernels (resident.h)
    CGState *cgst = 0;        // [2] ping-pong
    CGState *cgst_host = 0;   // pinned
    int *count_host = 0;      // pinned [2]: active blocks, unknowns
    int *unk_count = 0;       // [1] device

    // --- multi-GPU (dist.cu, xch.h): replicated state in a symmetric heap, work cut into k-slabs, results delivered by
    // peer-memory stores.  `sharded` is set once every rank's heap is mapped (flip_dist_p2p_import); until then several
    // ranks are plain replicas.
    SymHeap heap;
    int rank = 0, nranks = 1;
    void *nccl = 0;           // ncclComm_t (rendezvous / barrier only: nothing on the substep path calls NCCL)
    bool sharded = false;
    Link *link = 0;           // this rank's hand-shake counters (in the heap: peers store into it)
    Link **link_peers = 0;    // device table [FLIP_MAX_RANKS]
    Cuts *cuts = 0;           // slab cuts of the current substep (device)
    int *plane_count = 0;     // [nk + 1] liquid cells per k-plane
    double **part_peers = 0;  // device table [FLIP_MAX_RANKS]: `part` of every rank
    unsigned long long xch_epoch = 1;   // bumped whenever the exchange set-up changes: captured graphs are keyed by it
    // A CG solve is only cut into slabs when it is big enough to pay for its per-kernel hand-shakes (a few microseconds
    // each, three per pressure iteration): below this many unknowns (whole system; decided from the previous solve so that
    // every rank takes the same branch) every rank runs the pressure solve in full on its own copy.  Measured on 2 B200:
    // 0.62 M unknowns, 9.4 ms replicated vs 22 ms sharded; 8 ranks, 2.0 M unknowns (sheet 512^3), 12 ms vs 15 ms.
    long long shard_min_unknowns = 8000000;
    int pres_last_unknowns = 0;
    int xch_nbr_wait = 0;               // 1: ghost-plane consumers wait for their two k-neighbours only (xch.h xch_enter); off until measured at N >= 4
    double xch_timeout_s = 20.0;        // a rank that waits longer than this for its peers gives up (Link::status)
    int *xch_status_host = 0;           // pinned copy of Link::status, fetched with every convergence poll

    // asynchronous position export (flip_get_positions_async): copy stream, two staging buffers, their events
    void *out_stream = 0, *out_ready[2] = {0, 0}, *out_done[2] = {0, 0};
    float *out_buf[2] = {0, 0};
    size_t out_cap[2] = {0, 0};
    bool out_used[2] = {false, false};
    int out_next = 0;
    void *user_ev[4] = {0, 0, 0, 0};   // cudaEvent_t slots of flip_event_record (device-side timing for callers)

    // stats of the last substep
    SolveStats pres_stats = {0, 0, 0, 0, 0, 0, 0, 0};
    SolveStats visc_stats = {0, 0, 0, 0, 0, 0, 0, 0};
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float visc_setup_ms = 0;
    long long substeps = 0;
    long long kernel_launches = 0;

    float *vc(int c) { return vel + (size_t)c * g.total; }
};

// a stage that every rank runs in full on its own copy: the handle looks unsharded for the duration of the guard
struct ReplicateGuard {
    Sim &s; bool was;
    ReplicateGuard(Sim &sim, bool on) : s(sim), was(sim.sharded) { if (on) s.sharded = false; }
    ~ReplicateGuard() { s.sharded = was; }
};

// api.cu
template <class T>
static inline void heap_alloc(Sim &s, T *&p, size_t n) {   // zero-filled
    p = (T *)s.heap.alloc(n * sizeof(T));
    CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s.stream));
}
template <class T>
static inline void heap_free(Sim &s, T *&p) { if (p) s.heap.release(p); p = nullptr; }
void sim_alloc(Sim &s, int ni, int nj, int nk, float dx);
void sim_free(Sim &s);
void sim_reserve_particles(Sim &s, long long n);

// fields.cu
void grid_list_ensure(Sim &s);               // (re)build the near-liquid block list if particles or the liquid SDF changed
void grid_list_mark_all_dirty(Sim &s);       // a caller wrote arbitrary field data: next lists cover every block once
void grid_list_end_substep(Sim &s, bool all_stages_ran);
void solid_precompute(Sim &s);               // weights, cell-centre solid phi, face states
void stage_add_body_force(Sim &s, float dt);
void extrapolate_velocity(Sim &s);
void stage_constrain(Sim &s);
float compute_max_velocity(Sim &s);
void apply_pressure(Sim &s, float dt);

// particles.cu
void bin_particles(Sim &s);
void stage_update_liquid_sdf(Sim &s);
void stage_advect_velocity_field(Sim &s);    // P2G + mask + extrapolate + save
void stage_advect_particles(Sim &s, float dt);

// scene.cu
void scene_reset_boundary(Sim &s);
void scene_add_boundary(Sim &s, const float *verts, int nv, const int *tris, int nt, bool inverted);
long long scene_add_liquid(Sim &s, const float *verts, int nv, const int *tris, int nt);
void scene_mesh_sdf(Sim &s, const float *verts, int nv, const int *tris, int nt, float *out_host);
void scene_srand(unsigned int seed);
int scene_rand_next();

// pressure.cu / viscosity.cu
void stage_project(Sim &s, float dt);
void solve_pressure(Sim &s, float dt);
void stage_apply_viscosity(Sim &s, float dt);
void viscosity_volumes(Sim &s);
void viscosity_free(Sim &s);
int viscosity_time_kernel(Sim &s, const char *name, int reps, float *ms_per_launch, unsigned long long *alg_bytes);

// dist.cu
void dist_init(Sim &s, int rank, int nranks, const void *unique_id);
void dist_shutdown(Sim &s);
void dist_get_unique_id(void *out128);
int dist_p2p_blob_size();
void dist_p2p_export(Sim &s, void *out);
void dist_p2p_import(Sim &s, const void *all_blobs);
void dist_p2p_shutdown(Sim &s);
// exchange helpers; every one is a no-op unless the handle is sharded
Xch xch_of(Sim &s);                          // what the kernels take (nranks == 1 when not sharded)
const Cuts *xch_cuts(Sim &s);                // device cuts, or null when not sharded
int xch_rank(Sim &s);
void xch_update_cuts(Sim &s);                // balanced k-slabs from the current liquid SDF
void xch_push_halo(Sim &s, const Grid &g, void *field, size_t elem, int ncomp, int level, int halo);   // ghost planes -> k-neighbours
void xch_push_gather(Sim &s, const Grid &g, void *field, size_t elem, int ncomp, int level);           // whole slab -> every rank
void xch_push_rows(Sim &s, const int *rng_dev, void *base, size_t row_bytes);                          // row range -> every rank
void xch_barrier(Sim &s);                    // everything pushed so far has arrived everywhere
void xch_check(Sim &s);                      // throws if a hand-shake timed out
void xch_status_fetch(Sim &s);               // enqueue a copy of the hand-shake status next to a convergence poll ...
bool xch_status_bad(Sim &s);                 // ... and read it after the stream synchronisation

// substep driver (api.cu)
void sim_substep(Sim &s, float dt);
int sim_advance(Sim &s, float dt);
