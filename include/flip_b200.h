/*
 * flip_b200.h — C ABI of the B200-native FLIP substep library (libflip_b200.so).
 *
 * This is the drop-in boundary for the reference's FluidSimulation class
 * (rlguy/FLIPViscosity3D, src/fluidsimulation.h:50-134).  The reference has no FFI of its
 * own: its seam is the C++ class, so every entry point below cites the member function
 * (reference file:line) whose behaviour it replaces.  Plain pointers and sizes only; no
 * C++ or torch types cross this boundary.  All functions return 0 on success and a
 * negative FLIP_E* code on failure; flip_last_error() gives the message.  Nothing throws
 * across the ABI.  A handle is bound to the CUDA device that was current at flip_create()
 * and is not thread-safe.  There is NO CPU fallback: without a CUDA device flip_create fails.
 *
 * Wire formats (host memory):
 *   grids      float32, x fastest: idx = i + W*(j + H*k)  (reference Array3d, src/array3d.h:397-400)
 *              cell grids   W,H,D = ni,nj,nk          nodal grids  ni+1,nj+1,nk+1
 *              U faces      ni+1,nj,nk    V faces  ni,nj+1,nk    W faces  ni,nj,nk+1
 *   particles  AoS {pos.xyz, vel.xyz}, 24-byte stride   (FluidParticle, src/fluidsimulation.h:39-48)
 */
#ifndef FLIP_B200_H
#define FLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flip_sim flip_sim;

enum {
    FLIP_OK = 0,
    FLIP_EINVAL = -1,   /* bad argument (the reference would FLUIDSIM_ASSERT -> abort) */
    FLIP_ECUDA = -2,    /* CUDA runtime / kernel failure */
    FLIP_ENOMEM = -3,
    FLIP_ENCCL = -4
};

/* grid fields addressable through flip_get_field / flip_set_field */
enum {
    FLIP_F_LIQUID_SDF = 0,  /* cells  : ParticleLevelSet::_phi      (src/particlelevelset.h:66) */
    FLIP_F_SOLID_SDF = 1,   /* nodes  : MeshLevelSet::_phi          (src/meshlevelset.h) */
    FLIP_F_U = 2, FLIP_F_V = 3, FLIP_F_W = 4,                /* _MACVelocity           */
    FLIP_F_SAVED_U = 5, FLIP_F_SAVED_V = 6, FLIP_F_SAVED_W = 7, /* _savedVelocityField  */
    FLIP_F_WEIGHT_U = 8, FLIP_F_WEIGHT_V = 9, FLIP_F_WEIGHT_W = 10, /* _weightGrid      */
    FLIP_F_PRESSURE = 11,   /* cells  : result of PressureSolver::solve */
    FLIP_F_VISCOSITY = 12,  /* nodes-sized (ni+1)(nj+1)(nk+1) : _viscosity */
    FLIP_F_VOL_CENTER = 13, /* ViscosityVolumeGrid (src/viscositysolver.h:55-90): cells */
    FLIP_F_VOL_U = 14, FLIP_F_VOL_V = 15, FLIP_F_VOL_W = 16,   /* face-sized */
    FLIP_F_VOL_EDGE_U = 17, /* ni, nj+1, nk+1 */
    FLIP_F_VOL_EDGE_V = 18, /* ni+1, nj, nk+1 */
    FLIP_F_VOL_EDGE_W = 19  /* ni+1, nj+1, nk */
};

typedef struct flip_stats {
    int64_t substeps;            /* substeps taken since creation */
    int64_t particles;
    int64_t kernel_launches;     /* kernels this library launched since creation */
    int32_t pressure_iterations; /* last substep */
    int32_t pressure_converged;
    int32_t pressure_active_blocks;
    int32_t viscosity_iterations;
    int32_t viscosity_converged;
    int32_t viscosity_active_blocks;
    int32_t viscosity_applied;   /* 1 if the solution was written back */
    int32_t reserved;
    double pressure_residual;    /* max|r| at exit */
    double viscosity_residual;
    double pressure_rhs_max;
    double viscosity_rhs_max;
    /* device time (CUDA events) of the last substep, milliseconds:
     * [0] liquid sdf  [1] P2G+extrapolate+save  [2] body force  [3] viscosity
     * [4] projection  [5] constrain  [6] G2P+advect  [7] whole substep */
    float stage_ms[8];
    float pressure_solve_ms;
    float viscosity_solve_ms;
    int64_t pressure_unknowns;   /* rows of the last pressure system */
    int64_t viscosity_unknowns;  /* rows (U+V+W faces) of the last viscosity system */
    float viscosity_setup_ms;    /* part of viscosity_solve_ms spent before the first CG iteration (volumes, rows, multigrid set-up) */
    float reserved2;
} flip_stats;

/* ---- lifetime: FluidSimulation::initialize (src/fluidsimulation.cpp:26-43).  The domain-box
 * boundary the reference builds in _initializeBoundary is NOT created here: upload the solid
 * SDF with flip_set_solid_sdf (the C++ shim in host/ computes it like the reference does). */
int flip_create(int ni, int nj, int nk, float dx, flip_sim **out);
int flip_destroy(flip_sim *h);
const char *flip_last_error(flip_sim *h);   /* h may be NULL: error of a failed flip_create */
int flip_synchronize(flip_sim *h);

/* ---- scene state ---- */
/* _solidSDF after addBoundary/resetBoundary (src/fluidsimulation.cpp:45-62): nodal phi, negative
 * inside solid.  Recomputes the face weights (_computeWeights, :549-582) and the viscosity
 * face states (src/viscositysolver.cpp:80-123), which only depend on the solid. */
int flip_set_solid_sdf(flip_sim *h, const float *phi_nodal);
/* public `particles` vector (src/fluidsimulation.h:63): replaces / reads all particles, in the
 * caller's order (the library sorts internally and un-sorts on read) */
int flip_set_particles(flip_sim *h, const float *pos_vel_aos, int64_t n);
int flip_get_particles(flip_sim *h, float *pos_vel_aos, int64_t capacity, int64_t *n_out);
int flip_num_particles(flip_sim *h, int64_t *n_out);
/* Output path (src/main.cpp:14-40 writes positions only): positions in the caller's order, 12 bytes per particle, copied to
 * `xyz_pinned` (from flip_host_alloc) on a separate copy stream while the next substeps run; flip_output_wait blocks until
 * every outstanding export has landed.  Two exports may be in flight. */
int flip_get_positions_async(flip_sim *h, float *xyz_pinned, int64_t capacity, int64_t *n_out);
int flip_output_wait(flip_sim *h);
/* setViscosity(float) / setViscosity(Array3d<float>&) (src/fluidsimulation.cpp:99-124): v >= 0 */
int flip_set_viscosity_uniform(flip_sim *h, float v);
int flip_set_viscosity_grid(flip_sim *h, const float *v_nodesized);
/* setGravity (src/fluidsimulation.cpp:126-132) */
int flip_set_gravity(flip_sim *h, float gx, float gy, float gz);

/* ---- scene construction on the device (optional: flip_set_solid_sdf / flip_set_particles take host-built scenes) ----
 * Same results as the reference's init-time code wherever the substep can tell the difference (csrc/scene.cu): exact
 * point-triangle distances and crossing-parity signs (MeshLevelSet::calculateSignedDistanceField,
 * src/meshlevelset.cpp:138-347), the reference's seeding order and libc rand() sequence.  Meshes: nv x 3 float vertices,
 * nt x 3 int32 vertex indices, inside the domain (else FLIP_EINVAL, where the reference asserts).
 * flip_reset_boundary = FluidSimulation::resetBoundary / the boundary initialize() builds (src/fluidsimulation.cpp:60-62,
 * 198-239); flip_add_boundary_mesh = addBoundary (:45-58); flip_add_liquid_mesh = addLiquid (:64-97), appending to the
 * particles already loaded (before the first substep). */
int flip_reset_boundary(flip_sim *h);
int flip_add_boundary_mesh(flip_sim *h, const float *verts, int nv, const int32_t *tris, int nt, int inverted);
int flip_add_liquid_mesh(flip_sim *h, const float *verts, int nv, const int32_t *tris, int nt, int64_t *added_out);
/* nodal signed distance of a mesh on the handle's grid, (ni+1)(nj+1)(nk+1) floats, x fastest (tests, tools) */
int flip_mesh_sdf(flip_sim *h, const float *verts, int nv, const int32_t *tris, int nt, float *out_nodal);
/* the library's restatement of glibc rand() that seeding draws from: process-wide like libc's, unseeded = srand(1) */
int flip_srand(unsigned int seed);
int flip_rand(void);

/* ---- time stepping ---- */
/* FluidSimulation::advance(float dt) (src/fluidsimulation.cpp:135-168): CFL substep loop */
int flip_advance(flip_sim *h, float dt, int *substeps_out);
/* body of that loop for one substep of the given size */
int flip_substep(flip_sim *h, float substep);
/* _cfl() (src/fluidsimulation.cpp:241-269): 5*dx / max|u|, +inf when the field is zero */
int flip_cfl(flip_sim *h, float *out);

/* ---- individual stages of the substep, for stage-by-stage parity checks ---- */
int flip_stage_update_liquid_sdf(flip_sim *h);              /* _updateLiquidSDF        :354-362 */
int flip_stage_advect_velocity_field(flip_sim *h);          /* _advectVelocityField    :500-519 */
int flip_stage_add_body_force(flip_sim *h, float dt);       /* _addBodyForce           :271-312 */
int flip_stage_apply_viscosity(flip_sim *h, float dt);      /* _applyViscosity         :170-196 */
int flip_stage_project(flip_sim *h, float dt);              /* _project                :522-531 */
int flip_stage_constrain(flip_sim *h);                      /* _constrainVelocityField :696-729 */
int flip_stage_advect_particles(flip_sim *h, float dt);     /* _advectFluidParticles   :315-339 */
int flip_solve_pressure(flip_sim *h, float dt);             /* _solvePressure          :584-596 */
int flip_apply_pressure(flip_sim *h, float dt);             /* _applyPressure          :598-688 */
int flip_extrapolate(flip_sim *h);                          /* _extrapolateVelocityField :690-694 */
int flip_viscosity_volumes(flip_sim *h);                    /* ViscositySolver::_computeVolumeGrid */

/* ---- field access (wire format above) ---- */
int flip_get_field(flip_sim *h, int field, float *out);
int flip_set_field(flip_sim *h, int field, const float *in);
/* ValidVelocityComponentGrid (src/macvelocityfield.h:38-53): one byte per face, comp 0/1/2 */
int flip_get_valid(flip_sim *h, int comp, uint8_t *out);
int flip_set_valid(flip_sim *h, int comp, const uint8_t *in);

/* ---- parameters and diagnostics ----
 * names: "pressure_tol" "pressure_maxit" "viscosity_tol" "viscosity_maxit" "viscosity_accept" "pic_ratio" "cfl"
 *        (defaults = the reference's: src/pressuresolver.h:224-226, src/viscositysolver.h:200-202);
 *        "maxit_scale" (default 40: both solves may take maxit x 40 iterations before the reference's cap / accept /
 *        fail rule applies - the pressure solve here is diagonal-preconditioned and needs ~7x the reference's MIC(0)
 *        iterations; 1 = the reference's caps) "cg_chunk" "verbose";
 *        "mg_compact" 1 = the sweeps of the first explicit multigrid level read compact rows (default), "mg_tma",
 *        "mg_xgroup", "pressure_resident", "use_block_lists": kernel variants, all on by default, kept switchable for A/B;
 *        "viscosity_precond" 2 = Galerkin multigrid (default), 0 = diagonal;
 *        "viscosity_operator" 0 = rows with the exact face-volume term (default), 1 = rows with the reference's
 *        fp32-rounded diagonal, bit for bit (strict parity in the stiff regime, several times more iterations);
 *        "mg_sweeps" "mg_coarse_sweeps" "mg_omega" "mg_levels" "mg_chunk" "mg_flexible" tune the V-cycle */
int flip_set_param(flip_sim *h, const char *name, double value);
int flip_get_stats(flip_sim *h, flip_stats *out);
/* Measurement aid (no reference counterpart): average duration, by CUDA events on the library's stream, of `reps`
 * back-to-back launches of a named hot kernel on the data of the last viscosity solve, and the algorithmic bytes
 * one launch moves.  names: "gmg_sweep_l1" (Jacobi sweep on the first explicit multigrid level), "visc_apply"
 * (matrix-free coupled-face stencil apply of the CG). */
int flip_time_kernel(flip_sim *h, const char *name, int reps, float *ms_per_launch, uint64_t *algorithmic_bytes);

/* Device-side timing for callers (no reference counterpart): records a CUDA event on the library's stream into one of
 * four slots; flip_event_elapsed_ms waits for `slot_to` and returns the device time between the two records.  bench.py
 * brackets its timed region with these (torch.cuda.Event only sees torch's own stream). */
int flip_event_record(flip_sim *h, int slot);
int flip_event_elapsed_ms(flip_sim *h, int slot_from, int slot_to, float *ms);

/* ---- multi-GPU (no reference counterpart: the reference is single-threaded) ----
 * One process per GPU.  Rank 0 calls flip_dist_unique_id and ships the 128 bytes to the other
 * ranks (bench.py uses torch.distributed for that); every rank then calls flip_dist_init on its own
 * handle.  Every rank must hold the same scene (same solid SDF, particles, parameters) and issue the
 * same calls: the CG solves are decomposed into k-slabs with NCCL halo exchange and all-reduce, the
 * particle/grid stages run replicated (multi-GPU). */
int flip_dist_unique_id(void *out128);
int flip_dist_init(flip_sim *h, int rank, int nranks, const void *unique_id128);
/* Optional, after flip_dist_init: map the other ranks' exchange buffers (CUDA IPC) so that the per-
 * iteration reductions and halo planes go through peer memory over NVLink instead of NCCL.  Every
 * rank exports flip_dist_p2p_blob_size() bytes; the caller concatenates them in rank order (e.g. with
 * an all-gather) and hands the result to every rank's import. */
int flip_dist_p2p_blob_size(void);
int flip_dist_p2p_export(flip_sim *h, void *out);
int flip_dist_p2p_import(flip_sim *h, const void *all_blobs_in_rank_order);

/* pinned host buffers for callers that want asynchronous copies */
int flip_host_alloc(void **ptr, uint64_t bytes);
int flip_host_free(void *ptr);

/* library identification: "flip_b200 <version> sm_100a" (or "... cpu-emu" for the test-only build) */
const char *flip_version(void);

#ifdef __cplusplus
}
#endif
#endif
