/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * C-ABI harness around the UNMODIFIED reference sources (rlguy/FLIPViscosity3D,
 * compiled where they lie under /root/reference/src by oracle/Makefile into
 * oracle/_ref/libflipref.so).  It exposes the reference's private per-substep
 * stages (src/fluidsimulation.cpp:135-168) one by one so the CUDA path can be
 * parity-checked stage by stage with teacher forcing, and times them for the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
 * may load this library.  Nothing in flipviscosity3d_b200/ does.
 *
 * Private members are reached with the `#define private public` trick; the std
 * headers are included first so that only the reference's own classes are opened.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cerrno>
#include <string>
#include <vector>
#include <queue>
#include <sstream>
#include <fstream>
#include <iostream>
#include <limits>
#include <algorithm>
#include <stdexcept>
#include <chrono>
#include <time.h>
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define private public
#define protected public
#include "fluidsimulation.h"
#undef private
#undef protected

namespace {

struct Ref {
    FluidSimulation sim;
    int ni, nj, nk;
    float dx;
    // last solver diagnostics
    int viscIters = 0;
    double viscResid = 0.0;
    int viscOk = 0;
    int viscUnknowns = 0;
    int presUnknowns = 0;
    std::vector<float> lastPressure;
    double stageSeconds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool quiet = true;
};

struct StdoutMute {
    // the reference prints progress lines unconditionally; silence them for tests
    std::streambuf *old;
    std::ostringstream sink;
    bool on;
    int savedFd;
    explicit StdoutMute(bool enable) : old(nullptr), on(enable), savedFd(-1) {
        if (on) {
            old = std::cout.rdbuf(sink.rdbuf());
        }
    }
    ~StdoutMute() {
        if (on) {
            std::cout.rdbuf(old);
        }
    }
};

template <class T>
void copyOut(Array3d<T> &a, T *out) {
    std::memcpy(out, a.getRawArray(), sizeof(T) * (size_t)a.width * a.height * a.depth);
}

template <class T>
void copyIn(Array3d<T> &a, const T *in) {
    std::memcpy(a.getRawArray(), in, sizeof(T) * (size_t)a.width * a.height * a.depth);
}

TriangleMesh makeMesh(const float *verts, int nv, const int *tris, int nt) {
    TriangleMesh m;
    m.vertices.resize(nv);
    for (int i = 0; i < nv; i++) {
        m.vertices[i] = vmath::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
    }
    m.triangles.resize(nt);
    for (int i = 0; i < nt; i++) {
        m.triangles[i] = Triangle(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
    }
    return m;
}

double now() {
    return std::chrono::duration<double>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void *ref_create(int ni, int nj, int nk, float dx) {
    Ref *r = new Ref();
    r->ni = ni; r->nj = nj; r->nk = nk; r->dx = dx;
    StdoutMute mute(true);
    r->sim.initialize(ni, nj, nk, dx);
    return r;
}

void ref_destroy(void *h) { delete (Ref *)h; }

void ref_set_quiet(void *h, int q) { ((Ref *)h)->quiet = q != 0; }

/* src/fluidsimulation.cpp:45-58 */
void ref_add_boundary(void *h, const float *verts, int nv, const int *tris, int nt, int inverted) {
    Ref *r = (Ref *)h;
    TriangleMesh m = makeMesh(verts, nv, tris, nt);
    r->sim.addBoundary(m, inverted != 0);
}

void ref_reset_boundary(void *h) { ((Ref *)h)->sim.resetBoundary(); }

/* src/fluidsimulation.cpp:64-97 (uses libc rand(); call ref_srand(1) first for the
 * unseeded-process sequence) */
void ref_add_liquid(void *h, const float *verts, int nv, const int *tris, int nt) {
    Ref *r = (Ref *)h;
    TriangleMesh m = makeMesh(verts, nv, tris, nt);
    r->sim.addLiquid(m);
}

void ref_srand(unsigned int seed) { srand(seed); }

/* standalone mesh SDF (src/meshlevelset.cpp:138-151), for checking the host-side
 * mesh level set of the new library.  out has (ni+1)(nj+1)(nk+1) floats. */
void ref_mesh_sdf(int ni, int nj, int nk, float dx, const float *verts, int nv,
                  const int *tris, int nt, int band, float *out) {
    TriangleMesh m = makeMesh(verts, nv, tris, nt);
    MeshLevelSet ls(ni, nj, nk, dx);
    ls.calculateSignedDistanceField(m, band);
    copyOut(ls._phi, out);
}

void ref_set_viscosity(void *h, float v) { ((Ref *)h)->sim.setViscosity(v); }

void ref_set_viscosity_grid(void *h, const float *v) {
    Ref *r = (Ref *)h;
    copyIn(r->sim._viscosity, v);
}

void ref_set_gravity(void *h, float gx, float gy, float gz) {
    ((Ref *)h)->sim.setGravity(gx, gy, gz);
}

long long ref_num_particles(void *h) { return (long long)((Ref *)h)->sim.particles.size(); }

/* AoS {pos.xyz, vel.xyz}, 24 B stride (src/fluidsimulation.h:39-48) */
void ref_get_particles(void *h, float *out) {
    Ref *r = (Ref *)h;
    size_t n = r->sim.particles.size();
    for (size_t i = 0; i < n; i++) {
        const FluidParticle &p = r->sim.particles[i];
        out[6 * i + 0] = p.position.x; out[6 * i + 1] = p.position.y; out[6 * i + 2] = p.position.z;
        out[6 * i + 3] = p.velocity.x; out[6 * i + 4] = p.velocity.y; out[6 * i + 5] = p.velocity.z;
    }
}

void ref_set_particles(void *h, const float *in, long long n) {
    Ref *r = (Ref *)h;
    r->sim.particles.resize((size_t)n);
    for (long long i = 0; i < n; i++) {
        r->sim.particles[i].position = vmath::vec3(in[6 * i], in[6 * i + 1], in[6 * i + 2]);
        r->sim.particles[i].velocity = vmath::vec3(in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]);
    }
}

/* grids, in the reference's Array3d layout (x fastest, src/array3d.h:397-400) */
void ref_get_solid_sdf(void *h, float *out) { copyOut(((Ref *)h)->sim._solidSDF._phi, out); }
void ref_set_solid_sdf(void *h, const float *in) { copyIn(((Ref *)h)->sim._solidSDF._phi, in); }
void ref_get_liquid_sdf(void *h, float *out) { copyOut(((Ref *)h)->sim._liquidSDF._phi, out); }
void ref_set_liquid_sdf(void *h, const float *in) { copyIn(((Ref *)h)->sim._liquidSDF._phi, in); }

void ref_get_mac(void *h, float *u, float *v, float *w) {
    Ref *r = (Ref *)h;
    copyOut(r->sim._MACVelocity._u, u);
    copyOut(r->sim._MACVelocity._v, v);
    copyOut(r->sim._MACVelocity._w, w);
}
void ref_set_mac(void *h, const float *u, const float *v, const float *w) {
    Ref *r = (Ref *)h;
    copyIn(r->sim._MACVelocity._u, u);
    copyIn(r->sim._MACVelocity._v, v);
    copyIn(r->sim._MACVelocity._w, w);
}
void ref_get_saved_mac(void *h, float *u, float *v, float *w) {
    Ref *r = (Ref *)h;
    copyOut(r->sim._savedVelocityField._u, u);
    copyOut(r->sim._savedVelocityField._v, v);
    copyOut(r->sim._savedVelocityField._w, w);
}
void ref_set_saved_mac(void *h, const float *u, const float *v, const float *w) {
    Ref *r = (Ref *)h;
    if (r->sim._savedVelocityField._u.width != r->ni + 1) {
        r->sim._savedVelocityField = MACVelocityField(r->ni, r->nj, r->nk, r->dx);
    }
    copyIn(r->sim._savedVelocityField._u, u);
    copyIn(r->sim._savedVelocityField._v, v);
    copyIn(r->sim._savedVelocityField._w, w);
}
void ref_get_weights(void *h, float *u, float *v, float *w) {
    Ref *r = (Ref *)h;
    copyOut(r->sim._weightGrid.U, u);
    copyOut(r->sim._weightGrid.V, v);
    copyOut(r->sim._weightGrid.W, w);
}
void ref_get_valid(void *h, unsigned char *u, unsigned char *v, unsigned char *w) {
    Ref *r = (Ref *)h;
    copyOut(r->sim._validVelocities.validU, (bool *)u);
    copyOut(r->sim._validVelocities.validV, (bool *)v);
    copyOut(r->sim._validVelocities.validW, (bool *)w);
}
void ref_set_valid(void *h, const unsigned char *u, const unsigned char *v, const unsigned char *w) {
    Ref *r = (Ref *)h;
    copyIn(r->sim._validVelocities.validU, (const bool *)u);
    copyIn(r->sim._validVelocities.validV, (const bool *)v);
    copyIn(r->sim._validVelocities.validW, (const bool *)w);
}

/* ---- stages of the substep (src/fluidsimulation.cpp:135-168) ---- */

float ref_cfl(void *h) { return ((Ref *)h)->sim._cfl(); }

void ref_stage_update_liquid_sdf(void *h) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._updateLiquidSDF();
}

/* raw P2G of one component before masking (src/fluidsimulation.cpp:364-438) */
void ref_p2g_component(void *h, int dir, float *field, unsigned char *isset) {
    Ref *r = (Ref *)h;
    int w = r->ni + (dir == 0), hh = r->nj + (dir == 1), d = r->nk + (dir == 2);
    Array3d<float> f(w, hh, d, 0.0f);
    Array3d<bool> s(w, hh, d, false);
    r->sim._computeVelocityScalarField(f, s, dir);
    copyOut(f, field);
    copyOut(s, (bool *)isset);
}

void ref_stage_advect_velocity_field(void *h) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._advectVelocityField();
}

void ref_stage_add_body_force(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._addBodyForce(dt);
}

void ref_extrapolate(void *h) {
    Ref *r = (Ref *)h;
    r->sim._extrapolateVelocityField(r->sim._MACVelocity, r->sim._validVelocities);
}

/* Viscosity volumes only (src/viscositysolver.cpp:135-270): fills the 7 grids. */
void ref_viscosity_volumes(void *h, float *center, float *U, float *V, float *W,
                           float *eU, float *eV, float *eW) {
    Ref *r = (Ref *)h;
    ViscositySolverParameters params;
    params.cellwidth = r->sim._dx;
    params.deltaTime = 0.01f;
    params.velocityField = &r->sim._MACVelocity;
    params.liquidSDF = &r->sim._liquidSDF;
    params.solidSDF = &r->sim._solidSDF;
    params.viscosity = &r->sim._viscosity;
    ViscositySolver vs;
    vs._initialize(params);
    vs._computeFaceStateGrid();
    vs._computeVolumeGrid();
    copyOut(vs._volumes.center, center);
    copyOut(vs._volumes.U, U);
    copyOut(vs._volumes.V, V);
    copyOut(vs._volumes.W, W);
    copyOut(vs._volumes.edgeU, eU);
    copyOut(vs._volumes.edgeV, eV);
    copyOut(vs._volumes.edgeW, eW);
}

/* _applyViscosity (src/fluidsimulation.cpp:170-196) with the solver's private
 * tolerance / iteration cap (src/viscositysolver.h:200-202) overridable:
 * tol <= 0 or maxit <= 0 keep the reference defaults (1e-6, 700).
 * Returns 1 if the reference would have written the solution back. */
int ref_stage_apply_viscosity_ex(void *h, float dt, double tol, int maxit) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    FluidSimulation &s = r->sim;
    bool nonzero = false;
    for (int k = 0; k < s._viscosity.depth && !nonzero; k++)
        for (int j = 0; j < s._viscosity.height && !nonzero; j++)
            for (int i = 0; i < s._viscosity.width; i++)
                if (s._viscosity(i, j, k) > 0.0) { nonzero = true; break; }
    r->viscIters = 0; r->viscResid = 0; r->viscOk = 0; r->viscUnknowns = 0;
    if (!nonzero) return 0;

    ViscositySolverParameters params;
    params.cellwidth = s._dx;
    params.deltaTime = dt;
    params.velocityField = &s._MACVelocity;
    params.liquidSDF = &s._liquidSDF;
    params.solidSDF = &s._solidSDF;
    params.viscosity = &s._viscosity;

    ViscositySolver vs;
    if (tol > 0) vs._solverTolerance = tol;
    if (maxit > 0) vs._maxSolverIterations = maxit;
    // body of ViscositySolver::applyViscosityToVelocityField (src/viscositysolver.cpp:41-63)
    vs._initialize(params);
    vs._computeFaceStateGrid();
    vs._computeVolumeGrid();
    vs._computeMatrixIndexTable();
    int matsize = vs._matrixIndex.matrixSize;
    r->viscUnknowns = matsize;
    SparseMatrixd matrix(matsize);
    std::vector<double> rhs(matsize, 0);
    std::vector<double> soln(matsize, 0);
    vs._initializeLinearSystem(matrix, rhs);
    vs._destroyVolumeGrid();
    // body of _solveLinearSystem (src/viscositysolver.cpp:666-690) with diagnostics kept
    PCGSolver<double> solver;
    solver.setSolverParameters(vs._solverTolerance, vs._maxSolverIterations);
    double err = 0; int its = 0;
    bool ok = solver.solve(matrix, rhs, soln, err, its);
    r->viscIters = its; r->viscResid = err;
    bool accept = ok || (its == vs._maxSolverIterations && err < vs._acceptableTolerace);
    r->viscOk = accept ? 1 : 0;
    if (!accept) return 0;
    vs._applySolutionToVelocityField(soln);
    return 1;
}

/* Residual of a CANDIDATE solution in the reference's own assembled viscosity system: builds the system exactly as
 * ViscositySolver::applyViscosityToVelocityField does (src/viscositysolver.cpp:41-63, 374-664) from the current
 * pre-viscosity state, reads x from the candidate MAC fields through the reference's index table
 * (src/viscositysolver.cpp:692-727 in reverse) and evaluates r = b - A x with the reference's CSR multiply
 * (src/pcgsolver/sparsematrix.h:166-176).  out = {max|r|, max|b|, unknowns, max|x|, rows with a non-positive diagonal,
 * max|r| over rows whose face control volume is > 0}.  The state of the simulation is not changed. */
int ref_viscosity_residual(void *h, float dt, const float *u, const float *v, const float *w, double *out) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    FluidSimulation &s = r->sim;
    ViscositySolverParameters params;
    params.cellwidth = s._dx;
    params.deltaTime = dt;
    params.velocityField = &s._MACVelocity;
    params.liquidSDF = &s._liquidSDF;
    params.solidSDF = &s._solidSDF;
    params.viscosity = &s._viscosity;
    ViscositySolver vs;
    vs._initialize(params);
    vs._computeFaceStateGrid();
    vs._computeVolumeGrid();
    vs._computeMatrixIndexTable();
    int n = vs._matrixIndex.matrixSize;
    SparseMatrixd matrix(n);
    std::vector<double> rhs(n, 0), x(n, 0), ax(n, 0), vol(n, 0);
    vs._initializeLinearSystem(matrix, rhs);
    int ni = r->ni, nj = r->nj, nk = r->nk;
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni + 1; i++) {
        int m = vs._matrixIndex.U(i, j, k);
        if (m != -1) { x[m] = u[i + (size_t)(ni + 1) * (j + (size_t)nj * k)]; vol[m] = vs._volumes.U(i, j, k); }
    }
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj + 1; j++) for (int i = 0; i < ni; i++) {
        int m = vs._matrixIndex.V(i, j, k);
        if (m != -1) { x[m] = v[i + (size_t)ni * (j + (size_t)(nj + 1) * k)]; vol[m] = vs._volumes.V(i, j, k); }
    }
    for (int k = 0; k < nk + 1; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
        int m = vs._matrixIndex.W(i, j, k);
        if (m != -1) { x[m] = w[i + (size_t)ni * (j + (size_t)nj * k)]; vol[m] = vs._volumes.W(i, j, k); }
    }
    vs._destroyVolumeGrid();
    FixedSparseMatrix<double> fixed;
    fixed.fromMatrix(matrix);
    multiply(fixed, x, ax);
    double rmax = 0, bmax = 0, xmax = 0, rmax_vol = 0; long long badDiag = 0;
    for (int q = 0; q < n; q++) {
        double res = std::fabs(rhs[q] - ax[q]);
        rmax = std::max(rmax, res); bmax = std::max(bmax, std::fabs(rhs[q])); xmax = std::max(xmax, std::fabs(x[q]));
        if (vol[q] > 0) rmax_vol = std::max(rmax_vol, res);
        if (matrix(q, q) <= 0) badDiag++;
    }
    out[0] = rmax; out[1] = bmax; out[2] = (double)n; out[3] = xmax; out[4] = (double)badDiag; out[5] = rmax_vol;
    return n;
}

void ref_stage_apply_viscosity(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._applyViscosity(dt);
}

void ref_viscosity_diag(void *h, int *iters, double *resid, int *ok, int *unknowns) {
    Ref *r = (Ref *)h;
    *iters = r->viscIters; *resid = r->viscResid; *ok = r->viscOk; *unknowns = r->viscUnknowns;
}

void ref_compute_weights(void *h) { ((Ref *)h)->sim._computeWeights(); }

/* pressure solve only (src/fluidsimulation.cpp:584-596); out = ni*nj*nk floats */
void ref_solve_pressure(void *h, float dt, double tol, int maxit, float *out) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    FluidSimulation &s = r->sim;
    PressureSolverParameters params;
    params.cellwidth = s._dx;
    params.density = 1.0;
    params.deltaTime = dt;
    params.velocityField = &s._MACVelocity;
    params.liquidSDF = &s._liquidSDF;
    params.weightGrid = &s._weightGrid;
    PressureSolver solver;
    if (tol > 0) solver._pressureSolveTolerance = tol;
    if (maxit > 0) solver._maxCGIterations = maxit;
    Array3d<float> p = solver.solve(params);
    r->presUnknowns = solver._matSize;
    copyOut(p, out);
}

void ref_apply_pressure(void *h, float dt, const float *pressure) {
    Ref *r = (Ref *)h;
    Array3d<float> p(r->ni, r->nj, r->nk, 0.0f);
    copyIn(p, pressure);
    r->sim._applyPressure(dt, p);
}

void ref_stage_project(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._project(dt);
}

void ref_stage_constrain(void *h) { ((Ref *)h)->sim._constrainVelocityField(); }

void ref_stage_advect_particles(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim._advectFluidParticles(dt);
}

/* one substep of the given size, stages timed with <chrono>:
 * times[0..6] = sdf, advect-field(P2G+extrapolate), body force, viscosity,
 * project, constrain, advect particles; times[7] = total */
void ref_substep(void *h, float substep, double *times) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    FluidSimulation &s = r->sim;
    double t[8];
    t[0] = now(); s._updateLiquidSDF();
    t[1] = now(); s._advectVelocityField();
    t[2] = now(); s._addBodyForce(substep);
    t[3] = now(); s._applyViscosity(substep);
    t[4] = now(); s._project(substep);
    t[5] = now(); s._constrainVelocityField();
    t[6] = now(); s._advectFluidParticles(substep);
    t[7] = now();
    if (times) {
        for (int i = 0; i < 7; i++) times[i] = t[i + 1] - t[i];
        times[7] = t[7] - t[0];
    }
}

/* FluidSimulation::advance (src/fluidsimulation.cpp:135-168); returns #substeps */
int ref_advance(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    FluidSimulation &s = r->sim;
    float t = 0; int n = 0;
    while (t < dt) {
        float substep = s._cfl();
        if (t + substep > dt) substep = dt - t;
        s._updateLiquidSDF();
        s._advectVelocityField();
        s._addBodyForce(substep);
        s._applyViscosity(substep);
        s._project(substep);
        s._constrainVelocityField();
        s._advectFluidParticles(substep);
        t += substep; n++;
    }
    return n;
}

/* the shipped entry point itself, for checking ref_advance against it */
void ref_advance_native(void *h, float dt) {
    Ref *r = (Ref *)h; StdoutMute mute(r->quiet);
    r->sim.advance(dt);
}

/* PLY writer of the reference (src/trianglemesh.cpp:190-343) on the current particles */
void ref_write_particles_ply(void *h, const char *path) {
    Ref *r = (Ref *)h;
    TriangleMesh m;
    for (size_t i = 0; i < r->sim.particles.size(); i++) m.vertices.push_back(r->sim.particles[i].position);
    m.writeMeshToPLY(path);
}

/* OBJ writer of the reference (src/trianglemesh.cpp:381-418) on the current particles */
void ref_write_particles_obj(void *h, const char *path) {
    Ref *r = (Ref *)h;
    TriangleMesh m;
    for (size_t i = 0; i < r->sim.particles.size(); i++) m.vertices.push_back(r->sim.particles[i].position);
    m.writeMeshToOBJ(path);
}

/* reference PLY loader (src/trianglemesh.cpp:39-63); fails on files < 2048 B.
 * Returns 1 on success and fills counts; call again with buffers to fetch. */
int ref_load_ply(const char *path, float *verts, int *tris, int *nv, int *nt) {
    TriangleMesh m;
    if (!m.loadPLY(path)) return 0;
    *nv = (int)m.vertices.size(); *nt = (int)m.triangles.size();
    if (verts) for (int i = 0; i < *nv; i++) { verts[3*i] = m.vertices[i].x; verts[3*i+1] = m.vertices[i].y; verts[3*i+2] = m.vertices[i].z; }
    if (tris) for (int i = 0; i < *nt; i++) { tris[3*i] = m.triangles[i].tri[0]; tris[3*i+1] = m.triangles[i].tri[1]; tris[3*i+2] = m.triangles[i].tri[2]; }
    return 1;
}

}  // extern "C"
