#include "fluidsimulation.h"
#include <cstdio>
#include <cstdlib>
#include "../../include/flip_b200.h"

namespace {

void require(bool ok, const char *what) { flip_require(ok, what); }

void check(flip_sim *h, int rc, const char *what) {
    if (rc != FLIP_OK) {
        fprintf(stderr, "FluidSimulation: %s failed (%d): %s\n", what, rc, flip_last_error(h));
        abort();
    }
}

}  // namespace

FluidSimulation::FluidSimulation() : _h(0), _isize(0), _jsize(0), _ksize(0), _dx(0), _boundaryDirty(false), _lastSubsteps(0) {}

FluidSimulation::~FluidSimulation() {
    if (_h) flip_destroy(_h);
}

void FluidSimulation::initialize(int i, int j, int k, float dx) {
    if (_h) { flip_destroy(_h); _h = 0; }
    _isize = i; _jsize = j; _ksize = k; _dx = dx;
    int rc = flip_create(i, j, k, dx, &_h);
    if (rc != FLIP_OK) {
        fprintf(stderr, "FluidSimulation::initialize: flip_create failed (%d): %s\n", rc, flip_last_error(0));
        abort();
    }
    particles.clear();
    _scene.initialize(i, j, k, dx);
    _boundaryDirty = true;
}

void FluidSimulation::_uploadBoundary() {
    if (!_boundaryDirty) return;
    check(_h, flip_set_solid_sdf(_h, _scene.solidSDF.data().data()), "flip_set_solid_sdf");
    _boundaryDirty = false;
}

void FluidSimulation::addBoundary(TriangleMesh &boundary, bool isInverted) {
    require(_h != 0, "initialize() must be called first");
    _scene.addBoundary(boundary, isInverted);
    _boundaryDirty = true;
}

void FluidSimulation::resetBoundary() {
    require(_h != 0, "initialize() must be called first");
    _scene.resetBoundary();
    _boundaryDirty = true;
}

void FluidSimulation::addLiquid(TriangleMesh &mesh) {
    require(_h != 0, "initialize() must be called first");
    _scene.addLiquid(mesh, particles);
}

void FluidSimulation::setViscosity(float value) {
    require(_h != 0, "initialize() must be called first");
    require(value >= 0.0, "viscosity must be >= 0");
    check(_h, flip_set_viscosity_uniform(_h, value), "flip_set_viscosity_uniform");
}

void FluidSimulation::setViscosity(Array3d<float> &vgrid) {
    require(_h != 0, "initialize() must be called first");
    require(vgrid.width == _isize + 1 && vgrid.height == _jsize + 1 && vgrid.depth == _ksize + 1,
            "viscosity grid must be (i+1) x (j+1) x (k+1)");
    const float *v = vgrid.getRawArray();
    for (int n = 0; n < vgrid.getNumElements(); n++) require(v[n] >= 0.0, "viscosity must be >= 0");
    check(_h, flip_set_viscosity_grid(_h, v), "flip_set_viscosity_grid");
}

void FluidSimulation::setGravity(vmath::vec3 g) { setGravity(g.x, g.y, g.z); }

void FluidSimulation::setGravity(float gx, float gy, float gz) {
    require(_h != 0, "initialize() must be called first");
    check(_h, flip_set_gravity(_h, gx, gy, gz), "flip_set_gravity");
}

void FluidSimulation::setVerbose(bool v) {
    if (_h) flip_set_param(_h, "verbose", v ? 1.0 : 0.0);
}

// The whole CFL substep loop runs on the device.  `particles` is public and may have been edited
// by the caller between frames, so it is uploaded before and read back after every frame
// (24 B/particle each way, once per FRAME, not per substep).
void FluidSimulation::advance(float dt) {
    require(_h != 0, "initialize() must be called first");
    _uploadBoundary();
    // FluidParticle is 2 x vec3 = 6 packed floats: the wire format of flip_set_particles
    const float *in = particles.empty() ? 0 : &particles[0].position.x;
    check(_h, flip_set_particles(_h, in, (int64_t)particles.size()), "flip_set_particles");
    check(_h, flip_advance(_h, dt, &_lastSubsteps), "flip_advance");
    int64_t n = 0;
    float *out = particles.empty() ? 0 : &particles[0].position.x;
    check(_h, flip_get_particles(_h, out, (int64_t)particles.size(), &n), "flip_get_particles");
}

// ------------------------------------------------------------------------------------------
// plain-C helpers for non-C++ hosts (python tests / bench): scene construction on the host
// ------------------------------------------------------------------------------------------
extern "C" {

// signed distance of a mesh on the node grid; out has (ni+1)(nj+1)(nk+1) floats
void fliphost_mesh_sdf(int ni, int nj, int nk, float dx, const float *verts, int nv, const int *tris, int nt,
                       int band, float *out) {
    TriangleMesh m;
    for (int i = 0; i < nv; i++) m.vertices.push_back(vmath::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
    for (int i = 0; i < nt; i++) m.triangles.push_back(Triangle(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]));
    MeshLevelSet ls(ni, nj, nk, dx);
    ls.calculateSignedDistanceField(m, band);
    const std::vector<float> &d = ls.data();
    for (size_t i = 0; i < d.size(); i++) out[i] = d[i];
}

// device-free scene builder (FlipScene) for python hosts
void *fliphost_scene_create(int ni, int nj, int nk, float dx) {
    FlipScene *s = new FlipScene();
    s->initialize(ni, nj, nk, dx);
    return s;
}
void fliphost_scene_destroy(void *h) { delete (FlipScene *)h; }
static TriangleMesh to_mesh(const float *verts, int nv, const int *tris, int nt) {
    TriangleMesh m;
    for (int i = 0; i < nv; i++) m.vertices.push_back(vmath::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
    for (int i = 0; i < nt; i++) m.triangles.push_back(Triangle(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]));
    return m;
}
void fliphost_scene_add_boundary(void *h, const float *verts, int nv, const int *tris, int nt, int inverted) {
    TriangleMesh m = to_mesh(verts, nv, tris, nt);
    ((FlipScene *)h)->addBoundary(m, inverted != 0);
}
// appends to the scene's particle list; returns the new particle count
struct SceneParticles { std::vector<FluidParticle> p; };
long long fliphost_scene_add_liquid(void *h, const float *verts, int nv, const int *tris, int nt, void **particles_io) {
    TriangleMesh m = to_mesh(verts, nv, tris, nt);
    if (!*particles_io) *particles_io = new SceneParticles();
    SceneParticles *sp = (SceneParticles *)*particles_io;
    ((FlipScene *)h)->addLiquid(m, sp->p);
    return (long long)sp->p.size();
}
void fliphost_particles_get(void *particles, float *out_aos) {
    SceneParticles *sp = (SceneParticles *)particles;
    for (size_t i = 0; i < sp->p.size(); i++) {
        const FluidParticle &q = sp->p[i];
        out_aos[6 * i] = q.position.x; out_aos[6 * i + 1] = q.position.y; out_aos[6 * i + 2] = q.position.z;
        out_aos[6 * i + 3] = q.velocity.x; out_aos[6 * i + 4] = q.velocity.y; out_aos[6 * i + 5] = q.velocity.z;
    }
}
void fliphost_particles_free(void *particles) { delete (SceneParticles *)particles; }
void fliphost_scene_get_solid_sdf(void *h, float *out) {
    const std::vector<float> &d = ((FlipScene *)h)->solidSDF.data();
    for (size_t i = 0; i < d.size(); i++) out[i] = d[i];
}
void fliphost_srand(unsigned seed) { srand(seed); }
// PLY writer/reader of the drop-in TriangleMesh (positions only), for format checks
int fliphost_write_points_ply(const char *path, const float *xyz, long long n) {
    TriangleMesh m;
    m.vertices.reserve((size_t)n);
    for (long long i = 0; i < n; i++) m.vertices.push_back(vmath::vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    m.writeMeshToPLY(path);
    return 0;
}
int fliphost_write_points_obj(const char *path, const float *xyz, long long n) {
    TriangleMesh m;
    for (long long i = 0; i < n; i++) m.vertices.push_back(vmath::vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    m.writeMeshToOBJ(path);
    return 0;
}

}  // extern "C"
