// Device-free scene construction: solid SDF (domain box + user boundaries) and particle seeding.
// These are the init-time parts of the reference's FluidSimulation (src/fluidsimulation.cpp:45-97,
// 198-239) kept on the host with the reference's algorithms, so a scene starts from the same
// bits; FluidSimulation (fluidsimulation.h) owns one and uploads its results through the C ABI.
#ifndef FLIPB200_SCENE_H
#define FLIPB200_SCENE_H
#include <vector>
#include "meshlevelset.h"
#include "trianglemesh.h"
#include "vmath.h"

struct FluidParticle {
    vmath::vec3 position;
    vmath::vec3 velocity;
    FluidParticle() {}
    FluidParticle(vmath::vec3 p) : position(p) {}
    FluidParticle(vmath::vec3 p, vmath::vec3 v) : position(p), velocity(v) {}
};

class FlipScene {
public:
    FlipScene() : ni(0), nj(0), nk(0), dx(0) {}
    void initialize(int i, int j, int k, float cellsize);       // domain box boundary
    void addBoundary(TriangleMesh &boundary, bool isInverted);   // union with a mesh SDF
    void resetBoundary();
    void addLiquid(TriangleMesh &mesh, std::vector<FluidParticle> &particles);  // libc rand() seeding

    int ni, nj, nk;
    float dx;
    MeshLevelSet solidSDF;
};

// precondition failures print and abort, like FLUIDSIM_ASSERT (src/fluidsimassert.h:24-37)
void flip_require(bool ok, const char *what);
#endif
