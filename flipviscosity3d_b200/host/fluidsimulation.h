// Drop-in host class: the reference's FluidSimulation public surface
// (/root/reference/src/fluidsimulation.h:50-63) on top of the C ABI in include/flip_b200.h.
// The reference's main.cpp compiles against this header unchanged (include path swapped).
//
// What stays on the host (init time, reference algorithms so scenes start from the same bits):
// mesh -> SDF (meshlevelset.*), boundary union, particle seeding with libc rand().
// What runs on the GPU: everything inside advance().
#ifndef FLIPB200_FLUIDSIMULATION_H
#define FLIPB200_FLUIDSIMULATION_H

#include <vector>
#include "array3d.h"
#include "meshlevelset.h"
#include "scene.h"
#include "trianglemesh.h"
#include "vmath.h"

struct flip_sim;

class FluidSimulation {
public:
    FluidSimulation();
    ~FluidSimulation();

    void initialize(int i, int j, int k, float dx);
    void addBoundary(TriangleMesh &boundary, bool isInverted = false);
    void resetBoundary();
    void addLiquid(TriangleMesh &mesh);
    void setViscosity(float value);
    void setViscosity(Array3d<float> &vgrid);
    void setGravity(vmath::vec3 gravity);
    void setGravity(float gx, float gy, float gz);
    void advance(float dt);

    std::vector<FluidParticle> particles;   // host mirror, refreshed by advance()

    // --- additions (not in the reference) ---
    flip_sim *handle() { return _h; }
    void setVerbose(bool v);
    int lastSubsteps() const { return _lastSubsteps; }

private:
    FluidSimulation(const FluidSimulation &);
    FluidSimulation &operator=(const FluidSimulation &);
    void _uploadBoundary();

    flip_sim *_h;
    int _isize, _jsize, _ksize;
    float _dx;
    FlipScene _scene;
    bool _boundaryDirty;
    int _lastSubsteps;
};

#endif
