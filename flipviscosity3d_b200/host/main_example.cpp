// Example driver: the reference's shipped scene (a Stanford-bunny-shaped mass of liquid dropped
// inside a spherical container, /root/reference/src/main.cpp:42-87) running on the GPU library.
// It only uses the public surface the reference's main.cpp uses, so that file itself also compiles
// against these headers unchanged; this one adds command-line control of size / frames / output.
//   usage: flip_example [N=64] [frames=300] [viscosity=5] [mesh_dir=sample_meshes] [ply|obj|none]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include "fluidsimulation.h"
#include "trianglemesh.h"

static void export_particles(int frame, std::vector<FluidParticle> &particles, const std::string &fmt) {
    if (fmt == "none") return;
    TriangleMesh mesh;
    mesh.vertices.reserve(particles.size());
    for (size_t i = 0; i < particles.size(); i++) mesh.vertices.push_back(particles[i].position);
    // NNNN.obj / NNNN.ply, 4-digit zero filled: the naming src/blender/render_particles.py expects
    std::ostringstream ss;
    ss << frame;
    std::string name = ss.str();
    if (name.size() < 4) name.insert(name.begin(), 4 - name.size(), '0');
    if (fmt == "obj") mesh.writeMeshToOBJ(name + ".obj");
    else mesh.writeMeshToPLY(name + ".ply");
    std::cout << "Exporting Particles to file: " << name << "." << fmt << std::endl;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 64;
    int frames = argc > 2 ? atoi(argv[2]) : 300;
    float viscosity = argc > 3 ? (float)atof(argv[3]) : 5.0f;
    std::string dir = argc > 4 ? argv[4] : "sample_meshes";
    std::string fmt = argc > 5 ? argv[5] : "ply";

    FluidSimulation fluidsim;
    float dx = 1.0f / n;
    fluidsim.initialize(n, n, n, dx);

    TriangleMesh boundaryMesh;
    if (!boundaryMesh.loadPLY(dir + "/sphere_large.ply")) {
        std::cout << "Error loading boundary mesh" << std::endl;
        return 1;
    }
    fluidsim.addBoundary(boundaryMesh, true);

    TriangleMesh liquidMesh;
    if (!liquidMesh.loadPLY(dir + "/stanford_bunny.ply")) {
        std::cout << "Error loading liquid mesh" << std::endl;
        return 1;
    }
    fluidsim.addLiquid(liquidMesh);
    fluidsim.setViscosity(viscosity);
    fluidsim.setGravity(0.0f, -9.81f, 0.0f);
    std::cout << fluidsim.particles.size() << " particles" << std::endl;

    float timestep = 0.01f;
    for (int frame = 0; frame < frames; frame++) {
        export_particles(frame, fluidsim.particles, fmt);
        fluidsim.advance(timestep);
        std::cout << "frame " << frame << ": " << fluidsim.lastSubsteps() << " substeps" << std::endl;
    }
    return 0;
}
