// Test driver for the C++ drop-in class (flipviscosity3d_b200/host/fluidsimulation.h): the calls the
// reference's main.cpp makes (/root/reference/src/main.cpp:42-87), plus resetBoundary(), at a small size.
//   usage: dropin_frames N mesh_dir out.bin [frames_before_reset=2] [frames_after_reset=1]
// out.bin = int64 particle count, then count * 6 float32 {pos, vel}.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <stdint.h>
#include "fluidsimulation.h"
#include "trianglemesh.h"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    int n = atoi(argv[1]);
    std::string dir = argv[2];
    int before = argc > 4 ? atoi(argv[4]) : 2, after = argc > 5 ? atoi(argv[5]) : 1;
    FluidSimulation fluidsim;
    fluidsim.initialize(n, n, n, 1.0f / n);
    TriangleMesh boundaryMesh, liquidMesh;
    if (!boundaryMesh.loadPLY(dir + "/sphere_large.ply") || !liquidMesh.loadPLY(dir + "/stanford_bunny.ply")) return 3;
    fluidsim.addBoundary(boundaryMesh, true);
    fluidsim.addLiquid(liquidMesh);
    fluidsim.setViscosity(5.0f);
    fluidsim.setGravity(0.0f, -9.81f, 0.0f);
    for (int f = 0; f < before; f++) fluidsim.advance(0.01f);
    // back to the bare domain box: the solid-derived fields (face weights, face states) must follow
    fluidsim.resetBoundary();
    for (int f = 0; f < after; f++) fluidsim.advance(0.01f);
    FILE *fp = fopen(argv[3], "wb");
    if (!fp) return 4;
    int64_t cnt = (int64_t)fluidsim.particles.size();
    fwrite(&cnt, sizeof(cnt), 1, fp);
    fwrite(fluidsim.particles.empty() ? 0 : &fluidsim.particles[0].position.x, 6 * sizeof(float), (size_t)cnt, fp);
    fclose(fp);
    std::cout << "DROPIN_OK " << cnt << std::endl;
    return 0;
}
