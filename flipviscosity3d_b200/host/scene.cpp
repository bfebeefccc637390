#include "scene.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

void flip_require(bool ok, const char *what) {
    if (!ok) {
        fprintf(stderr, "FluidSimulation: precondition failed: %s\n", what);
        abort();
    }
}

namespace {
// AABB(points) min / max corner as the reference computes them (src/aabb.cpp:71-108, 205-211)
void mesh_bounds(const std::vector<vmath::vec3> &pts, vmath::vec3 *mn, vmath::vec3 *mx) {
    double lo[3] = {pts[0].x, pts[0].y, pts[0].z}, hi[3] = {pts[0].x, pts[0].y, pts[0].z};
    for (size_t i = 0; i < pts.size(); i++) {
        const float c[3] = {pts[i].x, pts[i].y, pts[i].z};
        for (int a = 0; a < 3; a++) { lo[a] = std::fmin((double)c[a], lo[a]); hi[a] = std::fmax((double)c[a], hi[a]); }
    }
    const double eps = 1e-9;
    *mn = vmath::vec3((float)lo[0], (float)lo[1], (float)lo[2]);
    *mx = vmath::vec3(mn->x + (float)(hi[0] - lo[0] + eps), mn->y + (float)(hi[1] - lo[1] + eps), mn->z + (float)(hi[2] - lo[2] + eps));
}

bool inside_domain(vmath::vec3 p, int ni, int nj, int nk, float dx) {
    // AABB domain(0,0,0, ni*dx, ...).isPointInside (src/aabb.cpp:118-121)
    double w = ni * dx, h = nj * dx, d = nk * dx;
    return p.x >= 0.0f && p.y >= 0.0f && p.z >= 0.0f && p.x < 0.0f + w && p.y < 0.0f + h && p.z < 0.0f + d;
}

}  // namespace

void FlipScene::initialize(int i, int j, int k, float cellsize) {
    ni = i; nj = j; nk = k; dx = cellsize;
    resetBoundary();
}

// domain box inset by 3dx + 1e-6, as a closed mesh, turned inside out
// (src/fluidsimulation.cpp:198-239)
void FlipScene::resetBoundary() {
    double eps = 1e-6;
    double v = -3 * dx - eps;
    double half = 0.5 * v;
    float px = 0.0f - (float)half, py = 0.0f - (float)half, pz = 0.0f - (float)half;
    double w = (double)(ni * dx) + v, h = (double)(nj * dx) + v, d = (double)(nk * dx) + v;
    float x1 = px + (float)w, y1 = py + (float)h, z1 = pz + (float)d;
    TriangleMesh box;
    const float vx[8][3] = {{px, py, pz}, {x1, py, pz}, {x1, py, z1}, {px, py, z1},
                            {px, y1, pz}, {x1, y1, pz}, {x1, y1, z1}, {px, y1, z1}};
    for (int n = 0; n < 8; n++) box.vertices.push_back(vmath::vec3(vx[n][0], vx[n][1], vx[n][2]));
    const int tr[12][3] = {{0, 1, 2}, {0, 2, 3}, {4, 7, 6}, {4, 6, 5}, {0, 3, 7}, {0, 7, 4},
                           {1, 5, 6}, {1, 6, 2}, {0, 4, 5}, {0, 5, 1}, {3, 2, 6}, {3, 6, 7}};
    for (int n = 0; n < 12; n++) box.triangles.push_back(Triangle(tr[n][0], tr[n][1], tr[n][2]));
    solidSDF = MeshLevelSet(ni, nj, nk, dx);
    solidSDF.calculateSignedDistanceField(box, 3);
    solidSDF.negate();
}


void FlipScene::addBoundary(TriangleMesh &boundary, bool isInverted) {
    flip_require(!boundary.vertices.empty(), "boundary mesh is empty");
    vmath::vec3 mn, mx;
    mesh_bounds(boundary.vertices, &mn, &mx);
    flip_require(inside_domain(mn, ni, nj, nk, dx) && inside_domain(mx, ni, nj, nk, dx),
            "boundary mesh must lie inside the simulation domain");
    MeshLevelSet sdf(ni, nj, nk, dx);
    sdf.calculateSignedDistanceField(boundary, 3);
    if (isInverted) sdf.negate();
    solidSDF.calculateUnion(sdf);
}


// 8 jittered candidates per cell in k,j,i order, libc rand() never seeded
// (src/fluidsimulation.cpp:64-97, src/fluidsimulation.h:100-102)
void FlipScene::addLiquid(TriangleMesh &mesh, std::vector<FluidParticle> &particles) {
    flip_require(!mesh.vertices.empty(), "liquid mesh is empty");
    vmath::vec3 mn, mx;
    mesh_bounds(mesh.vertices, &mn, &mx);
    flip_require(inside_domain(mn, ni, nj, nk, dx) && inside_domain(mx, ni, nj, nk, dx),
            "liquid mesh must lie inside the simulation domain");
    MeshLevelSet sdf(ni, nj, nk, dx);
    sdf.calculateSignedDistanceField(mesh, 3);
    const double lo = 0.0, hi = dx;
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                vmath::vec3 corner((float)(i * (double)dx), (float)(j * (double)dx), (float)(k * (double)dx));
                for (int n = 0; n < 8; n++) {
                    float a = (float)(lo + (double)rand() / ((double)RAND_MAX / (hi - lo)));
                    float b = (float)(lo + (double)rand() / ((double)RAND_MAX / (hi - lo)));
                    float c = (float)(lo + (double)rand() / ((double)RAND_MAX / (hi - lo)));
                    vmath::vec3 pos = corner + vmath::vec3(a, b, c);
                    if (sdf.trilinearInterpolate(pos) < 0.0) {
                        if (solidSDF.trilinearInterpolate(pos) >= 0) particles.push_back(FluidParticle(pos));
                    }
                }
            }
}

