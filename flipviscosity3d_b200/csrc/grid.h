// Device grid layout and the float/double index conventions of the reference.
//
// Every field (cell-centred, U/V/W faces, nodes, edges) lives in ONE padded dense layout:
//   idx(i,j,k) = (i + PX) + AX * ((j + PY) + AY * (k + PZ)),   x fastest
// with AX a multiple of 8 and PX = 8 so that 8-wide x-rows of the 8x8x8 solver blocks start on
// 32-byte (fp32) / 64-byte (fp64) boundaries, and one zero ghost layer around j and k.  The
// reference's Array3d layout (i + W*(j + H*k), src/array3d.h:397-400) is only the wire format at
// the C ABI; conversion happens in api.cu.
#pragma once
#include "rt.h"
#include <cmath>

#define FLIP_PX 8
#define FLIP_PY 1
#define FLIP_PZ 1
#define FLIP_B 8  // solver block edge (8x8x8 cells)

struct Grid {
    int ni, nj, nk;      // cells
    int ax, ay, az;      // allocated extents
    int total;           // ax*ay*az
    int nbx, nby, nbz;   // 8^3 blocks covering [0,ni] x [0,nj] x [0,nk]
    int nblocks;
    float dx;            // as the reference's float _dx
    double dxd;          // (double)dx
    double invdx;        // 1.0 / (double)dx      (src/grid3d.h:60-65)
    float hdx;           // (float)(0.5*dx)
};

static inline Grid make_grid(int ni, int nj, int nk, float dx) {
    Grid g;
    g.ni = ni; g.nj = nj; g.nk = nk;
    g.ax = FLIP_PX + ((ni + 2 + 7) / 8) * 8;
    g.ay = nj + 1 + 2 * FLIP_PY;
    g.az = nk + 1 + 2 * FLIP_PZ;
    g.total = g.ax * g.ay * g.az;
    g.nbx = (ni + 1 + FLIP_B - 1) / FLIP_B;
    g.nby = (nj + 1 + FLIP_B - 1) / FLIP_B;
    g.nbz = (nk + 1 + FLIP_B - 1) / FLIP_B;
    g.nblocks = g.nbx * g.nby * g.nbz;
    g.dx = dx;
    g.dxd = (double)dx;
    g.invdx = 1.0 / (double)dx;
    g.hdx = (float)(0.5 * (double)dx);
    return g;
}

FLIP_HD int gidx(const Grid &g, int i, int j, int k) {
    return (i + FLIP_PX) + g.ax * ((j + FLIP_PY) + g.ay * (k + FLIP_PZ));
}
#define SY(g) ((g).ax)
#define SZ(g) ((g).ax * (g).ay)

// linear thread id -> (i,j,k) over an (nx,ny,nz) box; returns false when out of the box
FLIP_HD bool unflatten(long long t, int nx, int ny, int nz, int &i, int &j, int &k) {
    if (t >= (long long)nx * ny * nz) return false;
    i = (int)(t % nx);
    long long r = t / nx;
    j = (int)(r % ny);
    k = (int)(r / ny);
    return true;
}

// floor(p * invdx) with p promoted from float: Grid3d::positionToGridIndex(vec3, double dx)
FLIP_HD int pos_to_index(float p, double invdx) { return (int)floor((double)p * invdx); }
// (float)(i*dx): Grid3d::GridIndexToPosition (src/grid3d.h:81-83)
FLIP_HD float index_to_pos(int i, double dxd) { return (float)((double)i * dxd); }
// (float)(i*dx + 0.5*dx): Grid3d::GridIndexToCellCenter(int,int,int,double) (src/grid3d.h:104-107)
FLIP_HD float index_to_center(int i, double dxd) { return (float)((double)i * dxd + 0.5 * dxd); }

// Grid3d::isFaceBorderingValue{U,V,W} with the fluid mask = (phi < 0) on the ni*nj*nk cell grid
// (src/grid3d.h:496-530; mask built at src/fluidsimulation.cpp:501-510).  dir: 0=U,1=V,2=W.
FLIP_D bool face_borders_fluid(const Grid &g, const float *__restrict__ phi, int dir, int i, int j, int k) {
    int n = dir == 0 ? g.ni : (dir == 1 ? g.nj : g.nk);
    int c = dir == 0 ? i : (dir == 1 ? j : k);
    int st = dir == 0 ? 1 : (dir == 1 ? SY(g) : SZ(g));
    int id = gidx(g, i, j, k);
    if (c == n) return phi[id - st] < 0.0f;
    if (c > 0) return phi[id] < 0.0f || phi[id - st] < 0.0f;
    return phi[id] < 0.0f;
}

// Interpolation::trilinearInterpolate(vec3 p, double dx, Array3d<float>&) (src/interpolation.cpp:68-108):
// float (p - gpos), double weights, out-of-range corners read as 0.  (w,h,d) are the logical dims.
FLIP_D double trilinear_grid(const Grid &g, const float *__restrict__ f, int w, int h, int d,
                             float px, float py, float pz) {
    int gi = pos_to_index(px, g.invdx), gj = pos_to_index(py, g.invdx), gk = pos_to_index(pz, g.invdx);
    float gx = index_to_pos(gi, g.dxd), gy = index_to_pos(gj, g.dxd), gz = index_to_pos(gk, g.dxd);
    double x = (double)(px - gx) * g.invdx;
    double y = (double)(py - gy) * g.invdx;
    double z = (double)(pz - gz) * g.invdx;
    double p[8];
#define FLIP_CORNER(n, a, b, c)                                                            \
    {                                                                                      \
        int ii = gi + a, jj = gj + b, kk = gk + c;                                         \
        p[n] = (ii >= 0 && jj >= 0 && kk >= 0 && ii < w && jj < h && kk < d)               \
                   ? (double)f[gidx(g, ii, jj, kk)] : 0.0;                                 \
    }
    FLIP_CORNER(0, 0, 0, 0) FLIP_CORNER(1, 1, 0, 0) FLIP_CORNER(2, 0, 1, 0) FLIP_CORNER(3, 0, 0, 1)
    FLIP_CORNER(4, 1, 0, 1) FLIP_CORNER(5, 0, 1, 1) FLIP_CORNER(6, 1, 1, 0) FLIP_CORNER(7, 1, 1, 1)
#undef FLIP_CORNER
    // Interpolation::trilinearInterpolate(double[8], x, y, z) (src/interpolation.cpp:57-66)
    return p[0] * (1 - x) * (1 - y) * (1 - z) + p[1] * x * (1 - y) * (1 - z) + p[2] * (1 - x) * y * (1 - z) +
           p[3] * (1 - x) * (1 - y) * z + p[4] * x * (1 - y) * z + p[5] * (1 - x) * y * z +
           p[6] * x * y * (1 - z) + p[7] * x * y * z;
}
