// Host-side signed distance field of a closed triangle mesh on the (ni+1)(nj+1)(nk+1) node grid.
// Init-time only (the solid is static, the liquid mesh is only used for seeding); the result is
// uploaded once through flip_set_solid_sdf.  Same algorithm and float/double arithmetic as the
// reference's MeshLevelSet::calculateSignedDistanceField (/root/reference/src/meshlevelset.cpp:
// 138-347: exact distances in a band around each triangle, breadth-first propagation of the
// closest triangle, inside/outside from x-ray intersection parity), so scenes initialise to the
// same bits; written against flat std::vector storage instead of the reference's Array3d.
#ifndef FLIPB200_MESHLEVELSET_H
#define FLIPB200_MESHLEVELSET_H
#include <vector>
#include "trianglemesh.h"

class MeshLevelSet {
public:
    MeshLevelSet() : _ni(0), _nj(0), _nk(0), _dx(0.0) {}
    MeshLevelSet(int ni, int nj, int nk, double dx);

    void calculateSignedDistanceField(TriangleMesh &mesh, int bandwidth = 3);
    void calculateUnion(MeshLevelSet &other);   // pointwise min (src/meshlevelset.cpp:152-183)
    void negate();
    float trilinearInterpolate(vmath::vec3 pos);  // nodal sampling, out-of-range corners read 0
    float get(int i, int j, int k) const { return _phi[_at(i, j, k)]; }
    void getGridDimensions(int *i, int *j, int *k) const { *i = _ni; *j = _nj; *k = _nk; }
    const std::vector<float> &data() const { return _phi; }   // (ni+1)(nj+1)(nk+1), x fastest
    std::vector<float> &data() { return _phi; }

private:
    size_t _at(int i, int j, int k) const { return (size_t)i + (size_t)(_ni + 1) * ((size_t)j + (size_t)(_nj + 1) * (size_t)k); }
    int _ni, _nj, _nk;
    double _dx;
    std::vector<float> _phi;
    std::vector<int> _closest;
};

// trilinear sample of an (w,h,d) x-fastest float grid at node positions; corners outside read 0
// (Interpolation::trilinearInterpolate, src/interpolation.cpp:68-108)
float flip_host_trilinear(const float *grid, int w, int h, int d, double dx, vmath::vec3 p);
#endif
