#include "meshlevelset.h"
#include <cassert>
#include <cmath>

using vmath::vec3;

namespace {

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline double min3(double a, double b, double c) { return std::fmin(a, std::fmin(b, c)); }
inline double max3(double a, double b, double c) { return std::fmax(a, std::fmax(b, c)); }

// distance from x0 to the segment x1-x2 (src/meshlevelset.cpp:438-451)
float segment_distance(const vec3 &x0, const vec3 &x1, const vec3 &x2) {
    vec3 e = x2 - x1;
    double m2 = vmath::lengthsq(e);
    float s = (float)(vmath::dot(x2 - x0, e) / m2);
    if (s < 0) s = 0;
    else if (s > 1) s = 1;
    return vmath::length(x0 - (s * x1 + (1 - s) * x2));
}

// distance from x0 to the triangle x1-x2-x3 (src/meshlevelset.cpp:350-391)
float triangle_distance(const vec3 &x0, const vec3 &x1, const vec3 &x2, const vec3 &x3) {
    vec3 x13 = x1 - x3, x23 = x2 - x3, x03 = x0 - x3;
    float m13 = vmath::lengthsq(x13), m23 = vmath::lengthsq(x23);
    float d = vmath::dot(x13, x23);
    float invdet = 1.0f / std::fmax(m13 * m23 - d * d, 1e-30f);
    float a = vmath::dot(x13, x03), b = vmath::dot(x23, x03);
    float w23 = invdet * (m23 * a - d * b);
    float w31 = invdet * (m13 * b - d * a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) return vmath::length(x0 - (w23 * x1 + w31 * x2 + w12 * x3));
    if (w23 > 0) return std::fmin(segment_distance(x0, x1, x2), segment_distance(x0, x1, x3));
    if (w31 > 0) return std::fmin(segment_distance(x0, x1, x2), segment_distance(x0, x2, x3));
    return std::fmin(segment_distance(x0, x1, x3), segment_distance(x0, x2, x3));
}

// twice the signed area of (0,0)-(x1,y1)-(x2,y2) with simulation-of-simplicity tie breaking
// (src/meshlevelset.cpp:453-473)
int orientation(double x1, double y1, double x2, double y2, double *area2) {
    *area2 = y1 * x2 - x1 * y2;
    if (*area2 > 0) return 1;
    if (*area2 < 0) return -1;
    if (y2 > y1) return 1;
    if (y2 < y1) return -1;
    if (x1 > x2) return 1;
    if (x1 < x2) return -1;
    return 0;
}

// is (x0,y0) inside the 2-D triangle; barycentric coordinates out (src/meshlevelset.cpp:395-434)
bool barycentric(double x0, double y0, double x1, double y1, double x2, double y2, double x3, double y3,
                 double *a, double *b, double *c) {
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    double oa, ob, oc;
    int sa = orientation(x2, y2, x3, y3, &oa);
    if (sa == 0) return false;
    if (orientation(x3, y3, x1, y1, &ob) != sa) return false;
    if (orientation(x1, y1, x2, y2, &oc) != sa) return false;
    double sum = oa + ob + oc;
    assert(sum != 0);
    double inv = 1.0 / sum;
    *a = oa * inv; *b = ob * inv; *c = oc * inv;
    return true;
}

}  // namespace

float flip_host_trilinear(const float *grid, int w, int h, int d, double dx, vec3 p) {
    double invdx = 1.0 / dx;
    int gi = (int)std::floor(p.x * invdx), gj = (int)std::floor(p.y * invdx), gk = (int)std::floor(p.z * invdx);
    float gx = (float)(gi * dx), gy = (float)(gj * dx), gz = (float)(gk * dx);
    double x = (p.x - gx) * invdx, y = (p.y - gy) * invdx, z = (p.z - gz) * invdx;
    double c[8];
    const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 0}, {1, 1, 1}};
    for (int n = 0; n < 8; n++) {
        int i = gi + off[n][0], j = gj + off[n][1], k = gk + off[n][2];
        c[n] = (i >= 0 && j >= 0 && k >= 0 && i < w && j < h && k < d)
                   ? (double)grid[(size_t)i + (size_t)w * ((size_t)j + (size_t)h * (size_t)k)] : 0.0;
    }
    return (float)(c[0] * (1 - x) * (1 - y) * (1 - z) + c[1] * x * (1 - y) * (1 - z) + c[2] * (1 - x) * y * (1 - z) +
                   c[3] * (1 - x) * (1 - y) * z + c[4] * x * (1 - y) * z + c[5] * (1 - x) * y * z +
                   c[6] * x * y * (1 - z) + c[7] * x * y * z);
}

MeshLevelSet::MeshLevelSet(int ni, int nj, int nk, double dx)
    : _ni(ni), _nj(nj), _nk(nk), _dx(dx), _phi((size_t)(ni + 1) * (nj + 1) * (nk + 1), 0.0f),
      _closest((size_t)(ni + 1) * (nj + 1) * (nk + 1), -1) {}

float MeshLevelSet::trilinearInterpolate(vec3 pos) {
    return flip_host_trilinear(_phi.data(), _ni + 1, _nj + 1, _nk + 1, _dx, pos);
}

void MeshLevelSet::negate() {
    for (size_t i = 0; i < _phi.size(); i++) _phi[i] = -_phi[i];
}

void MeshLevelSet::calculateUnion(MeshLevelSet &o) {
    assert(o._ni == _ni && o._nj == _nj && o._nk == _nk);
    for (size_t i = 0; i < _phi.size(); i++) {
        if (o._phi[i] < _phi[i]) _phi[i] = o._phi[i];
    }
}

void MeshLevelSet::calculateSignedDistanceField(TriangleMesh &mesh, int band) {
    const int W = _ni + 1, H = _nj + 1, D = _nk + 1;
    const double invdx = 1.0 / _dx;
    std::vector<int> crossings((size_t)W * H * D, 0);
    std::fill(_phi.begin(), _phi.end(), (float)((W + H + D) * _dx));
    std::fill(_closest.begin(), _closest.end(), -1);

    // 1. exact distances near each triangle + x-ray crossing counts
    for (size_t t = 0; t < mesh.triangles.size(); t++) {
        const Triangle &tr = mesh.triangles[t];
        vec3 p = mesh.vertices[tr.tri[0]], q = mesh.vertices[tr.tri[1]], r = mesh.vertices[tr.tri[2]];
        double fip = (double)p.x * invdx, fjp = (double)p.y * invdx, fkp = (double)p.z * invdx;
        double fiq = (double)q.x * invdx, fjq = (double)q.y * invdx, fkq = (double)q.z * invdx;
        double fir = (double)r.x * invdx, fjr = (double)r.y * invdx, fkr = (double)r.z * invdx;
        int i0 = clampi(int(min3(fip, fiq, fir)) - band, 0, W - 1), i1 = clampi(int(max3(fip, fiq, fir)) + band + 1, 0, W - 1);
        int j0 = clampi(int(min3(fjp, fjq, fjr)) - band, 0, H - 1), j1 = clampi(int(max3(fjp, fjq, fjr)) + band + 1, 0, H - 1);
        int k0 = clampi(int(min3(fkp, fkq, fkr)) - band, 0, D - 1), k1 = clampi(int(max3(fkp, fkq, fkr)) + band + 1, 0, D - 1);
        for (int k = k0; k <= k1; k++)
            for (int j = j0; j <= j1; j++)
                for (int i = i0; i <= i1; i++) {
                    vec3 gp((float)(i * _dx), (float)(j * _dx), (float)(k * _dx));
                    float dist = triangle_distance(gp, p, q, r);
                    size_t id = _at(i, j, k);
                    if (dist < _phi[id]) { _phi[id] = dist; _closest[id] = (int)t; }
                }
        j0 = clampi((int)std::ceil(min3(fjp, fjq, fjr)), 0, H - 1);
        k0 = clampi((int)std::ceil(min3(fkp, fkq, fkr)), 0, D - 1);
        j1 = clampi((int)std::floor(max3(fjp, fjq, fjr)), 0, H - 1);
        k1 = clampi((int)std::floor(max3(fkp, fkq, fkr)), 0, D - 1);
        for (int k = k0; k <= k1; k++)
            for (int j = j0; j <= j1; j++) {
                double a, b, c;
                if (!barycentric(j, k, fjp, fkp, fjq, fkq, fjr, fkr, &a, &b, &c)) continue;
                double fi = a * fip + b * fiq + c * fir;
                int cell = int(std::ceil(fi));
                if (cell < 0) crossings[_at(0, j, k)] += 1;
                else if (cell < W) crossings[_at(cell, j, k)] += 1;
            }
    }

    // 2. breadth-first propagation of the closest triangle to the rest of the grid
    std::vector<int> queue;  // flat node ids in visiting order
    queue.reserve((size_t)W * H * D);
    std::vector<char> seen((size_t)W * H * D, 0);
    for (size_t id = 0; id < _closest.size(); id++)
        if (_closest[id] != -1) { seen[id] = 1; queue.push_back((int)id); }
    const size_t first_unknown = queue.size();
    const int di[6] = {-1, 1, 0, 0, 0, 0}, dj[6] = {0, 0, -1, 1, 0, 0}, dk[6] = {0, 0, 0, 0, -1, 1};
    for (size_t head = 0; head < queue.size(); head++) {
        int id = queue[head];
        int i = id % W, j = (id / W) % H, k = id / (W * H);
        for (int n = 0; n < 6; n++) {
            int a = i + di[n], b = j + dj[n], c = k + dk[n];
            if (a < 0 || b < 0 || c < 0 || a >= W || b >= H || c >= D) continue;
            size_t nid = _at(a, b, c);
            if (!seen[nid]) { seen[nid] = 1; queue.push_back((int)nid); }
        }
    }
    for (size_t head = first_unknown; head < queue.size(); head++) {
        int id = queue[head];
        int i = id % W, j = (id / W) % H, k = id / (W * H);
        vec3 gp((float)(i * _dx), (float)(j * _dx), (float)(k * _dx));
        for (int n = 0; n < 6; n++) {
            int a = i + di[n], b = j + dj[n], c = k + dk[n];
            if (a < 0 || b < 0 || c < 0 || a >= W || b >= H || c >= D) continue;
            int tn = _closest[_at(a, b, c)];
            if (tn == -1) continue;
            const Triangle &tr = mesh.triangles[tn];
            double dist = triangle_distance(gp, mesh.vertices[tr.tri[0]], mesh.vertices[tr.tri[1]], mesh.vertices[tr.tri[2]]);
            if (dist < _phi[id]) { _phi[id] = (float)dist; _closest[id] = tn; }
        }
    }

    // 3. signs from crossing parity along +x
    for (int k = 0; k < D; k++)
        for (int j = 0; j < H; j++) {
            int total = 0;
            for (int i = 0; i < W; i++) {
                size_t id = _at(i, j, k);
                total += crossings[id];
                if (total % 2 == 1) _phi[id] = -_phi[id];
            }
        }
}
