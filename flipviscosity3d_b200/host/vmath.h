// Minimal float vector type with the layout and operators the reference's public surface uses
// (vmath::vec3 in FluidParticle and setGravity; /root/reference/src/vmath.h).  12 bytes, no padding.
#ifndef FLIPB200_VMATH_H
#define FLIPB200_VMATH_H
#include <cmath>

namespace vmath {
struct vec3 {
    float x, y, z;
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    vec3(float xx, float yy, float zz) : x(xx), y(yy), z(zz) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(float s, const vec3 &v) { return vec3(v.x * s, v.y * s, v.z * s); }
inline vec3 operator*(const vec3 &v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
inline vec3 &operator+=(vec3 &a, const vec3 &b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline vec3 &operator-=(vec3 &a, const vec3 &b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float lengthsq(const vec3 &v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline float length(const vec3 &v) { return std::sqrt(lengthsq(v)); }
}  // namespace vmath
#endif
