// Level-set fraction functions as device functions.  Same arithmetic (float, same operation
// order, no FMA contraction: the library is built with -fmad=false) as the reference's
// LevelsetUtils (src/levelsetutils.cpp:15-119 for the 2- and 4-point fractions,
// src/levelsetutils.cpp:189-235 + src/levelsetutils.h:39-77 for the tet/prism/cube volumes).
#pragma once
#include "rt.h"

// fraction of the segment (phiL, phiR) where phi < 0      (src/levelsetutils.cpp:15-27)
FLIP_HD float frac_inside2(float a, float b) {
    if (a < 0 && b < 0) return 1.0f;
    if (a < 0 && b >= 0) return a / (a - b);
    if (a >= 0 && b < 0) return b / (b - a);
    return 0.0f;
}

// fraction of the square (bl, br, tl, tr) where phi < 0, marching-squares cases
// (src/levelsetutils.cpp:38-119).  The reference rotates a 4-array {bl,br,tr,tl}; here the
// rotation is an index offset r, q(n) = element n of the rotated array.
FLIP_HD float frac_inside4(float bl, float br, float tl, float tr) {
    int inside = (bl < 0 ? 1 : 0) + (tl < 0 ? 1 : 0) + (br < 0 ? 1 : 0) + (tr < 0 ? 1 : 0);
    float L[4] = {bl, br, tr, tl};
#define Q(n) L[(r + (n)) & 3]
    if (inside == 4) return 1.0f;
    if (inside == 0) return 0.0f;
    int r = 0;
    if (inside == 3) {
        while (Q(0) < 0) r++;
        float s0 = 1 - frac_inside2(Q(0), Q(3));
        float s1 = 1 - frac_inside2(Q(0), Q(1));
        return 1.0f - 0.5f * s0 * s1;
    }
    if (inside == 2) {
        while (Q(0) >= 0 || !(Q(1) < 0 || Q(2) < 0)) r++;
        if (Q(1) < 0) {
            float sl = frac_inside2(Q(0), Q(3));
            float sr = frac_inside2(Q(1), Q(2));
            return 0.5f * (sl + sr);
        }
        float mid = 0.25f * (Q(0) + Q(1) + Q(2) + Q(3));
        if (mid < 0) {
            float area = 0;
            float s1 = 1 - frac_inside2(Q(0), Q(3));
            float s3 = 1 - frac_inside2(Q(2), Q(3));
            area += 0.5f * s1 * s3;
            float s2 = 1 - frac_inside2(Q(2), Q(1));
            float s0 = 1 - frac_inside2(Q(0), Q(1));
            area += 0.5f * s0 * s2;
            return 1.0f - area;
        } else {
            float area = 0;
            float s0 = frac_inside2(Q(0), Q(1));
            float s1 = frac_inside2(Q(0), Q(3));
            area += 0.5f * s0 * s1;
            float s2 = frac_inside2(Q(2), Q(1));
            float s3 = frac_inside2(Q(2), Q(3));
            area += 0.5f * s2 * s3;
            return area;
        }
    }
    // inside == 1
    while (Q(0) >= 0) r++;
    float s0 = frac_inside2(Q(0), Q(3));
    float s1 = frac_inside2(Q(0), Q(1));
    return 0.5f * s0 * s1;
#undef Q
}

FLIP_HD void cswap(float &a, float &b) {
    if (a > b) { float t = a; a = b; b = t; }
}

// volume fraction of a tetrahedron with corner values (a,b,c,d)    (src/levelsetutils.cpp:189-202)
FLIP_HD float tet_fraction(float a, float b, float c, float d) {
    // 5-comparator sorting network, src/levelsetutils.h:68-76
    cswap(a, b); cswap(c, d); cswap(a, c); cswap(b, d); cswap(b, c);
    if (d <= 0) return 1.0f;
    if (c <= 0) {
        // 1 - sortedTet(d, c, b, a)
        return 1 - d * d * d / ((d - c) * (d - b) * (d - a));
    }
    if (b <= 0) {
        // sortedPrism(a, b, c, d), src/levelsetutils.h:53-59
        float pa = a / (a - c);
        float pb = a / (a - d);
        float pc = b / (b - d);
        float pd = b / (b - c);
        return pa * pb * (1 - pd) + pb * (1 - pc) * pd + pc * pd;
    }
    if (a <= 0) return a * a * a / ((a - b) * (a - c) * (a - d));
    return 0.0f;
}

// cube -> average of the two 5-tet decompositions               (src/levelsetutils.cpp:219-235)
FLIP_HD float cube_fraction(float p000, float p100, float p010, float p110,
                            float p001, float p101, float p011, float p111) {
    return (tet_fraction(p000, p001, p101, p011) + tet_fraction(p000, p101, p100, p110) +
            tet_fraction(p000, p010, p011, p110) + tet_fraction(p101, p011, p111, p110) +
            2 * tet_fraction(p000, p011, p101, p110) + tet_fraction(p100, p101, p001, p111) +
            tet_fraction(p100, p001, p000, p010) + tet_fraction(p100, p110, p111, p010) +
            tet_fraction(p001, p111, p011, p010) + 2 * tet_fraction(p100, p111, p001, p010)) / 12.0f;
}
