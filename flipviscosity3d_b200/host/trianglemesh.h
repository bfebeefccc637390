// Triangle mesh container + PLY/OBJ I/O with the reference's public members and file formats
// (/root/reference/src/trianglemesh.h:36-58).  The PLY reader parses the whole file, so it also
// loads the small sample meshes the reference's own reader rejects (files < 2048 bytes,
// src/trianglemesh.cpp:426-444); writers are byte-compatible with src/trianglemesh.cpp:190-343
// (binary little-endian PLY) and :381-418 (OBJ text).
#ifndef FLIPB200_TRIANGLEMESH_H
#define FLIPB200_TRIANGLEMESH_H
#include <string>
#include <vector>
#include "vmath.h"

struct Triangle {
    int tri[3];
    Triangle() { tri[0] = tri[1] = tri[2] = 0; }
    Triangle(int a, int b, int c) { tri[0] = a; tri[1] = b; tri[2] = c; }
};

class TriangleMesh {
public:
    bool loadPLY(std::string filename);
    void writeMeshToPLY(std::string filename);
    void writeMeshToOBJ(std::string filename);
    int numVertices() { return (int)vertices.size(); }
    int numFaces() { return (int)triangles.size(); }
    int numTriangles() { return numFaces(); }
    void translate(vmath::vec3 t);

    std::vector<vmath::vec3> vertices;
    std::vector<vmath::vec3> vertexcolors;
    std::vector<vmath::vec3> normals;
    std::vector<Triangle> triangles;
};
#endif
