#include "trianglemesh.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {
bool header_count(const std::string &hdr, const char *key, int *n) {
    size_t p = hdr.find(key);
    if (p == std::string::npos) return false;
    return sscanf(hdr.c_str() + p + strlen(key), "%d", n) == 1;
}
}  // namespace

bool TriangleMesh::loadPLY(std::string filename) {
    std::ifstream f(filename.c_str(), std::ios::in | std::ios::binary);
    if (!f.is_open()) return false;
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const std::string endtag = "end_header\n";
    size_t he = data.find(endtag);
    if (he == std::string::npos) return false;
    he += endtag.size();
    std::string hdr = data.substr(0, he);
    if (hdr.find("binary_little_endian") == std::string::npos) return false;
    int nv = 0, nf = 0;
    if (!header_count(hdr, "element vertex ", &nv) || !header_count(hdr, "element face ", &nf)) return false;
    bool colors = hdr.find("property uchar red") != std::string::npos;
    size_t vstride = colors ? 15 : 12;
    if (data.size() < he + (size_t)nv * vstride + (size_t)nf * 13) return false;
    vertices.clear(); vertexcolors.clear(); triangles.clear();
    vertices.reserve(nv);
    const char *p = data.data() + he;
    for (int i = 0; i < nv; i++, p += vstride) {
        float v[3];
        memcpy(v, p, 12);
        vertices.push_back(vmath::vec3(v[0], v[1], v[2]));
        if (colors) {
            const unsigned char *c = (const unsigned char *)p + 12;
            vertexcolors.push_back(vmath::vec3(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f));
        }
    }
    triangles.reserve(nf);
    for (int i = 0; i < nf; i++, p += 13) {
        if ((unsigned char)p[0] != 3) return false;
        int t[3];
        memcpy(t, p + 1, 12);
        triangles.push_back(Triangle(t[0], t[1], t[2]));
    }
    return true;
}

void TriangleMesh::writeMeshToPLY(std::string filename) {
    bool colors = vertices.size() == vertexcolors.size();
    std::ostringstream h;
    h << "ply\nformat binary_little_endian 1.0\nelement vertex " << vertices.size()
      << "\nproperty float x\nproperty float y\nproperty float z\n";
    if (colors) h << "property uchar red\nproperty uchar green\nproperty uchar blue\n";
    h << "element face " << triangles.size() << "\nproperty list uchar int vertex_index\nend_header\n";
    std::string out = h.str();
    size_t vstride = colors ? 15 : 12;
    size_t off = out.size();
    out.resize(off + vertices.size() * vstride + triangles.size() * 13);
    char *p = &out[off];
    for (size_t i = 0; i < vertices.size(); i++, p += vstride) {
        float v[3] = {vertices[i].x, vertices[i].y, vertices[i].z};
        memcpy(p, v, 12);
        if (colors) {
            const vmath::vec3 &c = vertexcolors[i];
            unsigned char cc[3] = {(unsigned char)((c.x / 1.0) * 255.0), (unsigned char)((c.y / 1.0) * 255.0),
                                   (unsigned char)((c.z / 1.0) * 255.0)};
            memcpy(p + 12, cc, 3);
        }
    }
    for (size_t i = 0; i < triangles.size(); i++, p += 13) {
        p[0] = 0x03;
        memcpy(p + 1, triangles[i].tri, 12);
    }
    std::ofstream file(filename.c_str(), std::ios::out | std::ios::binary | std::ios::trunc);
    file.write(out.data(), (std::streamsize)out.size());
}

void TriangleMesh::writeMeshToOBJ(std::string filename) {
    std::ostringstream str;
    str << "# OBJ file format with ext .obj" << std::endl;
    str << "# vertex count = " << vertices.size() << std::endl;
    str << "# face count = " << triangles.size() << std::endl;
    for (size_t i = 0; i < vertices.size(); i++)
        str << "v " << vertices[i].x << " " << vertices[i].y << " " << vertices[i].z << std::endl;
    if (normals.size() == vertices.size())
        for (size_t i = 0; i < normals.size(); i++)
            str << "vn " << normals[i].x << " " << normals[i].y << " " << normals[i].z << std::endl;
    for (size_t i = 0; i < triangles.size(); i++) {
        int a = triangles[i].tri[0] + 1, b = triangles[i].tri[1] + 1, c = triangles[i].tri[2] + 1;
        str << "f " << a << "//" << a << " " << b << "//" << b << " " << c << "//" << c << std::endl;
    }
    std::ofstream out(filename.c_str());
    out << str.str();
}

void TriangleMesh::translate(vmath::vec3 t) {
    for (size_t i = 0; i < vertices.size(); i++) vertices[i] += t;
}
