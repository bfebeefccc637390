"""Host-side scene construction through libflip_host.so (C++11 FlipScene: the reference's
init-time algorithms — mesh SDF, boundary union, rand() seeding — see host/scene.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(_HERE, "lib", "libflip_host.so")
_lib = None


def host_library():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("flipviscosity3d_b200: %s not found; run __graft_entry__.build()" % HOST_LIB)
        _lib = C.CDLL(HOST_LIB)
        _lib.fliphost_scene_create.restype = C.c_void_p
        _lib.fliphost_scene_add_liquid.restype = C.c_longlong
    return _lib


def read_ply(path):
    """binary little-endian PLY (positions + triangles) -> (verts float32 [nv,3], tris int32 [nt,3])"""
    data = open(path, "rb").read()
    hend = data.index(b"end_header\n") + len(b"end_header\n")
    hdr = data[:hend].decode().split("\n")
    nv = int([l for l in hdr if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in hdr if l.startswith("element face")][0].split()[-1])
    v = np.frombuffer(data, np.float32, nv * 3, hend).reshape(nv, 3).copy()
    rec = np.dtype([("n", "u1"), ("i", "<i4", 3)])
    f = np.frombuffer(data, rec, nf, hend + nv * 12)
    return v, f["i"].astype(np.int32).copy()


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def mesh_sdf(ni, nj, nk, dx, verts, tris, band=3):
    verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.int32)
    out = np.empty((nk + 1, nj + 1, ni + 1), np.float32)
    host_library().fliphost_mesh_sdf(ni, nj, nk, C.c_float(dx), _fp(verts), len(verts), _ip(tris), len(tris), band, _fp(out))
    return out


class Scene:
    """FlipScene: domain-box boundary (+ optional meshes) and liquid seeding, no device needed."""

    def __init__(self, ni, nj, nk, dx, seed=1):
        self.ni, self.nj, self.nk, self.dx = ni, nj, nk, float(np.float32(dx))
        self.lib = host_library()
        self.lib.fliphost_srand(C.c_uint(seed))  # glibc's unseeded state == srand(1)
        self.h = C.c_void_p(self.lib.fliphost_scene_create(ni, nj, nk, C.c_float(dx)))
        self._particles = C.c_void_p()
        self.n_particles = 0

    def add_boundary(self, verts, tris, inverted=False):
        verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.int32)
        self.lib.fliphost_scene_add_boundary(self.h, _fp(verts), len(verts), _ip(tris), len(tris), int(inverted))

    def add_liquid(self, verts, tris):
        verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.int32)
        self.n_particles = int(self.lib.fliphost_scene_add_liquid(self.h, _fp(verts), len(verts), _ip(tris), len(tris),
                                                                   C.byref(self._particles)))

    def solid_sdf(self):
        out = np.empty((self.nk + 1, self.nj + 1, self.ni + 1), np.float32)
        self.lib.fliphost_scene_get_solid_sdf(self.h, _fp(out))
        return out

    def particles(self):
        out = np.zeros((self.n_particles, 6), np.float32)
        if self.n_particles:
            self.lib.fliphost_particles_get(self._particles, _fp(out))
        return out

    def close(self):
        if self.h:
            self.lib.fliphost_scene_destroy(self.h); self.h = None
        if self._particles:
            self.lib.fliphost_particles_free(self._particles); self._particles = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_points_ply(path, xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    host_library().fliphost_write_points_ply(path.encode(), _fp(xyz), C.c_longlong(len(xyz)))


def write_points_obj(path, xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    host_library().fliphost_write_points_obj(path.encode(), _fp(xyz), C.c_longlong(len(xyz)))
