"""ORACLE wrapper (test infrastructure, NOT product code).

ctypes binding of oracle/_ref/libflipref.so = the unmodified reference sources
(/root/reference/src, commit 178b82f) + oracle/ref_harness.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under flipviscosity3d_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libflipref.so")
REF_SRC = os.environ.get("FLIP_REFERENCE_DIR", "/root/reference")


def build(force=False):
    """Compile the reference + harness if the sources are present (this container).
    On the GPU box only the prebuilt .so travels; returns False if neither exists."""
    if os.path.isdir(os.path.join(REF_SRC, "src")):
        if force or not os.path.exists(LIB_PATH) or (
                os.path.getmtime(os.path.join(_HERE, "ref_harness.cpp")) > os.path.getmtime(LIB_PATH)):
            subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "REF=" + REF_SRC])
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/libflipref.so missing: run `make -C oracle` where "
                               "/root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_num_particles.restype = C.c_longlong
        _lib.ref_cfl.restype = C.c_float
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def mesh_sdf(ni, nj, nk, dx, verts, tris, band=3):
    verts = np.ascontiguousarray(verts, np.float32)
    tris = np.ascontiguousarray(tris, np.int32)
    out = np.empty((nk + 1, nj + 1, ni + 1), np.float32)
    lib().ref_mesh_sdf(ni, nj, nk, C.c_float(dx), _fp(verts), len(verts), _ip(tris), len(tris), band, _fp(out))
    return out


def srand(seed=1):
    lib().ref_srand(C.c_uint(seed))


class RefSim:
    """The reference FluidSimulation, stage by stage.  Grids are numpy arrays shaped
    (depth, height, width) = (k, j, i), x fastest, like the reference's Array3d."""

    def __init__(self, ni, nj, nk, dx):
        self.ni, self.nj, self.nk, self.dx = ni, nj, nk, float(np.float32(dx))
        self.h = C.c_void_p(lib().ref_create(ni, nj, nk, C.c_float(dx)))

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # shapes
    def shape_u(self):
        return (self.nk, self.nj, self.ni + 1)

    def shape_v(self):
        return (self.nk, self.nj + 1, self.ni)

    def shape_w(self):
        return (self.nk + 1, self.nj, self.ni)

    def shape_c(self):
        return (self.nk, self.nj, self.ni)

    def shape_n(self):
        return (self.nk + 1, self.nj + 1, self.ni + 1)

    # scene
    def add_boundary(self, verts, tris, inverted=False):
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        lib().ref_add_boundary(self.h, _fp(verts), len(verts), _ip(tris), len(tris), int(inverted))

    def add_liquid(self, verts, tris):
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        lib().ref_add_liquid(self.h, _fp(verts), len(verts), _ip(tris), len(tris))

    def set_viscosity(self, v):
        if np.isscalar(v):
            lib().ref_set_viscosity(self.h, C.c_float(v))
        else:
            v = np.ascontiguousarray(v, np.float32)
            assert v.shape == self.shape_n()
            lib().ref_set_viscosity_grid(self.h, _fp(v))

    def set_gravity(self, gx, gy, gz):
        lib().ref_set_gravity(self.h, C.c_float(gx), C.c_float(gy), C.c_float(gz))

    # particles
    def num_particles(self):
        return int(lib().ref_num_particles(self.h))

    def get_particles(self):
        out = np.empty((self.num_particles(), 6), np.float32)
        lib().ref_get_particles(self.h, _fp(out))
        return out

    def set_particles(self, p):
        p = np.ascontiguousarray(p, np.float32)
        lib().ref_set_particles(self.h, _fp(p), C.c_longlong(len(p)))

    # grids
    def _get3(self, fn, dtype=np.float32):
        u = np.empty(self.shape_u(), dtype)
        v = np.empty(self.shape_v(), dtype)
        w = np.empty(self.shape_w(), dtype)
        cast = _fp if dtype == np.float32 else _bp
        fn(self.h, cast(u), cast(v), cast(w))
        return u, v, w

    def _set3(self, fn, u, v, w, dtype=np.float32):
        u = np.ascontiguousarray(u, dtype); v = np.ascontiguousarray(v, dtype); w = np.ascontiguousarray(w, dtype)
        assert u.shape == self.shape_u() and v.shape == self.shape_v() and w.shape == self.shape_w()
        cast = _fp if dtype == np.float32 else _bp
        fn(self.h, cast(u), cast(v), cast(w))

    def get_mac(self):
        return self._get3(lib().ref_get_mac)

    def set_mac(self, u, v, w):
        self._set3(lib().ref_set_mac, u, v, w)

    def get_saved_mac(self):
        return self._get3(lib().ref_get_saved_mac)

    def set_saved_mac(self, u, v, w):
        self._set3(lib().ref_set_saved_mac, u, v, w)

    def get_weights(self):
        return self._get3(lib().ref_get_weights)

    def get_valid(self):
        return self._get3(lib().ref_get_valid, np.uint8)

    def set_valid(self, u, v, w):
        self._set3(lib().ref_set_valid, u, v, w, np.uint8)

    def get_solid_sdf(self):
        out = np.empty(self.shape_n(), np.float32)
        lib().ref_get_solid_sdf(self.h, _fp(out))
        return out

    def set_solid_sdf(self, phi):
        phi = np.ascontiguousarray(phi, np.float32)
        assert phi.shape == self.shape_n()
        lib().ref_set_solid_sdf(self.h, _fp(phi))

    def get_liquid_sdf(self):
        out = np.empty(self.shape_c(), np.float32)
        lib().ref_get_liquid_sdf(self.h, _fp(out))
        return out

    def set_liquid_sdf(self, phi):
        phi = np.ascontiguousarray(phi, np.float32)
        assert phi.shape == self.shape_c()
        lib().ref_set_liquid_sdf(self.h, _fp(phi))

    # stages
    def cfl(self):
        return float(lib().ref_cfl(self.h))

    def update_liquid_sdf(self):
        lib().ref_stage_update_liquid_sdf(self.h)

    def p2g_component(self, d):
        shp = (self.shape_u(), self.shape_v(), self.shape_w())[d]
        f = np.empty(shp, np.float32)
        s = np.empty(shp, np.uint8)
        lib().ref_p2g_component(self.h, d, _fp(f), _bp(s))
        return f, s

    def advect_velocity_field(self):
        lib().ref_stage_advect_velocity_field(self.h)

    def add_body_force(self, dt):
        lib().ref_stage_add_body_force(self.h, C.c_float(dt))

    def extrapolate(self):
        lib().ref_extrapolate(self.h)

    def viscosity_volumes(self):
        ni, nj, nk = self.ni, self.nj, self.nk
        shapes = [(nk, nj, ni), (nk, nj, ni + 1), (nk, nj + 1, ni), (nk + 1, nj, ni),
                  (nk + 1, nj + 1, ni), (nk + 1, nj, ni + 1), (nk, nj + 1, ni + 1)]
        outs = [np.empty(s, np.float32) for s in shapes]
        lib().ref_viscosity_volumes(self.h, *[_fp(o) for o in outs])
        return outs  # center, U, V, W, edgeU, edgeV, edgeW

    def apply_viscosity(self, dt, tol=0.0, maxit=0):
        wrote = lib().ref_stage_apply_viscosity_ex(self.h, C.c_float(dt), C.c_double(tol), int(maxit))
        it = C.c_int(); res = C.c_double(); ok = C.c_int(); unk = C.c_int()
        lib().ref_viscosity_diag(self.h, C.byref(it), C.byref(res), C.byref(ok), C.byref(unk))
        return dict(wrote=int(wrote), iters=it.value, resid=res.value, ok=ok.value, unknowns=unk.value)

    def viscosity_residual(self, dt, u, v, w):
        """max|b - A x| of candidate MAC fields in the reference's own assembled viscosity system (state unchanged)"""
        u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32); w = np.ascontiguousarray(w, np.float32)
        assert u.shape == self.shape_u() and v.shape == self.shape_v() and w.shape == self.shape_w()
        out = (C.c_double * 6)()
        lib().ref_viscosity_residual(self.h, C.c_float(dt), _fp(u), _fp(v), _fp(w), out)
        return dict(resid=out[0], bmax=out[1], unknowns=int(out[2]), xmax=out[3], nonpositive_diagonals=int(out[4]),
                    resid_faces_with_volume=out[5])

    def compute_weights(self):
        lib().ref_compute_weights(self.h)

    def solve_pressure(self, dt, tol=0.0, maxit=0):
        out = np.empty(self.shape_c(), np.float32)
        lib().ref_solve_pressure(self.h, C.c_float(dt), C.c_double(tol), int(maxit), _fp(out))
        return out

    def apply_pressure(self, dt, p):
        p = np.ascontiguousarray(p, np.float32)
        lib().ref_apply_pressure(self.h, C.c_float(dt), _fp(p))

    def project(self, dt):
        lib().ref_stage_project(self.h, C.c_float(dt))

    def constrain(self):
        lib().ref_stage_constrain(self.h)

    def advect_particles(self, dt):
        lib().ref_stage_advect_particles(self.h, C.c_float(dt))

    def substep(self, dt):
        t = np.zeros(8, np.float64)
        lib().ref_substep(self.h, C.c_float(dt), t.ctypes.data_as(C.POINTER(C.c_double)))
        return t

    def advance(self, dt):
        return int(lib().ref_advance(self.h, C.c_float(dt)))

    def write_particles_ply(self, path):
        lib().ref_write_particles_ply(self.h, path.encode())

    def write_particles_obj(self, path):
        lib().ref_write_particles_obj(self.h, path.encode())

    def reset_boundary(self):
        lib().ref_reset_boundary(self.h)
