"""Python host-side mirror of the C ABI: FlipSim wraps one flip_sim handle.

Grids are numpy float32 arrays shaped (depth, height, width) = (k, j, i), x fastest: the
reference's Array3d layout (src/array3d.h:397-400).  Particles are (n, 6) float32 AoS
{pos.xyz, vel.xyz} like the reference's FluidParticle (src/fluidsimulation.h:39-48).
"""
import ctypes as C

import numpy as np

from . import _lib

# field ids (include/flip_b200.h)
F_LIQUID_SDF, F_SOLID_SDF = 0, 1
F_U, F_V, F_W = 2, 3, 4
F_SAVED_U, F_SAVED_V, F_SAVED_W = 5, 6, 7
F_WEIGHT_U, F_WEIGHT_V, F_WEIGHT_W = 8, 9, 10
F_PRESSURE, F_VISCOSITY = 11, 12
F_VOL_CENTER, F_VOL_U, F_VOL_V, F_VOL_W, F_VOL_EDGE_U, F_VOL_EDGE_V, F_VOL_EDGE_W = 13, 14, 15, 16, 17, 18, 19


class FlipError(RuntimeError):
    pass


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class FlipSim:
    def __init__(self, ni, nj, nk, dx, lib=None):
        self.lib = lib or _lib.default_library()
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.dx = float(np.float32(dx))
        h = C.c_void_p()
        rc = self.lib.flip_create(self.ni, self.nj, self.nk, C.c_float(dx), C.byref(h))
        if rc != 0:
            raise FlipError("flip_create failed (%d): %s" % (rc, (self.lib.flip_last_error(None) or b"").decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.flip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FlipError("flip_b200 error %d: %s" % (rc, (self.lib.flip_last_error(self.h) or b"").decode()))

    # ---- shapes ----
    def field_shape(self, f):
        ni, nj, nk = self.ni, self.nj, self.nk
        cells, nodes = (nk, nj, ni), (nk + 1, nj + 1, ni + 1)
        u, v, w = (nk, nj, ni + 1), (nk, nj + 1, ni), (nk + 1, nj, ni)
        shapes = {F_LIQUID_SDF: cells, F_SOLID_SDF: nodes, F_U: u, F_V: v, F_W: w, F_SAVED_U: u, F_SAVED_V: v,
                F_SAVED_W: w, F_WEIGHT_U: u, F_WEIGHT_V: v, F_WEIGHT_W: w, F_PRESSURE: cells, F_VISCOSITY: nodes,
                F_VOL_CENTER: cells, F_VOL_U: u, F_VOL_V: v, F_VOL_W: w, F_VOL_EDGE_U: (nk + 1, nj + 1, ni),
                F_VOL_EDGE_V: (nk + 1, nj, ni + 1), F_VOL_EDGE_W: (nk, nj + 1, ni + 1)}
        if f not in shapes:
            raise FlipError("unknown field id %r (see include/flip_b200.h)" % (f,))
        return shapes[f]

    # ---- scene ----
    def set_solid_sdf(self, phi):
        phi = np.ascontiguousarray(phi, np.float32)
        assert phi.shape == self.field_shape(F_SOLID_SDF)
        self._ck(self.lib.flip_set_solid_sdf(self.h, _fp(phi)))

    def set_particles(self, p):
        p = np.ascontiguousarray(p, np.float32)
        assert p.ndim == 2 and p.shape[1] == 6
        self._ck(self.lib.flip_set_particles(self.h, _fp(p), len(p)))

    def num_particles(self):
        n = C.c_int64()
        self._ck(self.lib.flip_num_particles(self.h, C.byref(n)))
        return n.value

    def get_particles(self, out=None):
        n = self.num_particles()
        if out is None:
            out = np.empty((n, 6), np.float32)
        m = C.c_int64()
        self._ck(self.lib.flip_get_particles(self.h, _fp(out), len(out), C.byref(m)))
        return out[: m.value]

    # ---- scene construction on the device (csrc/scene.cu) ----
    @staticmethod
    def _mesh(verts, tris):
        verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.int32)
        assert verts.ndim == 2 and verts.shape[1] == 3 and tris.ndim == 2 and tris.shape[1] == 3
        return verts, tris, _fp(verts), tris.ctypes.data_as(C.POINTER(C.c_int32))

    def reset_boundary(self):
        """the domain-box boundary FluidSimulation::initialize / resetBoundary builds"""
        self._ck(self.lib.flip_reset_boundary(self.h))

    def add_boundary(self, verts, tris, inverted=False):
        verts, tris, vp, tp = self._mesh(verts, tris)
        self._ck(self.lib.flip_add_boundary_mesh(self.h, vp, len(verts), tp, len(tris), int(inverted)))

    def add_liquid(self, verts, tris):
        verts, tris, vp, tp = self._mesh(verts, tris)
        n = C.c_int64()
        self._ck(self.lib.flip_add_liquid_mesh(self.h, vp, len(verts), tp, len(tris), C.byref(n)))
        return n.value

    def mesh_sdf(self, verts, tris):
        verts, tris, vp, tp = self._mesh(verts, tris)
        out = np.empty(self.field_shape(F_SOLID_SDF), np.float32)
        self._ck(self.lib.flip_mesh_sdf(self.h, vp, len(verts), tp, len(tris), _fp(out)))
        return out

    def srand(self, seed=1):
        self.lib.flip_srand(C.c_uint(seed))

    def get_positions_async(self, out):
        """start an asynchronous export of the positions into `out` ((n, 3) float32, ideally pinned); output_wait() completes it"""
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.ndim == 2 and out.shape[1] == 3
        n = C.c_int64()
        self._ck(self.lib.flip_get_positions_async(self.h, _fp(out), len(out), C.byref(n)))
        return n.value

    def output_wait(self):
        self._ck(self.lib.flip_output_wait(self.h))

    def set_viscosity(self, v):
        if np.isscalar(v):
            self._ck(self.lib.flip_set_viscosity_uniform(self.h, C.c_float(v)))
        else:
            v = np.ascontiguousarray(v, np.float32)
            assert v.shape == self.field_shape(F_VISCOSITY)
            self._ck(self.lib.flip_set_viscosity_grid(self.h, _fp(v)))

    def set_gravity(self, gx, gy, gz):
        self._ck(self.lib.flip_set_gravity(self.h, C.c_float(gx), C.c_float(gy), C.c_float(gz)))

    # ---- stepping ----
    def advance(self, dt):
        n = C.c_int()
        self._ck(self.lib.flip_advance(self.h, C.c_float(dt), C.byref(n)))
        return n.value

    def substep(self, dt):
        self._ck(self.lib.flip_substep(self.h, C.c_float(dt)))

    def cfl(self):
        v = C.c_float()
        self._ck(self.lib.flip_cfl(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._ck(self.lib.flip_synchronize(self.h))

    # ---- stages ----
    def update_liquid_sdf(self):
        self._ck(self.lib.flip_stage_update_liquid_sdf(self.h))

    def advect_velocity_field(self):
        self._ck(self.lib.flip_stage_advect_velocity_field(self.h))

    def add_body_force(self, dt):
        self._ck(self.lib.flip_stage_add_body_force(self.h, C.c_float(dt)))

    def apply_viscosity(self, dt):
        self._ck(self.lib.flip_stage_apply_viscosity(self.h, C.c_float(dt)))

    def project(self, dt):
        self._ck(self.lib.flip_stage_project(self.h, C.c_float(dt)))

    def constrain(self):
        self._ck(self.lib.flip_stage_constrain(self.h))

    def advect_particles(self, dt):
        self._ck(self.lib.flip_stage_advect_particles(self.h, C.c_float(dt)))

    def solve_pressure(self, dt):
        self._ck(self.lib.flip_solve_pressure(self.h, C.c_float(dt)))

    def apply_pressure(self, dt):
        self._ck(self.lib.flip_apply_pressure(self.h, C.c_float(dt)))

    def extrapolate(self):
        self._ck(self.lib.flip_extrapolate(self.h))

    def viscosity_volumes(self):
        self._ck(self.lib.flip_viscosity_volumes(self.h))

    # ---- fields ----
    def get_field(self, f):
        out = np.empty(self.field_shape(f), np.float32)
        self._ck(self.lib.flip_get_field(self.h, f, _fp(out)))
        return out

    def set_field(self, f, a):
        a = np.ascontiguousarray(a, np.float32)
        assert a.shape == self.field_shape(f), (a.shape, self.field_shape(f))
        self._ck(self.lib.flip_set_field(self.h, f, _fp(a)))

    def get_mac(self):
        return self.get_field(F_U), self.get_field(F_V), self.get_field(F_W)

    def set_mac(self, u, v, w):
        self.set_field(F_U, u); self.set_field(F_V, v); self.set_field(F_W, w)

    def get_saved_mac(self):
        return self.get_field(F_SAVED_U), self.get_field(F_SAVED_V), self.get_field(F_SAVED_W)

    def set_saved_mac(self, u, v, w):
        self.set_field(F_SAVED_U, u); self.set_field(F_SAVED_V, v); self.set_field(F_SAVED_W, w)

    def get_weights(self):
        return self.get_field(F_WEIGHT_U), self.get_field(F_WEIGHT_V), self.get_field(F_WEIGHT_W)

    def get_valid(self):
        outs = []
        for c, f in enumerate((F_U, F_V, F_W)):
            o = np.empty(self.field_shape(f), np.uint8)
            self._ck(self.lib.flip_get_valid(self.h, c, o.ctypes.data_as(C.POINTER(C.c_uint8))))
            outs.append(o)
        return tuple(outs)

    def set_valid(self, u, v, w):
        for c, (f, a) in enumerate(zip((F_U, F_V, F_W), (u, v, w))):
            a = np.ascontiguousarray(a, np.uint8)
            assert a.shape == self.field_shape(f)
            self._ck(self.lib.flip_set_valid(self.h, c, a.ctypes.data_as(C.POINTER(C.c_uint8))))

    # ---- multi-GPU ----
    def dist_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self.lib.flip_dist_unique_id(buf)
        if rc != 0:
            raise FlipError("flip_dist_unique_id failed (%d)" % rc)
        return buf.raw

    def dist_init(self, rank, nranks, unique_id=None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._ck(self.lib.flip_dist_init(self.h, int(rank), int(nranks), buf))

    def dist_p2p_export(self):
        n = self.lib.flip_dist_p2p_blob_size()
        buf = C.create_string_buffer(n)
        self._ck(self.lib.flip_dist_p2p_export(self.h, buf))
        return buf.raw

    def dist_p2p_import(self, blobs_in_rank_order):
        data = b"".join(blobs_in_rank_order)
        buf = C.create_string_buffer(data, len(data))
        self._ck(self.lib.flip_dist_p2p_import(self.h, buf))

    # ---- params / stats ----
    def set_param(self, name, value):
        self._ck(self.lib.flip_set_param(self.h, name.encode(), C.c_double(value)))

    def time_kernel(self, name, reps=20):
        """(ms per launch, algorithmic bytes per launch) of a named hot kernel on the last solve's data."""
        ms, nb = C.c_float(), C.c_uint64()
        self._ck(self.lib.flip_time_kernel(self.h, name.encode(), int(reps), C.byref(ms), C.byref(nb)))
        return float(ms.value), int(nb.value)

    def event_record(self, slot):
        self._ck(self.lib.flip_event_record(self.h, int(slot)))

    def event_elapsed_ms(self, slot_from, slot_to):
        ms = C.c_float()
        self._ck(self.lib.flip_event_elapsed_ms(self.h, int(slot_from), int(slot_to), C.byref(ms)))
        return float(ms.value)

    def stats(self):
        st = _lib.flip_stats()
        self._ck(self.lib.flip_get_stats(self.h, C.byref(st)))
        d = {k: getattr(st, k) for k, _ in st._fields_ if k != "stage_ms"}
        d["stage_ms"] = list(st.stage_ms)
        return d
