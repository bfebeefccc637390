"""dev: the two Galerkin-product kernels (mg_build 0 = lane-ordered scatter, 1 = gather) must give identical rows.
Usage: python dev/gmg_build_ab.py [emu|cuda] [n]"""
import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import common, parity_checks as pc
from oracle import refsim
refsim.build()
which = sys.argv[1] if len(sys.argv) > 1 else "emu"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
from flipviscosity3d_b200 import _lib
lib = common.emu_library() if which == "emu" else _lib.default_library()
lib.flip_debug_gmg_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

def levels(sim):
    out = []
    lvl = 1
    while True:
        info = (C.c_int * 10)()
        if lib.flip_debug_gmg_level(sim.h, lvl, info, None, None, None) != 0: break
        T, nrows, stride = info[6], info[7], info[9]
        rows = np.zeros(nrows, np.int32); S = np.zeros((nrows, stride), np.float32); d = np.zeros(3 * T, np.float32)
        assert lib.flip_debug_gmg_level(sim.h, lvl, info, rows.ctypes.data, S.ctypes.data, d.ctypes.data) == 0
        out.append((rows, S)); lvl += 1
    return out

sim, ref = pc.build_pair(lib, refsim, n=n)
pc.prepare_mid_substep(sim, ref)
res = []
for b in (0, 1):
    pc.sync_grid_state(sim, ref)
    sim.set_param("mg_build", b)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    print("mg_build", b, "iterations", st["viscosity_iterations"], "converged", st["viscosity_converged"])
    res.append(levels(sim))
for l, ((r0, S0), (r1, S1)) in enumerate(zip(*res)):
    assert np.array_equal(r0, r1)
    print("level", l + 1, "rows", len(r0), "identical" if np.array_equal(S0, S1) else "max rel diff %.3g" % (np.abs(S0 - S1).max() / np.abs(S0).max()),
          "fill %.3f" % ((S0[:, :235] != 0).mean()))
