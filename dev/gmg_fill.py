"""dev: how many of the 235 stored slots of a Galerkin row hold non-zeros (per level), on the bench scene.
Runs the set-up of one viscosity solve (CPU-emulated kernels by default: minutes at 128^3, ~half an hour at 256^3).
Usage: python dev/gmg_fill.py [emu|cuda] [n] [scene]"""
import ctypes as C, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import common, bench
which = sys.argv[1] if len(sys.argv) > 1 else "emu"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
scene = sys.argv[3] if len(sys.argv) > 3 else "bunny"
from flipviscosity3d_b200 import FlipSim, _lib
lib = common.emu_library() if which == "emu" else _lib.default_library()
lib.flip_debug_gmg_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
t0 = time.time()
phi, p = bench.build_scene(scene, n)
sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
sim.set_param("maxit_scale", 1); sim.set_param("viscosity_maxit", 1)      # the set-up is what is wanted, not the solve
print("scene %s %d^3: %d particles (%.0f s)" % (scene, n, len(p), time.time() - t0), flush=True)
sim.update_liquid_sdf(); sim.advect_velocity_field(); sim.add_body_force(0.01); sim.apply_viscosity(0.01)
print("set-up done (%.0f s), unknowns %d" % (time.time() - t0, sim.stats()["viscosity_unknowns"]), flush=True)
lvl = 1
while True:
    info = (C.c_int * 10)()
    if lib.flip_debug_gmg_level(sim.h, lvl, info, None, None, None) != 0: break
    T, nrows, stride = info[6], info[7], info[9]
    rows = np.zeros(nrows, np.int32); S = np.zeros((nrows, stride), np.float32); d = np.zeros(3 * T, np.float32)
    assert lib.flip_debug_gmg_level(sim.h, lvl, info, rows.ctypes.data, S.ctypes.data, d.ctypes.data) == 0
    nz = (S[:, :235] != 0).sum(1)
    print("level %d: %d rows, fill %.3f of 235 slots (non-zeros per row: mean %.1f, median %d, 10%% %d, 90%% %d, max %d)"
          % (lvl, nrows, nz.mean() / 235, nz.mean(), np.median(nz), np.percentile(nz, 10), np.percentile(nz, 90), nz.max()), flush=True)
    lvl += 1
