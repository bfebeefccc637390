"""dev: a parameter that must not change the viscosity solve at all (bit-identical velocities, same iteration count).
Usage: python dev/visc_param_ab.py [emu|cuda] [n] [param] [v0] [v1]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import common, parity_checks as pc
from oracle import refsim
refsim.build()
which = sys.argv[1] if len(sys.argv) > 1 else "emu"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
name = sys.argv[3] if len(sys.argv) > 3 else "mg_tail"
vals = [float(v) for v in sys.argv[4:6]] if len(sys.argv) > 5 else [0.0, 1.0]
from flipviscosity3d_b200 import _lib
lib = common.emu_library() if which == "emu" else _lib.default_library()
sim, ref = pc.build_pair(lib, refsim, n=n)
pc.prepare_mid_substep(sim, ref)
res = []
for v in vals:
    pc.sync_grid_state(sim, ref)
    sim.set_param(name, v)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    print(name, v, "iterations", st["viscosity_iterations"], "converged", st["viscosity_converged"], "resid %.3e" % st["viscosity_residual"])
    res.append([f.copy() for f in sim.get_mac()])
for a, b in zip(*res):
    print("identical" if np.array_equal(a, b) else "max diff %.3g" % np.abs(a - b).max())
