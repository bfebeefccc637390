"""Dev tool (CPU only): run the oracle on the bunny scene and dump the inputs of the viscosity solve
at chosen substeps to /tmp/visc_state_<n>_<step>.npz, for solver prototyping (dev/visc_proto.py).

  python dev/visc_dump_state.py 64 3 30 60
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import make_ref_scene  # noqa: E402

n = int(sys.argv[1])
steps = sorted(int(a) for a in sys.argv[2:])
ref = make_ref_scene(n, viscosity=5.0)
print("particles", ref.num_particles(), flush=True)
step = 0
t0 = time.time()
while step <= steps[-1]:
    dt = min(ref.cfl(), 0.01)
    ref.update_liquid_sdf()
    ref.advect_velocity_field()
    ref.add_body_force(dt)
    if step in steps:
        vols = ref.viscosity_volumes()
        u, v, w = ref.get_mac()
        np.savez_compressed("/tmp/visc_state_%d_%d.npz" % (n, step), n=n, dt=dt, u=u, v=v, w=w, solid=ref.get_solid_sdf(),
                            vc=vols[0], vu=vols[1], vv=vols[2], vw=vols[3], veu=vols[4], vev=vols[5], vew=vols[6])
        print("dumped step", step, "dt", dt, flush=True)
    info = ref.apply_viscosity(dt)
    ref.project(dt)
    ref.constrain()
    ref.advect_particles(dt)
    print("step", step, "dt %.5f" % dt, "visc", info["iters"], "%.3e" % info["resid"], "unk", info["unknowns"], "t %.0fs" % (time.time() - t0), flush=True)
    step += 1
