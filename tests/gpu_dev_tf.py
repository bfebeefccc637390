"""Dev helper: teacher forcing — load an oracle particle state and run one substep per preconditioner mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from flipviscosity3d_b200 import FlipSim
n = 256
phi, p0 = bench.build_scene(n)
p = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tmp_state", "p256_3.npy"))
for mode in [int(a) for a in sys.argv[1].split(",")]:
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    sim.set_param('viscosity_precond', mode)
    sim.set_param('verbose', 2)
    for step in range(2):
        sim.substep(0.01)
        st = sim.stats()
        print(n, 'TF mode', mode, 'step', step, 'visc it', st['viscosity_iterations'], 'ms %.1f' % st['viscosity_solve_ms'], 'unknowns', st['viscosity_unknowns'], flush=True)
    sim.close()
