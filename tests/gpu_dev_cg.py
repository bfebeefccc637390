"""Dev helper: CG launch-configuration experiments on the bench scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from flipviscosity3d_b200 import FlipSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
phi, p = bench.build_scene(n)
cases = [dict(use_graphs=0, cg_grid_mult=2), dict(use_graphs=1, cg_grid_mult=2), dict(use_graphs=1, cg_grid_mult=3), dict(use_graphs=1, cg_grid_mult=4),
         dict(use_graphs=1, cg_grid_mult=4, cg_chunk=64), ]
for kw in cases:
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    for k, v in kw.items(): sim.set_param(k, v)
    for step in range(3):
        sim.substep(0.01)
        st = sim.stats()
    print(n, kw, 'visc it', st['viscosity_iterations'], 'ms %.1f' % st['viscosity_solve_ms'], 'us/it %.1f' % (1e3 * st['viscosity_solve_ms'] / max(1, st['viscosity_iterations'])),
          'pres it', st['pressure_iterations'], 'pms %.2f' % st['pressure_solve_ms'], 'us/it %.1f' % (1e3 * st['pressure_solve_ms'] / max(1, st['pressure_iterations'])), 'total %.1f' % st['stage_ms'][7], flush=True)
    sim.close()
