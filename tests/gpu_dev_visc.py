"""Dev helper: viscosity solve iteration counts / timings, diagonal vs multigrid preconditioner."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from flipviscosity3d_b200 import FlipSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
phi, p = bench.build_scene(n)
for mode, kw in [(1, {}), (1, {'mg_sweeps': 3}), (1, {'mg_sweeps': 1}), (1, {'mg_omega': 0.55}), (0, {})]:
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    sim.set_param('viscosity_precond', mode)
    for k, v in kw.items(): sim.set_param(k, v)
    for step in range(3):
        sim.substep(0.01)
        st = sim.stats()
        print(n, 'mode', mode, kw, 'step', step, 'visc it', st['viscosity_iterations'], 'conv', st['viscosity_converged'], 'ms %.1f' % st['viscosity_solve_ms'],
              'unknowns', st['viscosity_unknowns'], 'pres it', st['pressure_iterations'], 'pms %.1f' % st['pressure_solve_ms'], 'total %.1f' % st['stage_ms'][7], flush=True)
    sim.close()
