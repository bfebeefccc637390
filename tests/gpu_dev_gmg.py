"""Dev helper: viscosity solve with the diagonal / rediscretised-MG / Galerkin-MG preconditioner."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from flipviscosity3d_b200 import FlipSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 0]
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
params = dict(a.split('=') for a in sys.argv[4:])
phi, p = bench.build_scene(n)
out = {}
for mode in modes:
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    sim.set_param('viscosity_precond', mode)
    for k, v in params.items(): sim.set_param(k, float(v))
    for step in range(nsteps):
        sim.substep(0.01)
        st = sim.stats()
        print(n, 'mode', mode, 'step', step, 'visc it', st['viscosity_iterations'], 'conv', st['viscosity_converged'], 'ms %.1f' % st['viscosity_solve_ms'],
              'unknowns', st['viscosity_unknowns'], 'pres it', st['pressure_iterations'], 'pms %.1f' % st['pressure_solve_ms'], 'total %.1f' % st['stage_ms'][7], flush=True)
    out[mode] = sim.get_particles()
    sim.close()
if len(modes) > 1:
    a, b = out[modes[0]], out[modes[1]]
    print('max |x_%d - x_%d| = %.3e' % (modes[0], modes[1], np.abs(a[:, :3] - b[:, :3]).max()))
