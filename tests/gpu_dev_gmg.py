"""Dev helper: a few substeps of the bench scene with a given viscosity preconditioner and parameters.
Usage: python tests/gpu_dev_gmg.py [n] [modes, e.g. 2 or 2,0] [substeps] [name=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from flipviscosity3d_b200 import FlipSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 0]
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
params = dict(a.split('=') for a in sys.argv[4:])
scene = params.pop("scene", "bunny")
out = {}
for mode in modes:
    sim = FlipSim(n, n, n, 1.0 / n)
    bench.load_scene_device(sim, scene)
    sim.set_viscosity(5.0)
    sim.set_param('viscosity_precond', mode)
    for k, v in params.items(): sim.set_param(k, float(v))
    for step in range(nsteps):
        sim.substep(0.01)
        st = sim.stats()
        print(n, 'mode', mode, 'step', step, 'visc it', st['viscosity_iterations'], 'conv', st['viscosity_converged'], 'ms %.2f' % st['viscosity_solve_ms'],
              'setup %.2f' % st['viscosity_setup_ms'], 'unknowns', st['viscosity_unknowns'], 'pres it', st['pressure_iterations'], 'conv', st['pressure_converged'],
              'res %.2e' % st['pressure_residual'], 'pms %.2f' % st['pressure_solve_ms'], 'total %.2f' % st['stage_ms'][7], flush=True)
    out[mode] = sim.get_particles()
    sim.close()
if len(modes) > 1:
    a, b = out[modes[0]], out[modes[1]]
    print('max |x_%d - x_%d| = %.3e' % (modes[0], modes[1], np.abs(a[:, :3] - b[:, :3]).max()))
