"""Dev helper: stage timings on the GPU for a few grid sizes (not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import _analytic_scene
from flipviscosity3d_b200 import FlipSim

for n, visc in [(64, 0.0), (64, 2.0), (128, 0.0), (128, 2.0), (256, 0.0), (256, 2.0)]:
    phi, p = _analytic_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(visc)
    t0 = time.time()
    for f in range(4):
        ns = sim.advance(0.01)
        st = sim.stats()
        print(n, visc, 'frame', f, 'substeps', ns, 'stage_ms', ['%.2f' % x for x in st['stage_ms']], 'pit', st['pressure_iterations'],
              'pconv', st['pressure_converged'], 'pms %.2f' % st['pressure_solve_ms'], 'vit', st['viscosity_iterations'], 'vconv',
              st['viscosity_converged'], 'vms %.2f' % st['viscosity_solve_ms'], 'blocks', st['pressure_active_blocks'], st['viscosity_active_blocks'], flush=True)
    print(n, visc, 'np', len(p), 'wall', time.time() - t0, flush=True)
    sim.close()
