"""Dev helper: per-substep iteration counts / times on the bench scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flipviscosity3d_b200 import FlipSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
phi, p = bench.build_scene(n)
for kw in [dict(cg_variant=0), dict(cg_variant=1), dict(cg_variant=1, cg_grid_mult=3), dict(cg_variant=1, cg_grid_mult=1)]:
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    for k, v in kw.items(): sim.set_param(k, v)
    line = []
    for step in range(4):
        sim.substep(0.01)
        st = sim.stats()
        line.append('%d:%dit/%.0fms(%.0fus)p%d/%.1f' % (step, st['viscosity_iterations'], st['viscosity_solve_ms'], 1e3 * st['viscosity_solve_ms'] / max(1, st['viscosity_iterations']), st['pressure_iterations'], st['pressure_solve_ms']))
    print(n, kw, ' '.join(line), flush=True)
    sim.close()
