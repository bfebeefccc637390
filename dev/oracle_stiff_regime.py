"""Does the reference's own MICCG(0) converge in the stiff regime of the 256^3 bench (dt*mu/dx^2 ~ 3300) when its
700-iteration cap is raised?  usage: oracle_stiff_regime.py N [mu]   (mu defaults to the value that gives the regime)"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import common
n = int(sys.argv[1])
mu = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0 * (256.0 / n) ** 2
DT = 0.01
ref = common.make_ref_scene(n, viscosity=mu)
print("n", n, "mu", mu, "particles", ref.num_particles(), flush=True)
t = time.time(); ref.substep(DT); print("substep", time.time() - t, flush=True)
ref.update_liquid_sdf(); ref.advect_velocity_field(); ref.add_body_force(DT)
pre = ref.get_mac()
for tol, maxit in ((1e-6, 700), (1e-6, 20000)):
    ref.set_mac(*pre)
    t = time.time()
    info = ref.apply_viscosity(DT, tol=tol, maxit=maxit)
    sol = ref.get_mac()
    ref.set_mac(*pre)
    res = ref.viscosity_residual(DT, *sol)
    print("tol", tol, "maxit", maxit, info, "seconds %.1f" % (time.time() - t), res, flush=True)
