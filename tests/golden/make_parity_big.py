"""Generates the teacher-forcing fixture for the headline-configuration parity check (tests/parity_big.py):
the ORACLE (compiled reference, CPU) takes the bunny-in-sphere scene at N^3 / viscosity mu one substep forward, runs the
grid stages of the next substep up to the body force, and then solves the viscosity system three ways from that same
state: with the reference's own settings (tol 1e-6, 700-iteration cap: the truncated iterate it would really return at
256^3) and with the cap raised to 20000 at tol 1e-6 and 1e-8; the pressure system is then solved from the tightest
viscosity result.  Arrays are cropped to the bounding box of the liquid (+8 cells) and stored compressed under
tests/golden_big/ (git-ignored: ~tens of MB; it travels to the GPU box with the snapshot).

  python tests/golden/make_parity_big.py [N=256] [mu=5]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import common

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mu = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
DT = 0.01
out = os.path.join(ROOT, "tests", "golden_big", "parity_%d_mu%g.npz" % (n, mu))
t0 = time.time()
ref = common.make_ref_scene(n, viscosity=mu)
print("scene %.0f s, %d particles" % (time.time() - t0, ref.num_particles()), flush=True)
t0 = time.time(); ref.substep(DT); print("substep 1: %.0f s" % (time.time() - t0), flush=True)
ref.update_liquid_sdf(); ref.advect_velocity_field(); ref.add_body_force(DT)
phi = ref.get_liquid_sdf()
pre = ref.get_mac()
kk, jj, ii = np.where(phi < 0)
pad = 8
lo = [max(0, int(a.min()) - pad) for a in (kk, jj, ii)]
hi = [min(n, int(a.max()) + pad + 1) for a in (kk, jj, ii)]
crop = lambda a: np.ascontiguousarray(a[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1][:min(a.shape[0], hi[0] + 1) - lo[0]])
def crop3(a):
    return np.ascontiguousarray(a[lo[0]:min(a.shape[0], hi[0] + 1), lo[1]:min(a.shape[1], hi[1] + 1), lo[2]:min(a.shape[2], hi[2] + 1)])
for a in pre:   # nothing non-zero may fall outside the crop
    b = a.copy(); b[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = 0
    assert not b.any()
save = dict(n=n, mu=mu, dt=DT, lo=np.array(lo), hi=np.array(hi), phi=crop3(phi), phi_fill=np.float32(phi[0, 0, 0]))
for name, a in zip("uvw", pre):
    save["pre_" + name] = crop3(a)
diag = {}
for tag, tol, maxit in (("ref700", 1e-6, 700), ("ref1e6", 1e-6, 20000), ("ref1e8", 1e-8, 40000)):
    ref.set_mac(*pre)
    t0 = time.time()
    info = ref.apply_viscosity(DT, tol=tol, maxit=maxit)
    sol = ref.get_mac()
    info["seconds"] = time.time() - t0
    ref.set_mac(*pre)
    info.update({"true_" + k: v for k, v in ref.viscosity_residual(DT, *sol).items()})
    print(tag, info, flush=True)
    diag[tag] = info
    for name, a in zip("uvw", sol):
        save[tag + "_" + name] = crop3(a)
    last = sol
save["diag"] = np.array(repr(diag))
# pressure from the tightest viscosity result
ref.set_mac(*last)
ref.compute_weights()
t0 = time.time()
pr = ref.solve_pressure(DT)
print("pressure %.0f s, max|p| %g" % (time.time() - t0, np.abs(pr).max()), flush=True)
save["pressure"] = crop3(pr)
np.savez_compressed(out, **save)
print("wrote", out, os.path.getsize(out) / 1e6, "MB", flush=True)
