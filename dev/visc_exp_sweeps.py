"""experiment (dev only): sweeps per level vs iteration count, with the per-launch costs measured at 256^3 on B200"""
import sys, time
import numpy as np
sys.path.insert(0, "dev")
from visc_proto import *

def cost_us(pl):
    l0, l1, l2 = pl[0], pl[1], pl[2]
    return 100 + 85 + 25 * (2 * l0 - 1) + (28 + 80 * (2 * l1)) + (15 * 2 * l2 + 10) + 3 * 8 * (2 * l2 + 2)

states = []
for path in sys.argv[1:]:
    st = np.load(path)
    L0 = make_level0(st)
    A, b = assemble(L0, with_rhs=True)
    states.append((L0, A, b))
configs = [[3, 1, 2], [3, 1, 1], [2, 1, 2], [2, 1, 1], [4, 1, 2], [4, 1, 1], [5, 1, 2], [3, 2, 2], [4, 2, 1], [6, 1, 1]]
for pl in configs:
    its = []
    for L0, A, b in states:
        mg = MG(L0, A, nlev=8, coarse_exact=False, verbose=False, galerkin=True, minvol=0.0, smoother="l1")
        mg.l1_scale = 1.6
        mg.pre_levels = pl + [pl[2]] * 5
        x, it = pcg(A, b, mg.vcycle, maxit=400)
        its.append(it)
    c = cost_us(pl)
    print("sweeps (L0, L1, deeper) =", pl, "iterations", its, "cost/iteration %d us -> solve %.1f ms (mean)" % (c, np.mean(its) * c / 1000), flush=True)
