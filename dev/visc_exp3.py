"""experiments: Galerkin MG smoothing variants, h-dependence (dev only)"""
import sys, time
import numpy as np
sys.path.insert(0, "dev")
from visc_proto import *
st = np.load(sys.argv[1])
L0 = make_level0(st)
A, b = assemble(L0, with_rhs=True)
d = A.diagonal()
print(sys.argv[1], "unknowns", L0.nunk, flush=True)
if "--jac" in sys.argv:
    t0 = time.time(); x, it = pcg(A, b, lambda r: r / d); print("jacobi", it, "%.1fs" % (time.time() - t0), flush=True)
for kw in [dict(galerkin=True), dict(galerkin=True, pre=1), dict(galerkin=True, pre=1, omega=0.7), dict(galerkin=True, pre=2, omega=0.7),
           dict(galerkin=True, pre=3, omega=0.6),
           dict(galerkin=True, smoother="cheb", pre=2), dict(galerkin=True, smoother="cheb", pre=3), dict(galerkin=False)]:
    t0 = time.time()
    mg = MG(L0, A, verbose=False, **kw)
    x, it = pcg(A, b, mg.vcycle, maxit=3000)
    print(kw, "levels", len(mg.A), "iterations", it, "%.1fs" % (time.time() - t0), flush=True)
