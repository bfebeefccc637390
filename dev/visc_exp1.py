"""experiments: MG variants as PCG preconditioner (dev only)"""
import sys, time
import numpy as np
sys.path.insert(0, "dev")
from visc_proto import *

st = np.load(sys.argv[1])
L0 = make_level0(st)
A, b = assemble(L0, with_rhs=True)
d = A.diagonal()
print("unknowns", L0.nunk)
t0 = time.time(); x, it = pcg(A, b, lambda r: r / d); print("jacobi", it, "%.1fs" % (time.time() - t0))
for kw in [dict(galerkin=False, coarse_exact=True), dict(galerkin=True, coarse_exact=True),
           dict(galerkin=False, nlev=2), dict(galerkin=True, nlev=2),
           dict(galerkin=True, nlev=2, pre=1), dict(galerkin=True, nlev=2, pre=3, omega=0.6),
           ]:
    t0 = time.time()
    mg = MG(L0, A, verbose=False, **kw)
    x, it = pcg(A, b, mg.vcycle, maxit=3000)
    print(kw, "levels", len(mg.A), "iterations", it, "%.1fs" % (time.time() - t0), flush=True)
