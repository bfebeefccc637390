"""experiment (dev only): recycle Ritz vectors of one substep's multigrid-PCG solve as a deflation space for the next
substep's solve.  python dev/visc_exp_recycle.py /tmp/visc_state_128_3.npz /tmp/visc_state_128_4.npz [/tmp/visc_state_128_5.npz]"""
import sys, time
import numpy as np
sys.path.insert(0, "dev")
from visc_proto import *


def setup(path):
    st = np.load(path)
    L0 = make_level0(st)
    A, b = assemble(L0, with_rhs=True)
    mg = MG(L0, A, nlev=8, coarse_exact=False, verbose=False, galerkin=True, minvol=0.0, smoother="l1")
    mg.l1_scale = 1.6
    mg.pre_levels = [3, 1, 2, 2, 2, 2, 2, 2]
    glob = np.nonzero(L0.num >= 0)[0]      # global (component * T + padded id) index of every unknown
    return L0, A, b, mg, glob


def pcg_lanczos(A, b, M, x0=None, keep=True, project=None, maxit=400):
    """PCG; returns x, iterations, and (Z, T) = M-orthonormal Lanczos basis and tridiagonal.
    project: (W, AW, G) -> deflated PCG: every preconditioned residual is made A-orthogonal to span(W)."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    tol = 1e-6 * np.abs(b).max()
    def prec(r):
        z = M(r)
        if project is not None:
            W, AW, G = project
            z = z - W @ (G @ (AW.T @ z))
        return z
    z = prec(r); s = z.copy(); rho = r @ z
    al, be, Z = [], [], []
    for it in range(1, maxit + 1):
        if keep: Z.append(z / np.sqrt(rho))
        q = A @ s; alpha = rho / (s @ q); x += alpha * s; r -= alpha * q
        al.append(alpha)
        if np.abs(r).max() <= tol: break
        z = prec(r); rn = r @ z; beta = rn / rho; be.append(beta); s = z + beta * s; rho = rn
    n = len(al); T = np.zeros((n, n))
    for k in range(n):
        T[k, k] = 1 / al[k] + (be[k - 1] / al[k - 1] if k > 0 else 0)
        if k < n - 1: T[k, k + 1] = T[k + 1, k] = np.sqrt(be[k]) / al[k]
    return x, it, (np.array(Z).T if keep else None), T


paths = sys.argv[1:]
L0a, Aa, ba, mga, ga = setup(paths[0])
x, it, Z, T = pcg_lanczos(Aa, ba, mga.vcycle)
w, V = np.linalg.eigh(T)
print(paths[0], "unknowns", len(ba), "iterations", it, "lowest Ritz values", np.round(w[:6], 4), flush=True)
for nxt in paths[1:]:
    L0b, Ab, bb, mgb, gb = setup(nxt)
    x0, it0, Zb, Tb = pcg_lanczos(Ab, bb, mgb.vcycle)
    print(nxt, "unknowns", len(bb), "baseline iterations", it0, flush=True)
    for k in (3, 6, 10):
        W_prev = Z @ V[:, :k]                       # Ritz vectors on the previous unknown set
        # map to the new unknown set by face index
        full = np.zeros(3 * L0a.T); Wn = np.zeros((len(bb), k))
        for c in range(k):
            full[:] = 0; full[ga] = W_prev[:, c]; Wn[:, c] = full[gb]
        AW = Ab @ Wn; G = np.linalg.inv(Wn.T @ AW)
        xi = Wn @ (G @ (Wn.T @ bb))                  # Galerkin start in span(W)
        _, it1, _, _ = pcg_lanczos(Ab, bb, mgb.vcycle, x0=xi, keep=False)
        _, it2, _, _ = pcg_lanczos(Ab, bb, mgb.vcycle, x0=xi, keep=False, project=(Wn, AW, G))
        print("   k=%d recycled Ritz vectors: init-only %d iterations, deflated PCG %d iterations" % (k, it1, it2), flush=True)
    # carry on: next pair uses this solve's Ritz vectors
    w, V = np.linalg.eigh(Tb); Z = Zb; L0a, ga = L0b, gb
