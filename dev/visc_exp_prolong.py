"""experiment (dev only): prolongation weights that reproduce linear functions where parents are missing at the free
surface (minimum-norm correction of the trilinear weights over the parents that exist), same parent sets / window."""
import sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "dev")
import visc_proto as vp
from visc_proto import *


def prolongation_lin(F, C, clip=None):
    rows, cols, vals = [], [], []
    nfix = 0
    for m in range(3):
        fid = np.arange(F.T)[F.unk[m]]
        f = [F.ijk[0][fid], F.ijk[1][fid], F.ijk[2][fid]]
        r = F.num[m * F.T + fid]

        def parents(own, n):
            p0 = n >> 1
            odd = (n & 1) == 1
            if own:   # vertex-centred: fine position n/2
                return (p0, np.where(odd, p0 + 1, p0), np.where(odd, 0.5, 1.0), np.where(odd, 0.5, 0.0),
                        np.where(odd, -0.5, 0.0), np.where(odd, 0.5, 0.0))
            p1 = np.where(odd, p0 + 1, p0 - 1)   # cell-centred: fine position n/2 - 1/4 (in coarse cell-centre coordinates)
            return (p0, p1, np.full(n.shape, 0.75), np.full(n.shape, 0.25), np.where(odd, -0.25, 0.25), np.where(odd, 0.75, -0.75))

        P3 = [parents(m == a, f[a]) for a in range(3)]
        W = np.zeros((len(fid), 8)); OK = np.zeros((len(fid), 8), bool); CN = np.zeros((len(fid), 8), np.int64)
        D = np.zeros((len(fid), 8, 3))
        q = 0
        for a in range(2):
            for bb in range(2):
                for c in range(2):
                    w = P3[0][2 + a] * P3[1][2 + bb] * P3[2][2 + c]
                    I, J, K = P3[0][a], P3[1][bb], P3[2][c]
                    ok = (w > 0) & (I >= 0) & (J >= 0) & (K >= 0) & (I <= C.ni) & (J <= C.nj) & (K <= C.nk)
                    cid = (np.clip(I, -1, C.ni + 1) + 1) + C.sy * (np.clip(J, -1, C.nj + 1) + 1) + C.sz * (np.clip(K, -1, C.nk + 1) + 1)
                    cn = C.num[m * C.T + cid]
                    ok &= cn >= 0
                    W[:, q] = np.where(ok, w, 0); OK[:, q] = ok; CN[:, q] = cn
                    D[:, q, 0] = P3[0][4 + a]; D[:, q, 1] = P3[1][4 + bb]; D[:, q, 2] = P3[2][4 + c]
                    q += 1
        tot = W.sum(1)
        full = tot > 0.999999
        Wn = W / np.where(tot > 0, tot, 1)[:, None]
        part = np.nonzero((~full) & (tot > 0))[0]
        for i in part:
            sel = np.nonzero(OK[i])[0]
            w0 = Wn[i, sel]
            Cm = np.vstack([np.ones(len(sel)), D[i, sel].T])          # 4 x np
            t = np.array([1.0, 0, 0, 0])
            corr = np.linalg.lstsq(Cm, t - Cm @ w0, rcond=1e-9)[0]     # minimum-norm correction
            w1 = w0 + corr
            if clip is not None and (np.abs(w1).max() > clip):
                w1 = w0
            else:
                nfix += 1
            Wn[i, sel] = w1
        for q in range(8):
            ok = OK[:, q]
            rows.append(r[ok]); cols.append(CN[ok, q]); vals.append(Wn[ok, q])
    print("   linear-reproducing weights on %d surface rows" % nfix, flush=True)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(F.nunk, C.nunk))


for path in sys.argv[1:]:
    st = np.load(path)
    L0 = make_level0(st)
    A, b = assemble(L0, with_rhs=True)
    for label, fn in [("renormalised trilinear (shipping)", None), ("linear-reproducing, unclipped", lambda F, C, rn, tr: prolongation_lin(F, C)),
                      ("linear-reproducing, |w| <= 2", lambda F, C, rn, tr: prolongation_lin(F, C, clip=2.0))]:
        orig = vp.prolongation
        if fn is not None:
            vp.prolongation = fn
        try:
            mg = MG(L0, A, nlev=8, coarse_exact=False, verbose=False, galerkin=True, minvol=0.0, smoother="l1")
        finally:
            vp.prolongation = orig
        mg.l1_scale = 1.6
        mg.pre_levels = [3, 1, 2, 2, 2, 2, 2, 2]
        x, it = pcg(A, b, mg.vcycle, maxit=400)
        print(path, label, "iterations", it, flush=True)
