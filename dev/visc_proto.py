"""Dev tool (CPU only, scipy): prototype preconditioners for the variational viscosity system.

Builds the same matrix-free operator as csrc/viscosity.cu (k_visc_coefs / k_visc_rows / k_visc_apply)
from a state dumped by dev/visc_dump_state.py, as a scipy CSR matrix, and measures PCG iteration
counts (stopping rule of the reference: max|r| <= 1e-6 max|b|) for candidate preconditioners.
Not shipped, not imported by the package.
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Level:
    pass


def padded(a, n3):
    out = np.zeros(n3, a.dtype)
    d, h, w = a.shape
    out[1:1 + d, 1:1 + h, 1:1 + w] = a
    return out


def make_level0(st, mu=5.0):
    n = int(st["n"])
    ni = nj = nk = n
    L = Level()
    L.ni, L.nj, L.nk = ni, nj, nk
    L.shape = (nk + 3, nj + 3, ni + 3)
    L.sy, L.sz = ni + 3, (ni + 3) * (nj + 3)
    L.T = L.shape[0] * L.shape[1] * L.shape[2]
    dx = np.float32(1.0 / n)
    factor = float(st["dt"]) / float(dx) ** 2
    P = lambda a: padded(np.asarray(a, np.float64), L.shape).ravel()
    vc, vu, vv, vw = P(st["vc"]), P(st["vu"]), P(st["vv"]), P(st["vw"])
    veu, vev, vew = P(st["veu"]), P(st["vev"]), P(st["vew"])
    L.cc = 2 * factor * mu * vc
    L.cu, L.cv, L.cw = factor * mu * veu, factor * mu * vev, factor * mu * vew
    L.vol = [vu, vv, vw]
    # face states
    ps = np.asarray(st["solid"], np.float64)
    sc = 0.125 * (ps[:-1, :-1, :-1] + ps[:-1, :-1, 1:] + ps[:-1, 1:, :-1] + ps[:-1, 1:, 1:] + ps[1:, :-1, :-1] + ps[1:, :-1, 1:] +
                  ps[1:, 1:, :-1] + ps[1:, 1:, 1:])
    su = np.ones((nk, nj, ni + 1), bool); su[:, :, 1:-1] = (sc[:, :, :-1] + sc[:, :, 1:]) <= 0
    sv = np.ones((nk, nj + 1, ni), bool); sv[:, 1:-1, :] = (sc[:, :-1, :] + sc[:, 1:, :]) <= 0
    sw = np.ones((nk + 1, nj, ni), bool); sw[1:-1, :, :] = (sc[:-1, :, :] + sc[1:, :, :]) <= 0
    wall = []
    for s in (su, sv, sw):
        o = np.ones(L.shape, bool)
        d, h, w = s.shape
        o[1:1 + d, 1:1 + h, 1:1 + w] = s
        wall.append(o.ravel())
    L.wall = wall
    kk, jj, ii = np.meshgrid(np.arange(-1, nk + 2), np.arange(-1, nj + 2), np.arange(-1, ni + 2), indexing="ij")
    L.interior = ((ii >= 1) & (ii < ni) & (jj >= 1) & (jj < nj) & (kk >= 1) & (kk < nk)).ravel()
    L.ijk = (ii.ravel(), jj.ravel(), kk.ravel())
    sh = lambda f, o: np.roll(f, -o)
    sy, sz = L.sy, L.sz
    anyU = (vu > 0) | (vc > 0) | (sh(vc, -1) > 0) | (sh(vew, sy) > 0) | (vew > 0) | (sh(vev, sz) > 0) | (vev > 0)
    anyV = (vv > 0) | (sh(vew, 1) > 0) | (vew > 0) | (vc > 0) | (sh(vc, -sy) > 0) | (sh(veu, sz) > 0) | (veu > 0)
    anyW = (vw > 0) | (sh(vev, 1) > 0) | (vev > 0) | (sh(veu, sy) > 0) | (veu > 0) | (vc > 0) | (sh(vc, -sz) > 0)
    L.unk = [L.interior & ~wall[0] & anyU, L.interior & ~wall[1] & anyV, L.interior & ~wall[2] & anyW]
    L.vel = [P(st["u"]), P(st["v"]), P(st["w"])]
    return L


def stencil(L):
    """list per component of (col_comp, offset, coef_array) incl. diagonal; coef arrays are full-size."""
    sh = lambda f, o: np.roll(f, -o)
    sy, sz = L.sy, L.sz
    cc, cu, cv, cw = L.cc, L.cu, L.cv, L.cw
    out = []
    # U
    fR, fL, fT, fB, fF, fK = cc, sh(cc, -1), sh(cw, sy), cw, sh(cv, sz), cv
    dU = L.vol[0] + fR + fL + fT + fB + fF + fK
    out.append([(0, 0, dU), (0, 1, -fR), (0, -1, -fL), (0, sy, -fT), (0, -sy, -fB), (0, sz, -fF), (0, -sz, -fK),
                (1, sy, -fT), (1, -1 + sy, fT), (1, 0, fB), (1, -1, -fB), (2, sz, -fF), (2, -1 + sz, fF), (2, 0, fK), (2, -1, -fK)])
    # V
    fR, fL, fT, fB, fF, fK = sh(cw, 1), cw, cc, sh(cc, -sy), sh(cu, sz), cu
    dV = L.vol[1] + fR + fL + fT + fB + fF + fK
    out.append([(1, 0, dV), (1, 1, -fR), (1, -1, -fL), (1, sy, -fT), (1, -sy, -fB), (1, sz, -fF), (1, -sz, -fK),
                (0, 1, -fR), (0, 1 - sy, fR), (0, 0, fL), (0, -sy, -fL), (2, sz, -fF), (2, -sy + sz, fF), (2, 0, fK), (2, -sy, -fK)])
    # W
    fR, fL, fT, fB, fF, fK = sh(cv, 1), cv, sh(cu, sy), cu, cc, sh(cc, -sz)
    dW = L.vol[2] + fR + fL + fT + fB + fF + fK
    out.append([(2, 0, dW), (2, 1, -fR), (2, -1, -fL), (2, sy, -fT), (2, -sy, -fB), (2, sz, -fF), (2, -sz, -fK),
                (0, 1, -fR), (0, 1 - sz, fR), (0, 0, fL), (0, -sz, -fL), (1, sy, -fT), (1, sy - sz, fT), (1, 0, fB), (1, -sz, -fB)])
    return out


def assemble(L, with_rhs=False):
    """CSR over the unknowns of level L (numbering: comp-major, padded id ascending)."""
    T = L.T
    unk = np.concatenate(L.unk)
    num = -np.ones(3 * T, np.int64)
    num[unk] = np.arange(unk.sum())
    L.num, L.nunk = num, int(unk.sum())
    rows, cols, vals = [], [], []
    b = np.zeros(L.nunk)
    st = stencil(L)
    ids_all = np.arange(T)
    for m in range(3):
        rid = ids_all[L.unk[m]]
        r = num[m * T + rid]
        for (cm, off, coef) in st[m]:
            cid = rid + off
            c = num[cm * T + cid]
            v = coef[rid]
            ok = c >= 0
            rows.append(r[ok]); cols.append(c[ok]); vals.append(v[ok])
            if with_rhs and not (cm == m and off == 0):
                solid = L.wall[cm][cid]
                b[r[solid]] -= v[solid] * L.vel[cm][cid[solid]]
        if with_rhs:
            b[r] += L.vol[m][rid] * L.vel[m][rid]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(L.nunk, L.nunk))
    return (A, b) if with_rhs else A


# ---------------------------------------------------------------------------------------------
def pcg(A, b, M=None, tol_rel=1e-6, maxit=20000, x0=None, log=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    tol = tol_rel * np.abs(b).max()
    z = M(r) if M else r
    s = z.copy()
    rho = r @ z
    for it in range(1, maxit + 1):
        q = A @ s
        alpha = rho / (s @ q)
        x += alpha * s
        r -= alpha * q
        rm = np.abs(r).max()
        if log is not None:
            log.append(rm)
        if rm <= tol:
            return x, it
        z = M(r) if M else r
        rho_new = r @ z
        s = z + (rho_new / rho) * s
        rho = rho_new
    return x, maxit


# ---------------------------------------------------------------------------------------------
def coarse_dims(L):
    return (L.ni + 1) // 2, (L.nj + 1) // 2, (L.nk + 1) // 2


def new_level(ni, nj, nk):
    C = Level()
    C.ni, C.nj, C.nk = ni, nj, nk
    C.shape = (nk + 3, nj + 3, ni + 3)
    C.sy, C.sz = ni + 3, (ni + 3) * (nj + 3)
    C.T = C.shape[0] * C.shape[1] * C.shape[2]
    kk, jj, ii = np.meshgrid(np.arange(-1, nk + 2), np.arange(-1, nj + 2), np.arange(-1, ni + 2), indexing="ij")
    C.interior = ((ii >= 1) & (ii < ni) & (jj >= 1) & (jj < nj) & (kk >= 1) & (kk < nk)).ravel()
    C.ijk = (ii.ravel(), jj.ravel(), kk.ravel())
    return C


def fetch(L, f, i, j, k):
    """f at fine indices (arrays), 0 outside [0,n] box"""
    ok = (i >= 0) & (j >= 0) & (k >= 0) & (i <= L.ni) & (j <= L.nj) & (k <= L.nk)
    idx = (np.clip(i, -1, L.ni + 1) + 1) + L.sy * (np.clip(j, -1, L.nj + 1) + 1) + L.sz * (np.clip(k, -1, L.nk + 1) + 1)
    return np.where(ok, f[idx], 0)


def average(F, C, f, nx, ny, nz):
    I, J, K = C.ijk
    acc = np.zeros(C.T)
    for a in range(-1 if nx else 0, 2):
        for b in range(-1 if ny else 0, 2):
            for c in range(-1 if nz else 0, 2):
                w = (0.5 if (nx and a != 0) else 1.0) * (0.5 if (ny and b != 0) else 1.0) * (0.5 if (nz and c != 0) else 1.0)
                acc += w * fetch(F, f, 2 * I + a, 2 * J + b, 2 * K + c)
    return 0.125 * acc


def coarsen_rediscretize(F, minvol=0.02):
    """the scheme of csrc/vmg.h (k_vmg_coarsen_coefs / classify / coarsen_rows)"""
    C = new_level(*coarse_dims(F))
    C.cc = 0.25 * average(F, C, F.cc, False, False, False)
    C.cu = 0.25 * average(F, C, F.cu, False, True, True)
    C.cv = 0.25 * average(F, C, F.cv, True, False, True)
    C.cw = 0.25 * average(F, C, F.cw, True, True, False)
    C.vol = [average(F, C, F.vol[0], True, False, False), average(F, C, F.vol[1], False, True, False),
             average(F, C, F.vol[2], False, False, True)]
    I, J, K = C.ijk
    C.unk, C.wall = [], []
    for m in range(3):
        anyc = np.zeros(C.T, bool)
        wall = np.zeros(C.T, bool)
        for a in range(-1 if m == 0 else 0, 2):
            for b in range(-1 if m == 1 else 0, 2):
                for c in range(-1 if m == 2 else 0, 2):
                    anyc |= fetch(F, F.unk[m], 2 * I + a, 2 * J + b, 2 * K + c).astype(bool)
                    wall |= fetch(F, F.wall[m], 2 * I + a, 2 * J + b, 2 * K + c).astype(bool)
        unk = C.interior & anyc & (C.vol[m] >= minvol)
        C.unk.append(unk)
        C.wall.append(~unk & (wall | ~C.interior))
    return C


def prolongation(F, C, renorm=True, transverse="linear"):
    """sparse P (fine unknowns x coarse unknowns), weights of csrc/vmg.h vmg_parents + renormalisation"""
    rows, cols, vals = [], [], []
    for m in range(3):
        fid = np.arange(F.T)[F.unk[m]]
        fi, fj, fk = F.ijk[0][fid], F.ijk[1][fid], F.ijk[2][fid]
        r = F.num[m * F.T + fid]

        def parents(own, n):
            p0 = n >> 1
            if own:
                odd = (n & 1) == 1
                return p0, np.where(odd, p0 + 1, p0), np.where(odd, 0.5, 1.0), np.where(odd, 0.5, 0.0)
            p1 = np.where((n & 1) == 1, p0 + 1, p0 - 1)
            if transverse == "const":
                return p0, p1, np.full(n.shape, 1.0), np.full(n.shape, 0.0)
            return p0, p1, np.full(n.shape, 0.75), np.full(n.shape, 0.25)

        pi = parents(m == 0, fi); pj = parents(m == 1, fj); pk = parents(m == 2, fk)
        tot = np.zeros(len(fid))
        ent = []
        for a in range(2):
            for b in range(2):
                for c in range(2):
                    w = pi[2 + a] * pj[2 + b] * pk[2 + c]
                    I, J, K = pi[a], pj[b], pk[c]
                    ok = (w > 0) & (I >= 0) & (J >= 0) & (K >= 0) & (I <= C.ni) & (J <= C.nj) & (K <= C.nk)
                    cid = (np.clip(I, -1, C.ni + 1) + 1) + C.sy * (np.clip(J, -1, C.nj + 1) + 1) + C.sz * (np.clip(K, -1, C.nk + 1) + 1)
                    cn = C.num[m * C.T + cid]
                    ok &= cn >= 0
                    tot += np.where(ok, w, 0)
                    ent.append((ok, cn, w))
        for ok, cn, w in ent:
            ww = w / np.where(tot > 0, tot, 1) if renorm else w
            rows.append(r[ok]); cols.append(cn[ok]); vals.append(ww[ok])
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(F.nunk, C.nunk))
    return P


def number(L):
    unk = np.concatenate(L.unk)
    num = -np.ones(3 * L.T, np.int64)
    num[unk] = np.arange(unk.sum())
    L.num, L.nunk = num, int(unk.sum())


class MG:
    def __init__(self, L0, A0, nlev=6, galerkin=False, minvol=0.02, pre=2, omega=0.5, smoother="jacobi", cheb_deg=2,
                 coarse_exact=True, renorm=True, verbose=True, rscale=0.125, transverse="linear", alpha=1.0):
        self.alpha = alpha
        self.A, self.P, self.R, self.Dinv, self.lam = [A0], [], [], [], []
        self.pre, self.omega, self.smoother, self.cheb_deg = pre, omega, smoother, cheb_deg
        L = L0
        self.levels = [L0]
        for l in range(1, nlev):
            if min(L.ni, L.nj, L.nk) <= 4:
                break
            C = coarsen_rediscretize(L, minvol)
            number(C)
            if C.nunk < 10:
                break
            P = prolongation(L, C, renorm, transverse)
            R = (P.T * (rscale if not galerkin else 1.0)).tocsr()
            if galerkin:
                Ac = (P.T @ self.A[-1] @ P).tocsr()
                # drop coarse unknowns with empty rows
            else:
                Ac = assemble(C)
            self.P.append(P); self.R.append(R); self.A.append(Ac); self.levels.append(C)
            if verbose:
                print("  level %d: %dx%dx%d unknowns %d nnz/row %.1f" % (l, C.ni, C.nj, C.nk, C.nunk, Ac.nnz / max(1, C.nunk)))
            L = C
        for A in self.A:
            d = A.diagonal()
            d = np.where(d > 0, d, 1.0)
            self.Dinv.append(1.0 / d)
        if smoother == "cheb":
            for A, Di in zip(self.A, self.Dinv):
                # power iteration for lambda_max(D^-1 A)
                v = np.random.default_rng(0).standard_normal(A.shape[0])
                for _ in range(20):
                    v = Di * (A @ v)
                    lam = np.linalg.norm(v)
                    v /= lam
                self.lam.append(1.1 * lam)
        self.coarse_exact = coarse_exact
        if coarse_exact:
            self.lu = spla.splu(self.A[-1].tocsc())

    def smooth(self, l, x, b, n):
        A, Di = self.A[l], self.Dinv[l]
        if self.smoother == "l1" :
            # level 0: damped Jacobi; explicit levels: l1-Jacobi (1 / sum_j |a_ij|), unconditionally stable
            if l == 0:
                w = self.omega * Di
            else:
                if not hasattr(self, "l1"): self.l1 = {}
                if l not in self.l1: self.l1[l] = 1.0 / np.asarray(abs(A).sum(1)).ravel()
                w = np.minimum(self.omega * Di, self.l1_scale * self.l1[l])
            for _ in range(n):
                x = (w * b) if x is None else x + w * (b - A @ x)
            return x
        if self.smoother == "jacobi":
            for _ in range(n):
                x = (self.omega * Di * b) if x is None else x + self.omega * Di * (b - A @ x)
            return x
        # Chebyshev (degree n) on D^-1 A, interval [lam/ratio, lam]
        lmax = self.lam[l]; lmin = lmax / self.cheb_ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = b if x is None else b - A @ x
        d = Di * r / theta
        x = d.copy() if x is None else x + d
        for _ in range(n - 1):
            r = b - A @ x
            rho_new = 1.0 / (2 * sigma - rho)
            d = rho_new * rho * d + 2 * rho_new / delta * (Di * r)
            x = x + d
            rho = rho_new
        return x

    cheb_ratio = 4.0
    l1_scale = 1.0

    def vcycle(self, r, l=0):
        if l == len(self.A) - 1:
            if self.coarse_exact:
                return self.lu.solve(r)
            return self.smooth(l, None, r, 30)
        npre = self.pre_levels[l] if getattr(self, "pre_levels", None) else self.pre
        x = self.smooth(l, None, r, npre)
        res = r - self.A[l] @ x
        xc = self.vcycle(self.R[l] @ res, l + 1)
        x = x + self.alpha * (self.P[l] @ xc)
        x = self.smooth(l, x, r, npre)
        return x


if __name__ == "__main__":
    path = sys.argv[1]
    st = np.load(path)
    t0 = time.time()
    L0 = make_level0(st)
    A, b = assemble(L0, with_rhs=True)
    print("n", int(st["n"]), "dt", float(st["dt"]), "unknowns", L0.nunk, "nnz/row %.1f" % (A.nnz / L0.nunk), "assemble %.1fs" % (time.time() - t0))
    print("symmetry err", abs(A - A.T).max(), "max|b|", np.abs(b).max())
    d = A.diagonal()
    t0 = time.time()
    x, it = pcg(A, b, lambda r: r / d)
    print("jacobi-PCG iterations", it, "%.1fs" % (time.time() - t0))
    np.save("/tmp/visc_x.npy", x)
