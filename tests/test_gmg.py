"""Galerkin multigrid preconditioner of the viscosity CG (csrc/gmg.h).

The preconditioner has no counterpart in the reference (which uses MIC(0), src/pcgsolver/pcgsolver.h:62-214); what
must hold is that it changes NOTHING but the iteration count: same operator, same right-hand side, same stopping
rule, same converged velocities as the diagonal-preconditioned solve and as the reference.  Checked on the
CPU-emulation build (kernel logic, no GPU) and, marked `gpu`, on the sm_100a library.
"""
import ctypes as C

import numpy as np
import pytest

import common
import parity_checks as pc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "cuda_lib")


def _solve(sim, ref, precond, tol=1e-6):
    pc.sync_grid_state(sim, ref)
    sim.set_param("viscosity_precond", precond)
    sim.set_param("viscosity_tol", tol)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    assert st["viscosity_converged"] == 1 and st["viscosity_applied"] == 1
    return [f.copy() for f in sim.get_mac()], st


def test_multigrid_changes_only_the_iteration_count(lib, oracle):
    sim, ref = pc.build_pair(lib, oracle, n=32)
    pc.prepare_mid_substep(sim, ref)
    jac, st_j = _solve(sim, ref, 0, tol=1e-9)
    mg, st_m = _solve(sim, ref, 2, tol=1e-9)
    assert st_m["viscosity_unknowns"] == st_j["viscosity_unknowns"]
    # compared on the faces that border a fluid cell, like the reference parity check: faces whose control volume
    # holds no liquid form a singular block of the reference's system, arbitrary in any solver (parity_checks.py)
    masks = pc.fluid_border_masks(ref.get_liquid_sdf())
    for a, b, m in zip(mg, jac, masks):
        assert np.array_equal(a != 0, b != 0)
        assert np.abs(a - b)[m].max() <= 2e-6 * max(1.0, np.abs(b).max())
    # 32^3: 250+ diagonal-PCG iterations, ~25 with the V-cycle
    assert st_m["viscosity_iterations"] * 5 <= st_j["viscosity_iterations"]
    assert st_m["viscosity_iterations"] <= 45


def test_multigrid_solve_matches_reference(lib, oracle):
    """The default path (multigrid) against the reference's own converged solve."""
    sim, ref = pc.build_pair(lib, oracle, n=24)
    pc.prepare_mid_substep(sim, ref)
    sim.set_param("viscosity_precond", 2)
    pc.check_viscosity(sim, ref)


def test_stiff_system_keeps_its_mass_term(lib, oracle):
    """dt*mu/dx^2 ~ 5e4 (the 256^3 regime at 32^3): the six row factors add up to ~1e5, where one fp32 ulp (0.008) is
    larger than many face volumes.  With the diagonal formed in fp32 the matrix loses positive definiteness and CG
    wanders for hundreds of iterations; the difference-form operator (viscosity.cu, k_visc_apply) keeps the mass
    term exact and the solve stays short."""
    sim, ref = pc.build_pair(lib, oracle, n=32, viscosity=5000.0)
    pc.prepare_mid_substep(sim, ref)
    _, st = _solve(sim, ref, 2)
    assert st["viscosity_iterations"] <= 80
    _, st0 = _solve(sim, ref, 0)
    assert st0["viscosity_iterations"] >= 4 * st["viscosity_iterations"]


def _slot_table(m):
    """(column component, di, dj, dk) of the 235 slots of a row of component m (gmg.h, gmg_window)."""
    out = []
    for mp in range(3):
        lo, nn = [], []
        for a in range(3):
            if m == mp:
                lo.append(-1 if a == m else -2); nn.append(3 if a == m else 5)
            elif a == m:
                lo.append(-2); nn.append(4)
            elif a == mp:
                lo.append(-1); nn.append(4)
            else:
                lo.append(-2); nn.append(5)
        for dk in range(lo[2], lo[2] + nn[2]):
            for dj in range(lo[1], lo[1] + nn[1]):
                for di in range(lo[0], lo[0] + nn[0]):
                    out.append((mp, di, dj, dk))
    return out


def _level_matrix(lib, sim, level):
    import scipy.sparse as sp
    lib.flip_debug_gmg_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    info = (C.c_int * 10)()
    if lib.flip_debug_gmg_level(sim.h, level, info, None, None, None) != 0:
        return None
    ni, nj, nk, ax, ay, az, T, nrows, nlev, stride = list(info)
    rows = np.zeros(nrows, np.int32); S = np.zeros((nrows, stride), np.float32); diag = np.zeros(3 * T, np.float32)
    assert lib.flip_debug_gmg_level(sim.h, level, info, rows.ctypes.data, S.ctypes.data, diag.ctypes.data) == 0
    assert len(_slot_table(0)) == 235 and stride == 240
    assert (S[:, 235:] == 0).all()                     # padding slots
    m, idd = rows // T, rows % T
    rowmap = -np.ones(3 * T, np.int64); rowmap[rows] = np.arange(nrows)
    R, Cc, V = [], [], []
    for mm in range(3):
        sel = np.nonzero(m == mm)[0]
        for slot, (mp, di, dj, dk) in enumerate(_slot_table(mm)):
            v = S[sel, slot]
            nz = v != 0
            col = rowmap[mp * T + idd[sel][nz] + di + dj * ax + dk * ax * ay]
            assert (col >= 0).all(), "entry that points to a face that is not an unknown"
            R.append(sel[nz]); Cc.append(col); V.append(v[nz].astype(np.float64))
    A = sp.csr_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(Cc))), shape=(nrows, nrows))
    return A, diag[rows]


def test_galerkin_levels_are_symmetric_positive_and_safely_smoothed(lib, oracle):
    sim, ref = pc.build_pair(lib, oracle, n=32)
    pc.prepare_mid_substep(sim, ref)
    _solve(sim, ref, 2)
    level, seen = 1, 0
    while True:
        got = _level_matrix(lib, sim, level)
        if got is None:
            break
        A, dense_diag = got
        d = A.diagonal()
        assert (d > 0).all() and np.allclose(d, dense_diag, rtol=1e-6)
        assert abs(A - A.T).max() <= 2e-6 * abs(A).max()                 # P^T A P of a symmetric A
        # positive definite: a few random Rayleigh quotients and the rigid translations (x^T A x = coarse mass > 0)
        rng = np.random.default_rng(level)
        for _ in range(4):
            x = rng.standard_normal(A.shape[0])
            assert x @ (A @ x) > 0
        assert np.ones(A.shape[0]) @ (A @ np.ones(A.shape[0])) > 0
        # smoothing weights w = min(0.5/a_ii, 1.6/sum|a_ij|): lambda_max(W A) <= 1.6 < 2 (power iteration)
        w = np.minimum(0.5 / d, 1.6 / np.asarray(abs(A).sum(1)).ravel())
        v = rng.standard_normal(A.shape[0])
        for _ in range(60):
            v = w * (A @ v)
            lam = np.linalg.norm(v)
            v /= lam
        assert lam <= 1.6 + 1e-3
        seen += 1
        level += 1
    assert seen >= 2


@pytest.mark.gpu
def test_time_kernel_reports_the_level1_sweep(cuda_lib, oracle):
    sim, ref = pc.build_pair(cuda_lib, oracle, n=32)
    pc.prepare_mid_substep(sim, ref)
    _solve(sim, ref, 2)
    ms, nbytes = sim.time_kernel("gmg_sweep_l1", 5)
    # rows x (stored slots + 5) x 4 B: 160 stored slots with the compact rows of level 1 (default), else 235
    assert ms > 0 and nbytes > 0 and (nbytes % (165 * 4) == 0 or nbytes % (240 * 4) == 0)
    ms2, nb2 = sim.time_kernel("visc_apply", 5)
    assert ms2 > 0 and nb2 == 32 * sim.stats()["viscosity_unknowns"]
    with pytest.raises(Exception):
        sim.time_kernel("no_such_kernel", 1)


def test_compact_rows_of_the_first_explicit_level(lib, oracle):
    """The sweeps of level 1 read compact rows (csrc/gmg.h k_gmg_compact_rows): every non-zero of a full 235-slot row must
    be in its 160-slot compact row, at the slot the compact table names, and switching the compact rows off must change
    nothing but rounding (same iteration count, velocities equal to ~1e-7)."""
    dll = lib.dll if hasattr(lib, "dll") else lib
    dll.flip_debug_gmg_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    dll.flip_debug_gmg_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sim, ref = pc.build_pair(lib, oracle, n=32)
    pc.prepare_mid_substep(sim, ref)
    out = {}
    for flag in (0, 1):
        sim.set_param("mg_compact", flag)
        out[flag] = _solve(sim, ref, 2)
        meta = (C.c_int * 4)()
        assert dll.flip_debug_gmg_compact(sim.h, meta, None, None) == 0
        assert meta[0] == flag
    assert out[0][1]["viscosity_iterations"] == out[1][1]["viscosity_iterations"]
    for a, b in zip(out[0][0], out[1][0]):
        assert np.abs(a - b).max() <= 1e-6
    info = (C.c_int * 10)()
    assert dll.flip_debug_gmg_level(sim.h, 1, info, None, None, None) == 0
    T, nrows, stride = info[6], info[7], info[9]
    rows = np.zeros(nrows, np.int32); S = np.zeros((nrows, stride), np.float32); d = np.zeros(3 * T, np.float32)
    assert dll.flip_debug_gmg_level(sim.h, 1, info, rows.ctypes.data, S.ctypes.data, d.ctypes.data) == 0
    Sc = np.zeros((nrows, 160), np.float32); cslot = np.zeros((3, 160), np.int32)
    assert dll.flip_debug_gmg_compact(sim.h, meta, Sc.ctypes.data, cslot.ctypes.data) == 0
    assert max(meta[1], meta[2], meta[3]) <= 160 and min(meta[1], meta[2], meta[3]) > 0
    comp = rows // T
    for m in range(3):
        sl = cslot[m]
        kept = sl >= 0
        assert kept.sum() == meta[1 + m] and np.all(np.diff(sl[kept]) > 0)       # ascending slot order, padding last
        Sm, Scm = S[comp == m], Sc[comp == m]
        assert np.array_equal(Scm[:, kept], Sm[:, sl[kept]]) and not Scm[:, ~kept].any()
        dropped = np.setdiff1d(np.arange(235), sl[kept])
        assert not Sm[:, dropped].any()                                             # nothing non-zero was left behind
