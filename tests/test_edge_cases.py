"""Edge cases of the substep against the compiled reference, on the CPU-emulated kernels
(and on the device with -m gpu): non-cubic grids whose sizes are not multiples of the 8-cell solver
block, a variable viscosity grid, CFL-split frames, zero viscosity, empty particle sets, parameter
validation."""
import numpy as np
import pytest

import common
import parity_checks as pc
from flipviscosity3d_b200 import FlipError, FlipSim, fields as F, scene as hs

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "cuda_lib")


def _pair_noncubic(lib, oracle, ni, nj, nk, viscosity):
    """bunny in the domain box on an ni x nj x nk grid (dx = 1/max)"""
    dx = 1.0 / max(ni, nj, nk)
    oracle.srand(1)
    ref = oracle.RefSim(ni, nj, nk, dx)
    v, f = common.mesh("cube")
    # keep the mesh inside the smaller domain: scale the unit-cube meshes into the box
    scale = np.array([ni * dx, nj * dx, nk * dx], np.float32)
    ref.add_liquid(v * scale, f)
    ref.set_viscosity(viscosity)
    p = ref.get_particles()
    rng = np.random.default_rng(3)
    p[:, 3:] = (rng.standard_normal((len(p), 3)) * 0.3).astype(np.float32)
    ref.set_particles(p)
    sim = FlipSim(ni, nj, nk, dx, lib=lib)
    sim.set_solid_sdf(ref.get_solid_sdf()); sim.set_particles(p); sim.set_viscosity(viscosity)
    return sim, ref


@pytest.mark.parametrize("dims", [(20, 24, 28), (17, 13, 22)], ids=["20x24x28", "17x13x22"])
def test_noncubic_grid_frames(lib, oracle, dims):
    sim, ref = _pair_noncubic(lib, oracle, *dims, viscosity=1.5)
    assert sim.num_particles() > 500
    for _ in range(2):
        assert ref.advance(0.01) == sim.advance(0.01)
    a, b = sim.get_particles(), ref.get_particles()
    assert common.maxdiff(a[:, :3], b[:, :3]) <= 5e-6
    assert common.maxdiff(a[:, 3:], b[:, 3:]) <= 1e-4


def test_variable_viscosity_grid(lib, oracle):
    """setViscosity(Array3d<float>&) (src/fluidsimulation.cpp:110-124): a viscosity ramp in y."""
    n = 20
    sim, ref = pc.build_pair(lib, oracle, n=n, liquid="cube", boundary=None, viscosity=1.0)
    y = np.linspace(0.0, 4.0, n + 1, dtype=np.float32)
    grid = np.broadcast_to(y[None, :, None], (n + 1, n + 1, n + 1)).copy()
    sim.set_viscosity(grid); ref.set_viscosity(grid)
    pc.prepare_mid_substep(sim, ref)
    pc.sync_grid_state(sim, ref)
    sim.set_param("viscosity_tol", 1e-10)
    sim.apply_viscosity(pc.DT)
    info = ref.apply_viscosity(pc.DT, tol=1e-10, maxit=20000)
    assert info["wrote"] == 1 and sim.stats()["viscosity_converged"] == 1
    masks = pc.fluid_border_masks(ref.get_liquid_sdf())
    for a, b, m in zip(sim.get_mac(), ref.get_mac(), masks):
        # the edge-averaged viscosity is summed in one fixed order here, in two orders (U vs V rows)
        # in the reference: identical for a uniform field, last-bit different for a varying one
        assert np.abs(a - b)[m].max() <= 5e-6 * max(1.0, np.abs(b).max())
    with pytest.raises(FlipError):
        sim.set_viscosity(-grid - 1.0)            # FLUIDSIM_ASSERT(v >= 0) in the reference


def test_cfl_splits_the_frame_like_the_reference(lib, oracle):
    """Fast particles make _cfl() < dt: both sides must take the same number of substeps."""
    sim, ref = pc.build_pair(lib, oracle, n=16, liquid="cube", boundary=None, viscosity=0.0, random_velocity=False)
    p = ref.get_particles()
    p[:, 3] = 40.0                                   # 40 m/s along x: CFL step = 5*dx/40 << 0.01
    ref.set_particles(p); sim.set_particles(p)
    ref.substep(0.001); sim.substep(0.001)          # builds a grid velocity for _cfl()
    assert sim.cfl() == ref.cfl() and sim.cfl() < 0.01
    n1, n2 = ref.advance(0.01), sim.advance(0.01)
    assert n1 == n2 and n1 > 1
    a, b = sim.get_particles(), ref.get_particles()
    assert common.maxdiff(a[:, :3], b[:, :3]) <= 2e-5


def test_zero_viscosity_skips_the_solve(lib, oracle):
    sim, ref = pc.build_pair(lib, oracle, n=16, liquid="cube", boundary=None, viscosity=0.0)
    assert ref.advance(0.01) == sim.advance(0.01)
    st = sim.stats()
    assert st["viscosity_iterations"] == 0 and st["viscosity_applied"] == 0
    assert common.maxdiff(sim.get_particles(), ref.get_particles()) <= 1e-5


def test_empty_particle_set_and_bad_arguments(lib):
    sim = FlipSim(16, 16, 16, 1.0 / 16, lib=lib)
    sim.set_particles(np.zeros((0, 6), np.float32))
    assert sim.advance(0.01) == 1                    # first substep clamps to the frame (max|u| = 0)
    assert sim.num_particles() == 0
    with pytest.raises(FlipError):
        sim.set_param("no_such_parameter", 1.0)
    with pytest.raises(FlipError):
        sim.get_field(99)
    with pytest.raises(FlipError):
        FlipSim(2, 16, 16, 1.0 / 16, lib=lib)        # grids smaller than 4 cells are rejected


def test_viscosity_cap_accept_and_fail_branches(lib, oracle):
    """ViscositySolver::_solveLinearSystem (src/viscositysolver.cpp:676-689): a solve that stops at its iteration
    cap is still ACCEPTED when the residual is below the acceptable tolerance (10.0), otherwise it FAILS and
    FluidSimulation::_applyViscosity ignores the result (src/fluidsimulation.cpp:194-195): the field stays untouched.
    Same branches on the reference (raised / lowered through the harness) and on the library."""
    sim, ref = pc.build_pair(lib, oracle, n=16, liquid="cube", boundary=None, viscosity=4.0)
    pc.prepare_mid_substep(sim, ref)
    pc.sync_grid_state(sim, ref)
    before = [a.copy() for a in sim.get_mac()]
    # (1) cap hit, residual < acceptable tolerance -> accepted, written back, not converged
    sim.set_param("maxit_scale", 1)
    sim.set_param("viscosity_maxit", 2)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    assert st["viscosity_iterations"] == 2 and st["viscosity_converged"] == 0 and st["viscosity_applied"] == 1
    assert st["viscosity_residual"] < 10.0
    assert any(not np.array_equal(a, b) for a, b in zip(sim.get_mac(), before))
    info = ref.apply_viscosity(pc.DT, tol=1e-6, maxit=2)
    assert info["iters"] == 2 and info["resid"] > 1e-6 * 0 and info["ok"] == 1 and info["wrote"] == 1   # same branch: accepted at the cap
    # (2) cap hit, residual above the acceptable tolerance -> FAILED, field untouched
    pc.sync_grid_state(sim, ref)   # the reference wrote its truncated iterate: reload the common pre-solve state
    before = [a.copy() for a in sim.get_mac()]
    sim.set_param("viscosity_accept", 1e-30)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    assert st["viscosity_iterations"] == 2 and st["viscosity_converged"] == 0 and st["viscosity_applied"] == 0
    for a, b in zip(sim.get_mac(), before):
        assert np.array_equal(a, b)
    # (3) zero right-hand side (fluid at rest, no solid motion): success with 0 iterations (pcgsolver.h:254-258)
    sim.set_param("viscosity_accept", 10.0); sim.set_param("viscosity_maxit", 700); sim.set_param("maxit_scale", 40)
    zero = [np.zeros_like(a) for a in before]
    sim.set_mac(*zero)
    sim.apply_viscosity(pc.DT)
    st = sim.stats()
    assert st["viscosity_iterations"] == 0 and st["viscosity_converged"] == 1 and st["viscosity_applied"] == 1


def test_block_lists_equal_dense_sweeps(lib):
    """The grid stages run over the blocks near the liquid (fields.cu) instead of the whole grid: same bits in every
    field as with all blocks listed, while a small blob falls through a mostly empty 40^3 domain (the list follows it
    and the blocks it leaves go back to their defaults)."""
    n = 40
    dx = 1.0 / n
    c = (np.arange(n + 1) * dx).astype(np.float64)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    lo, hi = 3 * dx + 1e-6, 1 - 3 * dx - 1e-6
    phi = np.minimum.reduce([x - lo, hi - x, y - lo, hi - y, z - lo, hi - z]).astype(np.float32)
    rng = np.random.default_rng(5)
    cells = np.stack(np.meshgrid(np.arange(5, 11), np.arange(26, 33), np.arange(6, 12), indexing="ij"), -1).reshape(-1, 3)
    pos = (np.repeat(cells, 6, 0) + rng.random((len(cells) * 6, 3))) * dx
    p = np.zeros((len(pos), 6), np.float32)
    p[:, :3] = pos
    p[:, 3] = 1.5       # drifts in +x while it falls: blocks enter and leave the list
    sims = []
    for use in (1, 0):
        sim = FlipSim(n, n, n, dx, lib=lib)
        sim.set_param("use_block_lists", use)
        sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(0.5)
        sims.append(sim)
    for frame in range(6):
        for sim in sims:
            sim.advance(0.02)
        a, b = sims
        assert np.array_equal(a.get_particles(), b.get_particles()), frame
        for f in (F.F_LIQUID_SDF, F.F_U, F.F_V, F.F_W, F.F_SAVED_U, F.F_SAVED_V, F.F_SAVED_W, F.F_PRESSURE, F.F_VOL_CENTER,
                  F.F_VOL_U, F.F_VOL_EDGE_W):
            assert np.array_equal(a.get_field(f), b.get_field(f)), (frame, f)
        for u, v in zip(a.get_valid(), b.get_valid()):
            assert np.array_equal(u, v)
    st = sims[0].stats()
    assert st["viscosity_converged"] == 1 and st["pressure_converged"] == 1
    for sim in sims:
        sim.close()


def test_async_position_export(lib):
    """flip_get_positions_async / flip_output_wait: the positions the reference's exporters write (src/main.cpp:14-40), in
    the caller's order, captured at the moment of the call while later substeps run"""
    from __graft_entry__ import _analytic_scene
    n = 16
    phi, p = _analytic_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(1.0)
    bufs = [np.zeros((len(p), 3), np.float32) for _ in range(3)]
    want = []
    for f in range(3):
        sim.advance(0.01)
        want.append(sim.get_particles()[:, :3].copy())
        assert sim.get_positions_async(bufs[f]) == len(p)     # three exports back to back: the third re-uses a staging buffer
    sim.advance(0.01)                                          # keeps simulating while the copies drain
    sim.output_wait()
    for f in range(3):
        assert np.array_equal(bufs[f], want[f])
    sim.close()
