"""Stage-by-stage parity checks of the FLIP substep against the oracle (the compiled reference),
with teacher forcing: before each stage the candidate is loaded with the reference's state.
Shared by the CPU-emulation tests (-m "not gpu") and the CUDA tests (-m gpu): the same checks,
a different library behind the same C ABI.

Tolerances (fp32 fields, values O(1)):
  EXACT   bit-identical (integer/min/masks and float expressions evaluated in the same order)
  1e-5    float sums whose order differs from the reference's sequential particle loop
  1e-6    results of CG solves run to a tighter residual than the comparison tolerance
"""
import numpy as np

import common
from flipviscosity3d_b200 import FlipSim, fields as F

DT = 0.01


def build_pair(lib, oracle, n=24, liquid="stanford_bunny", boundary="sphere_large", viscosity=5.0, random_velocity=True,
               shuffle=False):
    ref = common.make_ref_scene(n, liquid=liquid, boundary=boundary, viscosity=viscosity)
    p = ref.get_particles()
    if random_velocity:
        rng = np.random.default_rng(0)
        p[:, 3:] = (rng.standard_normal((len(p), 3)) * 0.5).astype(np.float32)
    if shuffle:
        p = p[np.random.default_rng(1).permutation(len(p))]
    ref.set_particles(p)
    sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
    common.mirror_to(sim, ref, viscosity)
    return sim, ref


def fluid_border_masks(phi):
    fl = phi < 0
    nk, nj, ni = phi.shape
    bu = np.zeros((nk, nj, ni + 1), bool); bu[:, :, :-1] |= fl; bu[:, :, 1:] |= fl
    bv = np.zeros((nk, nj + 1, ni), bool); bv[:, :-1, :] |= fl; bv[:, 1:, :] |= fl
    bw = np.zeros((nk + 1, nj, ni), bool); bw[:-1] |= fl; bw[1:] |= fl
    return bu, bv, bw


def check_static_fields(sim, ref):
    ref.compute_weights()
    for a, b in zip(sim.get_weights(), ref.get_weights()):
        assert np.array_equal(a, b)                       # EXACT


def check_liquid_sdf(sim, ref):
    sim.update_liquid_sdf(); ref.update_liquid_sdf()
    a, b = sim.get_field(F.F_LIQUID_SDF), ref.get_liquid_sdf()
    assert (b < 0).sum() > 0
    assert np.array_equal(a, b)                           # EXACT: min is order independent


def check_p2g(sim, ref, tol=1e-5):
    sim.update_liquid_sdf(); ref.update_liquid_sdf()
    sim.advect_velocity_field(); ref.advect_velocity_field()
    for a, b in zip(sim.get_valid(), ref.get_valid()):
        assert np.array_equal(a, b)
    for a, b in zip(sim.get_mac(), ref.get_mac()):
        assert np.abs(b).max() > 0.1
        assert common.maxdiff(a, b) <= tol * max(1.0, np.abs(b).max())
    for a, b in zip(sim.get_saved_mac(), ref.get_saved_mac()):
        assert common.maxdiff(a, b) <= tol * max(1.0, np.abs(b).max())


def sync_grid_state(sim, ref):
    sim.set_field(F.F_LIQUID_SDF, ref.get_liquid_sdf())
    sim.set_mac(*ref.get_mac())
    sim.set_saved_mac(*ref.get_saved_mac())
    sim.set_valid(*ref.get_valid())


def check_body_force(sim, ref):
    sync_grid_state(sim, ref)
    sim.add_body_force(DT); ref.add_body_force(DT)
    for a, b in zip(sim.get_mac(), ref.get_mac()):
        assert np.array_equal(a, b)                       # EXACT


def check_extrapolate(sim, ref):
    sync_grid_state(sim, ref)
    sim.extrapolate(); ref.extrapolate()
    for a, b in zip(sim.get_mac(), ref.get_mac()):
        assert np.array_equal(a, b)                       # EXACT (Jacobi per layer == reference sweep)


def check_viscosity_volumes(sim, ref):
    sync_grid_state(sim, ref)
    sim.viscosity_volumes()
    rv = ref.viscosity_volumes()
    for fid, b in zip(range(F.F_VOL_CENTER, F.F_VOL_EDGE_W + 1), rv):
        a = sim.get_field(fid)
        assert b.sum() > 0
        assert common.maxdiff(a, b) <= 1e-6               # same float expressions; EXACT in practice


def check_viscosity(sim, ref, tol=2e-6):
    """Both solvers run to a residual far below the comparison tolerance.  Compared on the faces
    bordering a fluid cell: faces whose control volume holds no liquid form a singular block of the
    reference's own system (no mass term), their values are arbitrary in either solver and are
    discarded by _applyPressure (src/fluidsimulation.cpp:658-687)."""
    sync_grid_state(sim, ref)
    sim.set_param("viscosity_tol", 1e-10)
    sim.apply_viscosity(DT)
    info = ref.apply_viscosity(DT, tol=1e-10, maxit=20000)
    st = sim.stats()
    assert info["wrote"] == 1 and st["viscosity_applied"] == 1 and st["viscosity_converged"] == 1
    masks = fluid_border_masks(ref.get_liquid_sdf())
    for a, b, m in zip(sim.get_mac(), ref.get_mac(), masks):
        assert np.array_equal(a != 0, b != 0)             # same unknown set (the rest is cleared to 0)
        assert np.abs(a - b)[m].max() <= tol * max(1.0, np.abs(b).max())
    sim.set_param("viscosity_tol", 1e-6)


def check_pressure(sim, ref, tol=1e-6):
    sync_grid_state(sim, ref)
    ref.compute_weights()
    sim.solve_pressure(DT)
    rp = ref.solve_pressure(DT)
    st = sim.stats()
    assert st["pressure_converged"] == 1
    sp = sim.get_field(F.F_PRESSURE)
    assert np.abs(rp).max() > 0
    assert common.maxdiff(sp, rp) <= tol * max(1.0, np.abs(rp).max())
    # apply the REFERENCE pressure on both sides: isolates _applyPressure
    sim.set_field(F.F_PRESSURE, rp)
    sim.apply_pressure(DT); ref.apply_pressure(DT, rp)
    for a, b in zip(sim.get_valid(), ref.get_valid()):
        assert np.array_equal(a, b)
    for a, b in zip(sim.get_mac(), ref.get_mac()):
        assert np.array_equal(a, b)                       # EXACT
    sim.extrapolate(); ref.extrapolate()
    sim.constrain(); ref.constrain()
    for a, b in zip(sim.get_mac(), ref.get_mac()):
        assert np.array_equal(a, b)
    for a, b in zip(sim.get_saved_mac(), ref.get_saved_mac()):
        assert np.array_equal(a, b)
    assert sim.cfl() == ref.cfl()


def check_advect_particles(sim, ref):
    sync_grid_state(sim, ref)
    sim.set_particles(ref.get_particles())
    sim.advect_particles(DT); ref.advect_particles(DT)
    a, b = sim.get_particles(), ref.get_particles()
    assert np.array_equal(a, b)                           # EXACT: same double-weight trilinear, same order


def prepare_mid_substep(sim, ref):
    """Bring the reference to a state with a developed velocity field (one full substep), then
    redo the grid stages so every stage has non-trivial input."""
    ref.substep(DT)
    ref.update_liquid_sdf(); ref.advect_velocity_field(); ref.add_body_force(DT)
    sim.set_particles(ref.get_particles())
    sync_grid_state(sim, ref)


def check_free_run(sim, ref, frames=3, tol=5e-6):
    for _ in range(frames):
        n1 = ref.advance(DT); n2 = sim.advance(DT)
        assert n1 == n2
    a, b = sim.get_particles(), ref.get_particles()
    assert common.maxdiff(a[:, :3], b[:, :3]) <= tol
    assert common.maxdiff(a[:, 3:], b[:, 3:]) <= 20 * tol
