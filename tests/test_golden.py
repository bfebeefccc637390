"""Parity against committed golden fixtures (tests/golden/stage_golden_n16.npz, generated from the
compiled reference by tests/golden/make_golden.py).  Needs neither /root/reference nor oracle/_ref.
Same tolerances as tests/parity_checks.py."""
import os

import numpy as np
import pytest

import common
import parity_checks as pc
from flipviscosity3d_b200 import FlipSim, fields as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage_golden_n16.npz")
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "cuda_lib")


def test_golden_substep_stages(lib):
    g = np.load(GOLDEN)
    n, dt = int(g["n"]), float(g["dt"])
    sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
    sim.set_solid_sdf(g["solid_sdf"])
    sim.set_particles(g["particles0"])
    sim.set_viscosity(float(g["viscosity"]))
    for a, k in zip(sim.get_weights(), ("weight_u", "weight_v", "weight_w")):
        assert np.array_equal(a, g[k])
    sim.update_liquid_sdf()
    assert np.array_equal(sim.get_field(F.F_LIQUID_SDF), g["liquid_sdf"])
    sim.advect_velocity_field()
    for a, k in zip(sim.get_mac(), ("p2g_u", "p2g_v", "p2g_w")):
        assert common.maxdiff(a, g[k]) <= 1e-5 * max(1.0, np.abs(g[k]).max())
    for a, k in zip(sim.get_valid(), ("p2g_valid_u", "p2g_valid_v", "p2g_valid_w")):
        assert np.array_equal(a, g[k])
    # teacher forcing from here on
    sim.set_mac(g["p2g_u"], g["p2g_v"], g["p2g_w"])
    sim.add_body_force(dt)
    for a, k in zip(sim.get_mac(), ("force_u", "force_v", "force_w")):
        assert np.array_equal(a, g[k])
    sim.viscosity_volumes()
    for fid, k in zip(range(F.F_VOL_CENTER, F.F_VOL_EDGE_W + 1), ("c", "u", "v", "w", "eu", "ev", "ew")):
        assert common.maxdiff(sim.get_field(fid), g["vol_" + k]) <= 1e-6
    sim.set_param("viscosity_tol", 1e-10)
    sim.apply_viscosity(dt)
    masks = pc.fluid_border_masks(g["liquid_sdf"])
    for a, k, m in zip(sim.get_mac(), ("visc_u", "visc_v", "visc_w"), masks):
        assert np.array_equal(a != 0, g[k] != 0)
        assert np.abs(a - g[k])[m].max() <= 2e-6 * max(1.0, np.abs(g[k]).max())
    sim.set_mac(g["visc_u"], g["visc_v"], g["visc_w"])
    sim.solve_pressure(dt)
    assert common.maxdiff(sim.get_field(F.F_PRESSURE), g["pressure"]) <= 1e-6 * max(1.0, np.abs(g["pressure"]).max())
    sim.set_field(F.F_PRESSURE, g["pressure"])
    sim.apply_pressure(dt)
    for a, k in zip(sim.get_valid(), ("proj_valid_u", "proj_valid_v", "proj_valid_w")):
        assert np.array_equal(a, g[k])
    sim.extrapolate()
    sim.set_saved_mac(g["p2g_u"], g["p2g_v"], g["p2g_w"])
    sim.constrain()
    for a, k in zip(sim.get_mac(), ("final_u", "final_v", "final_w")):
        assert np.array_equal(a, g[k])
    for a, k in zip(sim.get_saved_mac(), ("saved_u", "saved_v", "saved_w")):
        assert np.array_equal(a, g[k])
    assert np.float32(sim.cfl()) == g["cfl"]
    sim.advect_particles(dt)
    assert np.array_equal(sim.get_particles(), g["particles1"])


def test_c_abi_exports_every_declared_symbol(emu_lib):
    """The header, the python prototypes and the library agree (no compute calls)."""
    import re
    from flipviscosity3d_b200 import _lib
    hdr = open(os.path.join(common.ROOT, "include", "flip_b200.h")).read()
    declared = set(re.findall(r"\b(flip_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    path = _lib.DEFAULT_LIB
    if os.path.exists(path):       # built by __graft_entry__.build(); loads without a GPU
        lib = _lib.load_library(path)
        assert lib.flip_version().startswith(b"flip_b200")
        assert b"sm_100a" in lib.flip_version()


def test_no_cuda_device_fails_loudly():
    """No CPU fallback: without a device flip_create must fail, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from flipviscosity3d_b200 import FlipError, _lib
    if not os.path.exists(_lib.DEFAULT_LIB):
        pytest.skip("library not built")
    with pytest.raises(FlipError):
        FlipSim(16, 16, 16, 1.0 / 16)
