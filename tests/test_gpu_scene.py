"""Scene construction on the device (csrc/scene.cu) against the oracle and the host layer:
  * the library's restatement of glibc rand() reproduces libc's sequence;
  * the mesh SDF equals the reference's wherever the true distance is <= 3 cells (bit for bit) and has the same sign
    everywhere; farther out it is the exact distance where the reference has its breadth-first approximation (>= exact);
  * seeding through flip_add_liquid_mesh gives the reference's particles, bit for bit and in the same order;
  * a scene built on the device evolves exactly like the same scene built by the reference."""
import ctypes as C

import numpy as np
import pytest

import common
from flipviscosity3d_b200 import FlipSim, fields as F

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "cuda_lib")


def test_rand_restatement_matches_libc(lib):
    libc = C.CDLL("libc.so.6")
    for seed in (1, 12345):
        libc.srand(seed); lib.flip_srand(seed)
        a = [libc.rand() for _ in range(5000)]
        b = [lib.flip_rand() for _ in range(5000)]
        assert a == b
    lib.flip_srand(1)


@pytest.mark.parametrize("name,n", [("stanford_bunny", 24), ("sphere_large", 20), ("cube", 16)])
def test_mesh_sdf_exact_band_and_sign(lib, oracle, name, n):
    v, f = common.mesh(name)
    sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
    a = sim.mesh_sdf(v, f)
    b = oracle.mesh_sdf(n, n, n, 1.0 / n, v, f)
    sim.close()
    assert np.array_equal(a < 0, b < 0)                       # same inside / outside everywhere
    near = np.abs(a) <= 3.0 / n
    assert near.sum() > 100
    assert np.array_equal(a[near], b[near])                   # bit-identical within 3 cells of the surface
    assert (np.abs(a) <= np.abs(b) * (1 + 1e-6)).all()        # farther out: exact <= the reference's approximation


@pytest.mark.parametrize("n", [16, 24])
def test_device_scene_matches_reference(lib, oracle, n):
    ref = common.make_ref_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
    sim.srand(1)
    sim.reset_boundary()
    sim.add_boundary(*common.mesh("sphere_large"), inverted=True)
    added = sim.add_liquid(*common.mesh("stanford_bunny"))
    assert added == ref.num_particles() == sim.num_particles()
    assert np.array_equal(sim.get_particles(), ref.get_particles())
    a, b = sim.get_field(F.F_SOLID_SDF), ref.get_solid_sdf()
    assert np.array_equal(a < 0, b < 0)
    near = np.abs(b) <= 2.0 / n
    assert np.array_equal(a[near], b[near])
    ref.compute_weights()
    for x, y in zip(sim.get_weights(), ref.get_weights()):
        assert np.array_equal(x, y)
    sim.set_viscosity(5.0)
    for _ in range(3):
        assert ref.advance(0.01) == sim.advance(0.01)
    p, q = sim.get_particles(), ref.get_particles()
    assert common.maxdiff(p[:, :3], q[:, :3]) <= 5e-6
    sim.close()


def test_device_and_host_scene_evolve_identically(lib, oracle):
    """same library, scene from the device path vs from the reference's arrays: bit-identical after 3 frames"""
    n = 16
    ref = common.make_ref_scene(n, liquid="cube", boundary=None, viscosity=2.0)
    a = FlipSim(n, n, n, 1.0 / n, lib=lib)
    a.srand(1); a.reset_boundary(); a.add_liquid(*common.mesh("cube")); a.set_viscosity(2.0)
    b = FlipSim(n, n, n, 1.0 / n, lib=lib)
    common.mirror_to(b, ref, 2.0)
    for _ in range(3):
        a.advance(0.01); b.advance(0.01)
    assert np.array_equal(a.get_particles(), b.get_particles())
    a.close(); b.close()


def test_mesh_outside_domain_is_rejected(lib):
    from flipviscosity3d_b200 import FlipError
    sim = FlipSim(16, 16, 16, 1.0 / 16, lib=lib)
    v, f = common.mesh("cube")
    with pytest.raises(FlipError):
        sim.add_liquid(v + 0.6, f)
    with pytest.raises(FlipError):
        sim.add_boundary(v - 0.5, f)
    sim.close()
