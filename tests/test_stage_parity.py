"""Stage-by-stage parity of the substep against the compiled reference (oracle/_ref).

Each check runs twice: on the CPU-emulation build of the kernels (no GPU needed; validates the
kernel logic in this container) and, marked `gpu`, on the real sm_100a library through the same
C ABI.  The product path is only the latter.
"""
import pytest

import parity_checks as pc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
SCENES = [
    pytest.param(dict(n=24, liquid="stanford_bunny", boundary="sphere_large", viscosity=5.0), id="bunny24"),
    pytest.param(dict(n=16, liquid="cube", boundary=None, viscosity=2.0), id="cube16"),
]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "cuda_lib")


@pytest.mark.parametrize("scene", SCENES)
def test_static_and_sdf(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, **scene)
    pc.check_static_fields(sim, ref)
    pc.check_liquid_sdf(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("shuffle", [False, True], ids=["ordered", "shuffled"])
def test_p2g(lib, oracle, scene, shuffle):
    sim, ref = pc.build_pair(lib, oracle, shuffle=shuffle, **scene)
    pc.check_p2g(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
def test_grid_stages(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, **scene)
    pc.prepare_mid_substep(sim, ref)
    pc.check_body_force(sim, ref)
    pc.check_extrapolate(sim, ref)
    pc.check_viscosity_volumes(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
def test_viscosity_solve(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, **scene)
    pc.prepare_mid_substep(sim, ref)
    pc.check_viscosity(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
def test_pressure_projection(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, **scene)
    pc.prepare_mid_substep(sim, ref)
    pc.check_pressure(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
def test_advect_particles(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, **scene)
    pc.prepare_mid_substep(sim, ref)
    pc.check_advect_particles(sim, ref)


@pytest.mark.parametrize("scene", SCENES)
def test_free_running_frames(lib, oracle, scene):
    sim, ref = pc.build_pair(lib, oracle, random_velocity=False, **scene)
    pc.check_free_run(sim, ref)
