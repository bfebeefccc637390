"""Host-side drop-in layer (C++11 FlipScene / TriangleMesh) against the compiled reference:
mesh SDF, boundary union, rand()-seeded particles and the PLY writer must be bit-identical."""
import os

import numpy as np
import pytest

import common
from flipviscosity3d_b200 import scene as hs


@pytest.mark.parametrize("name,n", [("stanford_bunny", 24), ("sphere_large", 20), ("cube", 16), ("rod", 32)])
def test_mesh_sdf_bit_exact(oracle, name, n):
    v, f = common.mesh(name)
    a = hs.mesh_sdf(n, n, n, 1.0 / n, v, f)
    b = oracle.mesh_sdf(n, n, n, 1.0 / n, v, f)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [16, 32])
def test_scene_matches_reference(oracle, n):
    ref = common.make_ref_scene(n)
    sc = hs.Scene(n, n, n, 1.0 / n)
    sc.add_boundary(*common.mesh("sphere_large"), inverted=True)
    sc.add_liquid(*common.mesh("stanford_bunny"))
    assert np.array_equal(sc.solid_sdf(), ref.get_solid_sdf())
    assert sc.n_particles == ref.num_particles()
    assert np.array_equal(sc.particles(), ref.get_particles())


def test_survey_anchor_particle_count():
    """config 1 (bunny in sphere, 64^3) seeds 73,176 particles."""
    sc = hs.Scene(64, 64, 64, 1.0 / 64)
    sc.add_boundary(*common.mesh("sphere_large"), inverted=True)
    sc.add_liquid(*common.mesh("stanford_bunny"))
    assert sc.n_particles == 73176


def test_ply_writer_byte_identical(oracle, tmp_path):
    ref = common.make_ref_scene(16)
    p = ref.get_particles()
    a, b = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    hs.write_points_ply(a, p[:, :3])
    ref.write_particles_ply(b)
    assert open(a, "rb").read() == open(b, "rb").read()
    v, f = hs.read_ply(a)
    assert np.array_equal(v, p[:, :3]) and len(f) == 0
