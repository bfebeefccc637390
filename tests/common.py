"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MESHES = os.path.join(ROOT, "tests", "data", "meshes")
EMU_DIR = os.path.join(ROOT, "tests", "cpu_emu")
EMU_LIB = os.path.join(EMU_DIR, "libflip_emu.so")


def read_ply(path):
    """Whole-file parse of the binary little-endian PLY the reference writes
    (src/trianglemesh.cpp:192-225, 306-332); also handles files < 2048 B, which the
    reference's own loader rejects."""
    data = open(path, "rb").read()
    hend = data.index(b"end_header\n") + len(b"end_header\n")
    hdr = data[:hend].decode().split("\n")
    nv = int([l for l in hdr if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in hdr if l.startswith("element face")][0].split()[-1])
    v = np.frombuffer(data, np.float32, nv * 3, hend).reshape(nv, 3).copy()
    rec = np.dtype([("n", "u1"), ("i", "<i4", 3)])
    f = np.frombuffer(data, rec, nf, hend + nv * 12)
    assert (f["n"] == 3).all()
    return v, f["i"].astype(np.int32).copy()


def mesh(name):
    return read_ply(os.path.join(MESHES, name + ".ply"))


def build_emu():
    """Build the CPU-emulation library (dev/test tooling, tests/cpu_emu/cuda_emu.h)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", EMU_DIR])
    return EMU_LIB


def emu_library():
    from flipviscosity3d_b200 import _lib
    build_emu()
    return _lib.load_library(EMU_LIB)


def make_ref_scene(n, liquid="stanford_bunny", boundary="sphere_large", viscosity=5.0, seed=1):
    """Reference scene: bunny (or other mesh) in the inverted sphere, as src/main.cpp:50-80."""
    from oracle import refsim
    refsim.srand(seed)
    r = refsim.RefSim(n, n, n, 1.0 / n)
    if boundary:
        v, f = mesh(boundary)
        r.add_boundary(v, f, True)
    v, f = mesh(liquid)
    r.add_liquid(v, f)
    r.set_viscosity(viscosity)
    r.set_gravity(0.0, -9.81, 0.0)
    return r


def mirror_to(sim, ref, viscosity=5.0):
    """Load the reference's current scene state into a FlipSim."""
    sim.set_solid_sdf(ref.get_solid_sdf())
    sim.set_particles(ref.get_particles())
    sim.set_viscosity(viscosity)
    sim.set_gravity(0.0, -9.81, 0.0)


def maxdiff(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0
