"""flipviscosity3d_b200 — B200-native FLIP substep (P2G/G2P, variational pressure and viscosity
solves) behind a C ABI.  See DESIGN.md.  The compute path is libflip_b200.so (hand-written CUDA,
sm_100a); this package is the thin Python host binding used by tests and bench.py."""
from ._lib import load_library, default_library, DEFAULT_LIB  # noqa: F401
from .sim import FlipSim, FlipError  # noqa: F401
from . import sim as fields  # noqa: F401
