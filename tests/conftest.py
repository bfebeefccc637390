import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_lib():
    from flipviscosity3d_b200 import _lib
    return _lib.default_library()  # raises if the sm_100a library is missing: no fallback


@pytest.fixture(scope="session")
def emu_lib():
    """CPU-emulation build of the kernels (dev/test tooling, tests/cpu_emu/cuda_emu.h)."""
    import common
    return common.emu_library()


@pytest.fixture(scope="session")
def cuda_lib():
    return _cuda_lib()


@pytest.fixture(scope="session")
def oracle():
    """The compiled reference (oracle/_ref/libflipref.so). Built here when /root/reference is
    present; on the GPU box the prebuilt .so travels with the snapshot."""
    from oracle import refsim
    if not refsim.build():
        pytest.skip("oracle/_ref/libflipref.so not available")
    return refsim
