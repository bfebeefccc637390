"""Stage parity in the regime the benchmark is quoted on (dt*mu/dx^2 ~ 3300: 256^3 at viscosity 5), teacher forced against
the oracle's raised-cap solves ("truncated reference solves").  The 128^3 / viscosity-20 fixture has
the same stiffness as 256^3 / viscosity 5 and is what the regular -m gpu suite runs; the 256^3 fixture itself is checked
when present.  Both are generated on the CPU by tests/golden/make_parity_big.py; the 128^3 fixture (5 MB) is committed
under tests/golden/, the 256^3 one (35 MB) lives in the git-ignored tests/golden_big/ and its result is kept in
profiles/r2_parity_256_mu5.json.

Tolerances (velocities: max |u| ~ 0.2 here), on the faces bordering fluid:
  * library with the reference's own rows (viscosity_operator = 1: the fp32-rounded diagonal of
    src/viscositysolver.cpp:429, bit for bit), multigrid-PCG at tol 1e-8, vs oracle (MICCG(0), tol 1e-8, cap raised):
    5e-6 absolute;
  * library with its default rows (exact face-volume term): 1e-4 absolute.  In this regime the six factors of a row add
    up to ~2e4, so the reference's fp32 diagonal carries +-1e-3 of rounding noise against a mass term <= 1; that noise
    alone moves the reference's SOLUTION by ~3e-5 (measured: library default vs library with the reference's rows).
    The default keeps the better-conditioned exact rows; the strict mode exists to prove the two coincide otherwise;
  * pressure: 1e-6 relative to max|p|."""
import glob
import os

import pytest

import common

BIG = os.path.join(common.ROOT, "tests", "golden_big")


def _check(res):
    strict = res["library_reference_operator_tol1e-08"]
    assert strict["converged"] == 1 and strict["applied"] == 1 and strict["same_unknown_set_as_oracle"]
    assert strict["linf_vs_oracle_1e8"] <= 5e-6, res
    for key in ("library_tol1e-08", "library_tol1e-06"):
        lib = res[key]
        assert lib["converged"] == 1 and lib["applied"] == 1 and lib["same_unknown_set_as_oracle"]
        assert lib["linf_vs_oracle_1e8"] <= 1e-4, res
    assert res["pressure"]["converged"] == 1 and res["pressure"]["linf_relative"] <= 1e-6, res


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["parity_128_mu20.npz", "parity_256_mu5.npz"])
def test_stage_parity_in_bench_regime(cuda_lib, name):
    path = os.path.join(common.ROOT, "tests", "golden", name)     # committed fixture
    if not os.path.exists(path):
        path = os.path.join(BIG, name)                           # generated locally (git-ignored)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated (tests/golden/make_parity_big.py)" % name)
    import parity_big
    res = parity_big.run(path, lib=cuda_lib)
    print(res)
    _check(res)


def test_parity_big_on_emulator_small(emu_lib, oracle, tmp_path, monkeypatch):
    """the same script end to end at 24^3 on the CPU-emulated kernels (fixture generated on the fly)"""
    import subprocess
    import sys
    import parity_big
    out = os.path.join(BIG, "parity_24_mu40.npz")
    if not os.path.exists(out):
        subprocess.check_call([sys.executable, os.path.join(common.ROOT, "tests", "golden", "make_parity_big.py"), "24", "40"])
    res = parity_big.run(out, lib=emu_lib)
    _check(res)
