"""Stage parity in the regime the benchmark is quoted on (dt*mu/dx^2 ~ 3300: 256^3 at viscosity 5), teacher forced against
the oracle's raised-cap solves ("truncated reference solves").  The 128^3 / viscosity-20 fixture has
the same stiffness as 256^3 / viscosity 5 and is what the regular -m gpu suite runs; the 256^3 fixture itself is checked
when present (tests/golden_big/ is generated on the CPU by tests/golden/make_parity_big.py and is git-ignored).

Tolerances: velocities are O(1) m/s (max |u| ~ 0.2-0.3 here).  Library (multigrid-PCG, tol 1e-8 relative residual) vs oracle
(MICCG(0), tol 1e-8, cap raised): 5e-6 absolute on the faces bordering fluid.  Pressure: 1e-6 relative to max|p|."""
import glob
import os

import pytest

import common

BIG = os.path.join(common.ROOT, "tests", "golden_big")


def _check(res):
    lib8 = res["library_tol1e-08"]
    assert lib8["converged"] == 1 and lib8["applied"] == 1 and lib8["same_unknown_set_as_oracle"]
    assert lib8["linf_vs_oracle_1e8"] <= 5e-6, res
    lib6 = res["library_tol1e-06"]
    assert lib6["converged"] == 1
    # at the shipping tolerance the library is at least as close to the converged answer as the reference's own 1e-6 solve
    assert lib6["linf_vs_oracle_1e8"] <= max(2 * res["oracle_1e6_vs_oracle_1e8"], 2e-5), res
    assert res["pressure"]["converged"] == 1 and res["pressure"]["linf_relative"] <= 1e-6, res


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["parity_128_mu20.npz", "parity_256_mu5.npz"])
def test_stage_parity_in_bench_regime(cuda_lib, name):
    path = os.path.join(BIG, name)
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
