"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
B200 arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--size", "24", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "substeps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "0", "--size", "24")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
