"""Multi-process paths.

CPU (gloo, world_size 2): the host-side plumbing bench.py uses — unique-id shipping with
broadcast_object_list and the replica contract (two processes stepping the same scene produce
bit-identical particles), on the CPU-emulation build.  The sharded protocol itself is exercised on the CPU by
tests/test_sharded_emu.py (ranks = threads of one process).
GPU (needs >= 2 devices): the sharded substep against a single-GPU run, via torchrun.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_GLOO_WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
import common
from __graft_entry__ import _analytic_scene
from flipviscosity3d_b200 import FlipSim
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = common.emu_library() if rank == 0 else None
dist.barrier()
lib = lib or common.emu_library()
sim = FlipSim(16, 16, 16, 1.0 / 16, lib=lib)
# unique-id plumbing (the emulator returns zeros; the point is the broadcast path)
box = [sim.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
assert len(box[0]) == 128
sim.dist_init(0, 1, box[0])           # single-process communicator on the emulator
phi, p = _analytic_scene(16)
sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(1.0)
for _ in range(2):
    sim.advance(0.01)
digest = hashlib.sha256(sim.get_particles().tobytes()).hexdigest()
got = [None] * world
dist.all_gather_object(got, digest)
assert len(set(got)) == 1, got
print("GLOO_OK", rank, digest[:12])
dist.destroy_process_group()
'''


def test_gloo_two_process_replicas(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("GLOO_OK") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["diagonal", "default"])
def test_two_gpu_solves_match_single_gpu(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29733" if mode == "default" else "29735",
                        os.path.join(ROOT, "tests", "mgpu_worker.py"), "64", mode],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_RESULT" in r.stdout
