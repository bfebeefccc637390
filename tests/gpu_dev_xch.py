"""Dev tool: cost of one synchronising kernel of the multi-GPU exchange (xch.h), measured on real GPUs.
  torchrun --nproc-per-node 2 tests/gpu_dev_xch.py [N=128]"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from __graft_entry__ import _analytic_scene  # noqa: E402
from flipviscosity3d_b200 import FlipSim  # noqa: E402
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
phi, p = _analytic_scene(n)
sim = FlipSim(n, n, n, 1.0 / n)
sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(2.0)
sim.update_liquid_sdf()
if world > 1:
    bench.dist_setup(sim, rank, world, True)
names = {0: "plain kernel (no hand-shake)", 1: "enter + leave", 2: "enter + leave(fence)", 3: "leave only", 4: "leave(fence) only",
         5: "push halo fp64 x3 (1 plane each way)", 6: "enter + coalesced remote stores + leave", 7: "enter + sector-strided remote stores + leave",
         8: "barrier kernel (1 CTA)", 9: "push halo fp32 x3"}
fn = sim.lib.flip_debug_xch_bench
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
for ctas in (1, 148, 296, 1184):
    for mode in (0, 1, 2, 3, 4, 6, 7):
        us = C.c_float()
        rc = fn(sim.h, mode, 200, ctas, C.byref(us))
        if world > 1:
            dist.barrier()
        if rank == 0:
            print("ctas %5d  %-48s %8.2f us/kernel (rc %d)" % (ctas, names[mode], us.value, rc), flush=True)
for mode in (5, 8, 9):
    us = C.c_float()
    rc = fn(sim.h, mode, 200, 0, C.byref(us))
    if world > 1:
        dist.barrier()
    if rank == 0:
        print("            %-48s %8.2f us/kernel (rc %d)" % (names[mode], us.value, rc), flush=True)
if world > 1:
    dist.destroy_process_group()
