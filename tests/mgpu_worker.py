"""Worker for the multi-GPU parity test: launched with torchrun, one rank per GPU.
  mgpu_worker.py N default      the shipping path: multigrid-preconditioned viscosity solve and pressure solve cut into
                                k-slabs, exchanges by peer-memory stores (xch.h)
  mgpu_worker.py N diagonal     the same with the diagonal preconditioner (every iteration is stencil + exchanges)
  FLIP_P2P=0                    no peer mapping: plain replicas
Rank 0 also steps a single-GPU simulation and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from __graft_entry__ import _analytic_scene  # noqa: E402
from flipviscosity3d_b200 import FlipSim  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mode = sys.argv[2] if len(sys.argv) > 2 else "default"
    phi, p = _analytic_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(2.0)
    if mode == "diagonal":
        sim.set_param("viscosity_precond", 0)
    box = [sim.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.dist_init(rank, world, box[0])
    if os.environ.get("FLIP_P2P", "1") == "1":
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.dist_p2p_export())
        sim.dist_p2p_import(blobs)
    for _ in range(3):
        sim.advance(0.01)
    out = sim.get_particles()
    st = sim.stats()
    # replicas must agree bit for bit
    t = torch.from_numpy(out.copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool((t == ref).all().item())
    ok = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        single = FlipSim(n, n, n, 1.0 / n)
        single.set_solid_sdf(phi); single.set_particles(p); single.set_viscosity(2.0)
        if mode == "diagonal":
            single.set_param("viscosity_precond", 0)
        for _ in range(3):
            single.advance(0.01)
        b = single.get_particles()
        st1 = single.stats()
        err = float(np.abs(out[:, :3] - b[:, :3]).max())
        print("MGPU_RESULT mode=" + mode + " world=%d replicas_identical=%d max_pos_diff_vs_single=%.3e visc_it=%d/%d pres_it=%d/%d unknowns=%d/%d"
              % (world, int(ok.item()), err, st["viscosity_iterations"], st1["viscosity_iterations"], st["pressure_iterations"],
                 st1["pressure_iterations"], st["viscosity_unknowns"], st1["viscosity_unknowns"]), flush=True)
        assert ok.item() == 1
        assert err < 1e-5
        assert st["viscosity_unknowns"] == st1["viscosity_unknowns"]
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
