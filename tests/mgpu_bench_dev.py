"""Dev helper (torchrun): per-iteration cost of the slab-decomposed solves, graphs on/off."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flipviscosity3d_b200 import FlipSim
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if rank == 0: bench.build_scene(n)
dist.barrier()
phi, p = bench.build_scene(n)
for p2p, variant in ((1, 0), (1, 1)):
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(5.0)
    box = [sim.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.dist_init(rank, world, box[0])
    sim.set_param('cg_variant_viscosity', variant)
    if p2p:
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.dist_p2p_export())
        sim.dist_p2p_import(blobs)
    line = []
    for step in range(3):
        sim.substep(0.01)
        st = sim.stats()
        line.append('%d:%dit/%.0fms(%.0fus)p%d/%.1fms' % (step, st['viscosity_iterations'], st['viscosity_solve_ms'], 1e3 * st['viscosity_solve_ms'] / max(1, st['viscosity_iterations']), st['pressure_iterations'], st['pressure_solve_ms']))
    if rank == 0: print('world', world, 'p2p', p2p, 'visc_variant', variant, ' '.join(line), flush=True)
    sim.close()
    dist.barrier()
dist.destroy_process_group()
