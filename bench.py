#!/usr/bin/env python
"""bench.py — substeps/s of the FLIP substep hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 256] [--viscosity 5]

Workload (config.workload): BASELINE.json configs[3], the scene the metric is quoted on — the
Stanford bunny dropped inside the inverted sphere on a 256^3 grid, 8 seeded candidates per cell
(4.7 M particles), viscosity 5 (the reference's main.cpp value), frame dt 0.01.  A "step" is one
substep of FluidSimulation::advance() (/root/reference/src/fluidsimulation.cpp:138-167).

  value     substeps/s with all state resident in HBM (device time of K substeps)
  e2e       the same through the C ABI with HOST particle buffers: every step uploads the particles
            (flip_set_particles), runs the substep and reads them back (flip_get_particles)
  roofline  dominant kernel = the Jacobi sweep on the first explicit level of the Galerkin multigrid that
            preconditions the viscosity CG (k_gmg_sweep<1>, ~30 % of the substep): algorithmic bytes =
            rows x (235 stencil coefficients + row index + weight + b + x in + x out) x 4 B, divided by its
            average launch duration measured with CUDA events on the library's stream (flip_time_kernel)
            right after the timed region, on the hierarchy of the last timed substep
  cpu_baseline  the reference's own C++ path (oracle/_ref, 1 thread: it has no threading) on a
            bounded sample: one substep of the same scene at 128^3
  --impl reference   the reference's CPU implementation on the 256^3 workload itself, bounded to
            ONE timed substep without warm-up (153 s/substep on one core)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MESHES = os.path.join(ROOT, "tests", "data", "meshes")
FRAME_DT = 0.01
ALG_BYTES_VISC = 13 * 8 + 16      # SURVEY.md §8(d), fp64 vectors
ALG_BYTES_PRES = 13 * 8 + 20


def build_scene(n, cache=True):
    """bunny in inverted sphere via the host C++ layer (bit-identical to the reference's init,
    tests/test_host_scene.py); cached under /tmp because the mesh SDF is single-threaded host work."""
    path = "/tmp/flip_scene_bunny_%d.npz" % n
    if cache and os.path.exists(path):
        d = np.load(path)
        return d["phi"], d["p"]
    from flipviscosity3d_b200 import scene as hs
    sc = hs.Scene(n, n, n, 1.0 / n)
    sc.add_boundary(*hs.read_ply(os.path.join(MESHES, "sphere_large.ply")), inverted=True)
    sc.add_liquid(*hs.read_ply(os.path.join(MESHES, "stanford_bunny.ply")))
    phi, p = sc.solid_sdf(), sc.particles()
    sc.close()
    if cache:
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, phi=phi, p=p)
        os.replace(tmp, path)
    return phi, p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), if present."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_reference_substeps(n, viscosity, steps, warmup):
    """Time the reference's own substep (oracle/_ref: unmodified reference sources) on the host."""
    from oracle import refsim
    if not refsim.build():
        return None
    phi, p = build_scene(n)
    ref = refsim.RefSim(n, n, n, 1.0 / n)
    ref.set_solid_sdf(phi)
    ref.set_particles(p)
    ref.set_viscosity(viscosity)
    for _ in range(warmup):
        ref.substep(FRAME_DT)
    t0 = time.perf_counter()
    stage = np.zeros(8)
    for _ in range(steps):
        stage += ref.substep(FRAME_DT)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "steps": steps, "particles": len(p), "stage_seconds": (stage / steps).tolist()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is single-process; other ranks exit 0 without work
    n, visc = args.size, args.viscosity
    steps = 1  # bounded: one 256^3 substep is ~150 s on one core
    r = cpu_reference_substeps(n, visc, steps, 0)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libflipref.so missing and /root/reference absent"}))
        return
    value = r["steps"] / r["seconds"]
    sample = "%d substep(s) of the full workload, no warm-up (one core, the reference has no threading)" % steps
    line = {
        "impl": "reference", "metric": "substeps_per_second", "value": value, "unit": "substeps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 0, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 fields / f64 CG vectors", "data": "synthetic",
        "config": workload_config(args, r["particles"]),
        "cpu_baseline": {"value": value, "unit": "substeps/s", "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_seconds": r["stage_seconds"], "host_cores_available": os.cpu_count(),
    }
    print(json.dumps(line))


def workload_config(args, particles):
    return {"workload": "bunny_in_sphere_%d^3_8ppc_viscosity%g (BASELINE.json configs[3])" % (args.size, args.viscosity),
            "grid": [args.size] * 3, "particles": int(particles), "viscosity": args.viscosity, "frame_dt": FRAME_DT,
            "step": "one substep of FluidSimulation::advance", "parallelism": "x%d: every rank runs the whole substep (multigrid viscosity solve not decomposed; pressure CG k-slab decomposed over %s only above 1 M unknowns per rank: 0.62 M here)" % (args.gpus, "NCCL" if getattr(args, "no_p2p", False) else "peer memory"),
            "viscosity_solver": "PCG, Galerkin multigrid V-cycle preconditioner, 3/1/2 sweeps on level 0/1/deeper (viscosity_precond=%d)" % getattr(args, "precond", 2),
            "l2": "working set (fields, CG vectors, 280 MB of level-1 multigrid coefficients) exceeds the 126 MB L2; no flush between steps"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from flipviscosity3d_b200 import FlipSim

    n, visc = args.size, args.viscosity
    if rank == 0:
        phi, p = build_scene(n)
    if world > 1:
        dist.barrier()
        if rank != 0:
            phi, p = build_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi)
    sim.set_particles(p)
    sim.set_viscosity(visc)
    sim.set_param("viscosity_precond", args.precond)
    if world > 1:
        # NCCL communicator of the library: rank 0 makes the unique id, torch.distributed ships it
        box = [sim.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sim.dist_init(rank, world, box[0])
        if not args.no_p2p:
            # per-iteration reductions / halos through peer memory (CUDA IPC) instead of NCCL; if any rank cannot map
            # its peers (no P2P path between two devices) every rank stays on NCCL
            ok = 1
            try:
                blobs = [None] * world
                dist.all_gather_object(blobs, sim.dist_p2p_export())
                sim.dist_p2p_import(blobs)
            except Exception as e:
                ok = 0
                print("bench.py: rank %d: peer-memory exchange unavailable (%s)" % (rank, e), file=sys.stderr)
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if ok:
                    print("bench.py: rank %d: falling back to NCCL exchanges because another rank could not map its peers" % rank, file=sys.stderr)
                args.no_p2p = True
                sim.set_param("dist_p2p", 0)

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sim.substep(FRAME_DT)
    st0 = sim.stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier_sync()
    t0 = time.perf_counter()
    dev_ms = 0.0
    stage = np.zeros(8)
    iters = {"p": 0, "v": 0, "p_ms": 0.0, "v_ms": 0.0, "p_unk": 0, "v_unk": 0}
    for _ in range(args.steps):
        sim.substep(FRAME_DT)
        st = sim.stats()
        dev_ms += st["stage_ms"][7]
        stage += np.array(st["stage_ms"])
        iters["p"] += st["pressure_iterations"]; iters["v"] += st["viscosity_iterations"]
        iters["p_ms"] += st["pressure_solve_ms"]; iters["v_ms"] += st["viscosity_solve_ms"]
        iters["p_unk"] += st["pressure_unknowns"] * st["pressure_iterations"]
        iters["v_unk"] += st["viscosity_unknowns"] * st["viscosity_iterations"]
    barrier_sync()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    st1 = sim.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_ms = t.tolist()
    ms_per_step = dev_ms / args.steps          # CUDA events on the library's stream, max over ranks
    value = 1e3 / ms_per_step

    # end to end through the C ABI with host buffers (pinned): H2D particles + substep + D2H particles
    e2e_steps = max(1, min(args.steps, 5))
    host = torch.empty((len(p), 6), dtype=torch.float32).pin_memory().numpy()
    host[:] = sim.get_particles()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.set_particles(host)
        sim.substep(FRAME_DT)
        sim.get_particles(out=host)
    barrier_sync()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    peak, peak_kind = measured_peak()
    # dominant kernel, timed live with CUDA events on the library's stream (hierarchy of the last substep)
    roof = None
    try:
        k_ms, k_bytes = sim.time_kernel("gmg_sweep_l1", 40)
        a_ms, a_bytes = sim.time_kernel("visc_apply", 40)
        roof = {"bound": "hbm", "kernel": "k_gmg_sweep<1> on multigrid level 1 (Jacobi sweep over explicit 235-slot Galerkin rows, one warp per row)",
                "achieved": k_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_kind": peak_kind,
                "ms_per_launch": k_ms, "algorithmic_bytes_per_launch": k_bytes, "traffic": ncu_traffic("gmg_sweep_l1"),
                "other_kernels": {"k_visc_apply": {"ms_per_launch": a_ms, "algorithmic_bytes_per_launch": a_bytes,
                                                    "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "traffic": ncu_traffic("visc_apply")}}}
        roof["frac"] = roof["achieved"] / peak
    except Exception as e:   # diagonal preconditioner selected: fall back to the CG-iteration figure
        vis_dom = iters["v_ms"] >= iters["p_ms"]
        alg = (ALG_BYTES_VISC * iters["v_unk"]) if vis_dom else (ALG_BYTES_PRES * iters["p_unk"])
        solve_ms = iters["v_ms"] if vis_dom else iters["p_ms"]
        achieved = alg / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "CG iteration (stencil apply + update + direction)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind, "traffic": None, "note": str(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cn = max(32, args.size // 2)
            r = cpu_reference_substeps(cn, visc, 1, 0)
            if r is not None:
                cpu = {"value": r["steps"] / r["seconds"], "unit": "substeps/s", "cores": 1, "kind": "reference",
                       "sample": "1 substep of the same scene at %d^3 (%d particles, 1/8 of the cells; the %d^3 substep takes "
                                 "~150 s on one core, see --impl reference)" % (cn, r["particles"], args.size),
                       "host_cores_available": os.cpu_count()}
        line = {
            "metric": "substeps_per_second", "value": value, "unit": "substeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 fields / f64 CG vectors", "data": "synthetic",
            "config": workload_config(args, len(p)),
            "e2e": {"value": 1.0 / e2e_s, "unit": "substeps/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(host.nbytes), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "stage_ms": (stage / args.steps).tolist(),
            "pressure": {"iterations_per_step": iters["p"] / args.steps, "solve_ms_per_step": iters["p_ms"] / args.steps,
                         "cell_updates_per_s": iters["p_unk"] / (iters["p_ms"] * 1e-3) if iters["p_ms"] > 0 else None},
            "viscosity": {"iterations_per_step": iters["v"] / args.steps, "solve_ms_per_step": iters["v_ms"] / args.steps,
                          "cell_updates_per_s": iters["v_unk"] / (iters["v_ms"] * 1e-3) if iters["v_ms"] > 0 else None},
            "p2g_particles_per_s": len(p) / (stage[1] / args.steps * 1e-3) if stage[1] > 0 else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--viscosity", type=float, default=5.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: keep NCCL for the per-iteration exchanges")
    ap.add_argument("--precond", type=int, default=2, help="viscosity preconditioner: 2 Galerkin multigrid (default), 0 diagonal")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
