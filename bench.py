#!/usr/bin/env python
"""bench.py — substeps/s of the FLIP substep hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scene bunny|sheet|cube|rod]
                  [--size 256] [--viscosity 5]

Workload (config.workload): BASELINE.json configs[3], the scene the metric is quoted on — the Stanford bunny dropped
inside the inverted sphere on a 256^3 grid, 8 seeded candidates per cell (4.7 M particles), viscosity 5 (the
reference's main.cpp value), frame dt 0.01.  A "step" is one substep of FluidSimulation::advance()
(/root/reference/src/fluidsimulation.cpp:138-167), with the substep size the loop itself would take: 5 dx / max|u|
clamped to the rest of the frame (_cfl, :241-269, :139-142).  Both arms start from the same scene at rest and take the
substeps in order: warm-up = substeps 1..W, timed = substeps W+1..W+K.

  value     substeps/s with all state resident in HBM: device time (CUDA events on the library's stream around the K
            substeps, CFL reductions included), max over ranks
  e2e       the SAME K substeps, replayed from the saved post-warm-up particles, through the C ABI with HOST particle
            buffers: every step uploads the particles (flip_set_particles), runs the substep and reads them back
            (flip_get_particles); the final particles must equal the resident run's bit for bit
  roofline  dominant kernel = the Jacobi sweep on the first explicit level of the Galerkin multigrid that preconditions
            the viscosity CG (k_gmg_sweep_tma<1>): algorithmic bytes = rows x (stored coefficients + row index + weight + b
            + x in + x out) x 4 B, divided by its average launch duration measured with CUDA events on the library's
            stream (flip_time_kernel) right after the timed region, on the hierarchy of the last timed substep;
            `stages` holds the same arithmetic for every stage in SURVEY.md 8(d) units
  cpu_baseline  the reference's own C++ path (oracle/_ref, 1 thread: it has no threading): the number the
            `--impl reference` run left on this box if there is one (full workload), else a bounded 128^3 sample
  --impl reference   the reference's CPU implementation on the same workload, scene built by the reference's own
            addBoundary/addLiquid, same substep rule; honours --steps/--warmup inside a wall-clock budget
            (REF_BUDGET_S, stated in the line): one 256^3 substep is ~80-150 s on one core
"""
import argparse
import hashlib
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
REF_BUDGET_S = 240.0
# algorithmic bytes per unit of work (fp64 CG vectors like the reference)
ALG_BYTES_VISC = 13 * 8 + 16      # per unknown per iteration
ALG_BYTES_PRES = 13 * 8 + 20      # per unknown per iteration
ALG_BYTES_P2G = 26                # per particle
ALG_BYTES_G2P = 52                # per particle

# BASELINE.json configs: name -> (liquid mesh, boundary mesh or None, inverted, configs[] index)
SCENES = {
    "bunny": ("stanford_bunny", "sphere_large", True, 3),
    "cube": ("cube", None, False, 1),
    "rod": ("rod", None, False, 2),
    "sheet": ("sheet", None, False, 4),
}


def read_ply(path):
    """binary little-endian PLY (positions + triangles), whole-file parse (also the < 2048-byte sample meshes the
    reference's own loader rejects)"""
    data = open(path, "rb").read()
    hend = data.index(b"end_header\n") + len(b"end_header\n")
    hdr = data[:hend].decode().split("\n")
    nv = int([l for l in hdr if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in hdr if l.startswith("element face")][0].split()[-1])
    v = np.frombuffer(data, np.float32, nv * 3, hend).reshape(nv, 3).copy()
    rec = np.dtype([("n", "u1"), ("i", "<i4", 3)])
    f = np.frombuffer(data, rec, nf, hend + nv * 12)
    return v, f["i"].astype(np.int32).copy()


def _scene_key(scene, n):
    h = hashlib.sha256()
    liquid, boundary, inverted, _ = SCENES[scene]
    for name in (liquid, boundary):
        if name:
            h.update(open(os.path.join(MESHES, name + ".ply"), "rb").read())
    host = os.path.join(ROOT, "flipviscosity3d_b200", "lib", "libflip_host.so")
    if os.path.exists(host):
        h.update(open(host, "rb").read())
    h.update(("%s %d %d" % (scene, n, int(inverted))).encode())
    return h.hexdigest()[:16]


def build_scene(scene, n, cache=True):
    """Scene through the host C++ layer (bit-identical to the reference's init, tests/test_host_scene.py); cached under
    /tmp, keyed by the mesh files and the host library, because the mesh SDF is single-threaded host work."""
    path = "/tmp/flip_scene_%s_%d_%s.npz" % (scene, n, _scene_key(scene, n))
    if cache and os.path.exists(path):
        d = np.load(path)
        return d["phi"], d["p"]
    from flipviscosity3d_b200 import scene as hs
    liquid, boundary, inverted, _ = SCENES[scene]
    sc = hs.Scene(n, n, n, 1.0 / n)
    if boundary:
        sc.add_boundary(*read_ply(os.path.join(MESHES, boundary + ".ply")), inverted=inverted)
    sc.add_liquid(*read_ply(os.path.join(MESHES, liquid + ".ply")))
    phi, p = sc.solid_sdf(), sc.particles()
    sc.close()
    if cache:
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, phi=phi, p=p)
        os.replace(tmp, path)
    return phi, p


def load_scene_device(sim, scene):
    """Scene built on the device through the C ABI (csrc/scene.cu: flip_reset_boundary / flip_add_boundary_mesh /
    flip_add_liquid_mesh): particles bit-identical to the reference's addLiquid (same seeding order, same rand() sequence),
    solid SDF identical within 3 cells of the surface and exact beyond (tests/test_gpu_scene.py).  Returns the seconds taken."""
    t0 = time.perf_counter()
    liquid, boundary, inverted, _ = SCENES[scene]
    sim.srand(1)
    sim.reset_boundary()
    if boundary:
        sim.add_boundary(*read_ply(os.path.join(MESHES, boundary + ".ply")), inverted=inverted)
    sim.add_liquid(*read_ply(os.path.join(MESHES, liquid + ".ply")))
    sim.synchronize()
    return time.perf_counter() - t0


class FrameStepper:
    """The substep loop of FluidSimulation::advance (src/fluidsimulation.cpp:135-168), one substep per call: float
    arithmetic like the reference, substep = cfl() clamped to the rest of the frame.  Works on FlipSim and on the oracle's
    RefSim (same cfl()/substep() surface)."""

    def __init__(self, sim, frame_dt=FRAME_DT):
        self.sim, self.dt, self.t = sim, np.float32(frame_dt), np.float32(0.0)

    def next_dt(self):
        with np.errstate(over="ignore", invalid="ignore"):
            sub = np.float32(self.sim.cfl())
            if not (self.t + sub <= self.dt):       # also catches +inf (field at rest) and NaN
                sub = self.dt - self.t
        return sub

    def advance_clock(self, sub):
        self.t = np.float32(self.t + sub)
        if not (self.t < self.dt):
            self.t = np.float32(0.0)

    def step(self):
        sub = self.next_dt()
        self.sim.substep(float(sub))
        self.advance_clock(sub)
        return float(sub)


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


def workload_name(args):
    return "%s_%d^3_8ppc_viscosity%g (BASELINE.json configs[%d])" % (
        {"bunny": "bunny_in_sphere"}.get(args.scene, args.scene), args.size, args.viscosity, SCENES[args.scene][3])


def workload_config(args, particles, parallelism=None):
    cfg = {"workload": workload_name(args), "grid": [args.size] * 3, "particles": int(particles),
           "viscosity": args.viscosity, "frame_dt": FRAME_DT,
           "step": "one substep of FluidSimulation::advance; substep size = the loop's own rule (5 dx / max|u|, clamped "
                   "to the frame), substeps taken in order from the scene at rest",
           "viscosity_solver": "PCG, Galerkin multigrid V-cycle preconditioner (viscosity_precond=%d)" % getattr(args, "precond", 2),
           "l2": "working set (fields, CG vectors, ~280 MB of level-1 multigrid coefficients) exceeds the 126 MB L2; no "
                 "flush between steps"}
    if parallelism:
        cfg["parallelism"] = parallelism
    return cfg


def ref_cache_path(args):
    return "/tmp/flip_ref_arm_%s_%d_%g.json" % (args.scene, args.size, args.viscosity)


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own C++ path on the host (oracle/_ref = unmodified reference sources)
# ---------------------------------------------------------------------------------------------------------------
def reference_scene(refsim, args):
    """scene built by the reference's OWN initialize/addBoundary/addLiquid (src/main.cpp:47-75)"""
    n = args.size
    liquid, boundary, inverted, _ = SCENES[args.scene]
    refsim.srand(1)
    ref = refsim.RefSim(n, n, n, 1.0 / n)
    if boundary:
        ref.add_boundary(*read_ply(os.path.join(MESHES, boundary + ".ply")), inverted)
    ref.add_liquid(*read_ply(os.path.join(MESHES, liquid + ".ply")))
    ref.set_viscosity(args.viscosity)
    ref.set_gravity(0.0, -9.81, 0.0)
    return ref


def time_reference(args, steps, warmup, budget_s):
    """Up to `warmup` untimed + `steps` timed substeps of the reference, in order from the scene at rest, inside a
    wall-clock budget: when the first substep shows that warm-up + steps do not fit, the warm-up is dropped (the first
    substep is then the first timed one) and as many timed substeps as fit are taken (at least one)."""
    from oracle import refsim
    if not refsim.build():
        return None
    ref = reference_scene(refsim, args)
    stepper = FrameStepper(ref)
    done, times, dts, stage = 0, [], [], np.zeros(8)
    t_begin = time.perf_counter()
    first = None
    timed_from = warmup
    while True:
        elapsed = time.perf_counter() - t_begin
        if done >= timed_from + steps:
            break
        if first is not None:
            per = elapsed / done
            if done < timed_from and elapsed + per * (timed_from - done + 1) > budget_s:
                timed_from = 0            # no room for the warm-up: everything taken so far counts as timed
            if done >= max(timed_from, 0) + 1 and elapsed + per > budget_s:
                break
        sub = stepper.next_dt()
        t0 = time.perf_counter()
        st = ref.substep(float(sub))
        dt = time.perf_counter() - t0
        stepper.advance_clock(sub)
        times.append(dt); dts.append(float(sub))
        if st is not None:
            stage += np.array(st)
        done += 1
        first = dt
    timed = times[timed_from:]
    return {"seconds": float(sum(timed)), "steps": len(timed), "warmup": timed_from, "particles": ref.num_particles(),
            "substep_dts": dts, "stage_seconds": (stage / max(1, done)).tolist(), "per_step_seconds": times}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is single-process; other ranks exit 0 without work
    r = time_reference(args, args.steps, args.warmup, REF_BUDGET_S)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libflipref.so missing and /root/reference absent"}))
        return
    value = r["steps"] / r["seconds"]
    sample = ("substeps %d..%d of the full workload in order from the scene at rest (asked: %d warm-up + %d timed; wall-clock "
              "budget %.0f s; one core, the reference has no threading)" % (r["warmup"] + 1, r["warmup"] + r["steps"],
                                                                            args.warmup, args.steps, REF_BUDGET_S))
    line = {
        "impl": "reference", "metric": "substeps_per_second", "value": value, "unit": "substeps/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 fields / f64 CG vectors", "data": "synthetic",
        "config": workload_config(args, r["particles"]),
        "cpu_baseline": {"value": value, "unit": "substeps/s", "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_seconds": r["stage_seconds"], "substep_dts": r["substep_dts"], "host_cores_available": os.cpu_count(),
    }
    try:   # left for the B200 arm's cpu_baseline on the same box
        json.dump({"value": value, "sample": sample, "steps": r["steps"]}, open(ref_cache_path(args), "w"))
    except Exception:
        pass
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def dist_setup(sim, rank, world, use_p2p=True):
    """NCCL communicator of the library (rank 0 makes the unique id, torch.distributed ships it) and the peer-memory
    mapping of the other ranks' symmetric heaps (CUDA IPC over NVLink).  Returns True when peer memory is active."""
    import torch
    import torch.distributed as dist
    box = [sim.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.dist_init(rank, world, box[0])
    if not use_p2p:
        return False
    ok = 1
    try:
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.dist_p2p_export())
        sim.dist_p2p_import(blobs)
    except Exception as e:
        ok = 0
        print("bench.py: rank %d: peer memory unavailable (%s)" % (rank, e), file=sys.stderr)
    flag = torch.tensor([ok], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        raise SystemExit("bench.py: peer-memory mapping failed on some rank; the sharded substep needs it (run with --gpus 1)")
    return True


def multi_gpu_self_check(rank, world, use_p2p):
    """Evidence for the sharded path where the driver can see it (its pytest lease has one GPU): a small scene stepped by
    all ranks together and, on rank 0, by a single-GPU handle; the replicas must be bit-identical to each other and
    agree with the single-GPU run to 1e-5."""
    import torch
    import torch.distributed as dist
    from __graft_entry__ import _analytic_scene
    from flipviscosity3d_b200 import FlipSim
    n = 64
    phi, p = _analytic_scene(n)
    sim = FlipSim(n, n, n, 1.0 / n)
    sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(2.0)
    dist_setup(sim, rank, world, use_p2p)
    for _ in range(3):
        sim.advance(FRAME_DT)
    out = sim.get_particles()
    st = sim.stats()
    t = torch.from_numpy(out.copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1 if bool((t == ref).all().item()) else 0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res = {"scene": "analytic block %d^3, viscosity 2, 3 frames" % n, "replicas_identical": int(same.item())}
    if rank == 0:
        single = FlipSim(n, n, n, 1.0 / n)
        single.set_solid_sdf(phi); single.set_particles(p); single.set_viscosity(2.0)
        for _ in range(3):
            single.advance(FRAME_DT)
        b = single.get_particles()
        res["max_pos_diff_vs_single_gpu"] = float(np.abs(out[:, :3] - b[:, :3]).max())
        res["viscosity_unknowns"] = [int(st["viscosity_unknowns"]), int(single.stats()["viscosity_unknowns"])]
        single.close()
    sim.close()
    dist.barrier()
    if rank == 0:
        assert res["replicas_identical"] == 1, res
        assert res["max_pos_diff_vs_single_gpu"] < 1e-5, res
    return res


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
    check = multi_gpu_self_check(rank, world, not args.no_p2p) if world > 1 and not args.no_self_check else None
    sim = FlipSim(n, n, n, 1.0 / n)
    if args.host_scene:
        if rank == 0:
            phi, p = build_scene(args.scene, n)
        if world > 1:
            dist.barrier()
            if rank != 0:
                phi, p = build_scene(args.scene, n)
        sim.set_solid_sdf(phi)
        sim.set_particles(p)
        scene_s = None
    else:
        scene_s = load_scene_device(sim, args.scene)      # every rank builds the (deterministic) scene on its own GPU
    p = np.empty((sim.num_particles(), 6), np.float32)    # only its length / shape is used below
    sim.set_viscosity(visc)
    sim.set_param("viscosity_precond", args.precond)
    for kv in args.param or []:
        k, v = kv.split("=")
        sim.set_param(k, float(v))
    if world > 1:
        dist_setup(sim, rank, world, not args.no_p2p)

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sim.synchronize()

    stepper = FrameStepper(sim)
    for _ in range(args.warmup):
        stepper.step()
    # state after the warm-up: the particles (everything else is rebuilt from them every substep) and the frame clock
    start = torch.empty((len(p), 6), dtype=torch.float32).pin_memory().numpy()
    sim.get_particles(out=start)
    st0 = sim.stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier_sync()
    t0 = time.perf_counter()
    stage = np.zeros(8)
    iters = {"p": 0, "v": 0, "p_ms": 0.0, "v_ms": 0.0, "p_unk": 0, "v_unk": 0, "v_setup_ms": 0.0}
    dts, flags = [], []
    sim.event_record(0)
    for _ in range(args.steps):
        dts.append(stepper.step())
        st = sim.stats()
        stage += np.array(st["stage_ms"])
        iters["p"] += st["pressure_iterations"]; iters["v"] += st["viscosity_iterations"]
        iters["p_ms"] += st["pressure_solve_ms"]; iters["v_ms"] += st["viscosity_solve_ms"]
        iters["v_setup_ms"] += st["viscosity_setup_ms"]
        iters["p_unk"] += st["pressure_unknowns"] * st["pressure_iterations"]
        iters["v_unk"] += st["viscosity_unknowns"] * st["viscosity_iterations"]
        flags.append((st["pressure_converged"], st["viscosity_converged"] if visc > 0 else 1,
                      st["viscosity_applied"] if visc > 0 else 1, st["pressure_residual"], st["viscosity_residual"]))
    sim.event_record(1)
    dev_ms = sim.event_elapsed_ms(0, 1)
    barrier_sync()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    st1 = sim.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    bad = [i for i, f in enumerate(flags) if not (f[0] and f[1] and f[2])]
    if bad:
        raise SystemExit("bench.py: a timed solve did not converge (substeps %s of the timed region: "
                         "(pressure_converged, viscosity_converged, viscosity_applied, residuals) = %s)" % (bad, [flags[i] for i in bad]))
    resident_final = sim.get_particles().copy()
    t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_ms = t.tolist()
    ms_per_step = dev_ms / args.steps          # CUDA events on the library's stream, max over ranks
    value = 1e3 / ms_per_step

    # end to end: the same K substeps replayed from the saved state, particles crossing the C ABI as HOST buffers
    host = torch.empty((len(p), 6), dtype=torch.float32).pin_memory().numpy()
    host[:] = start
    barrier_sync()
    t0 = time.perf_counter()
    for sub in dts:
        sim.set_particles(host)
        sim.cfl()                              # the loop's CFL reduce (its value is the recorded one: same window)
        sim.substep(sub)
        sim.get_particles(out=host)
    barrier_sync()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item() / args.steps
    same_window = bool(np.array_equal(host, resident_final))

    # strict-parity mode for context: the same first substeps of the window with the reference's fp32-rounded viscosity
    # rows (viscosity_operator = 1): more iterations for bit-level agreement with the reference's system
    strict = None
    if visc > 0 and args.precond == 2 and not args.no_strict:
        ks = min(3, len(dts))
        sim.set_particles(start)
        sim.set_param("viscosity_operator", 1)
        barrier_sync()
        sim.event_record(2)
        its = 0
        for sub in dts[:ks]:
            sim.cfl()
            sim.substep(sub)
            its += sim.stats()["viscosity_iterations"]
        sim.event_record(3)
        ms = sim.event_elapsed_ms(2, 3) / ks
        sim.set_param("viscosity_operator", 0)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strict = {"substeps_per_s": 1e3 / t.item(), "ms_per_step": t.item(), "steps": ks, "viscosity_iterations_per_step": its / ks,
                  "note": "viscosity_operator=1: rows with the reference's fp32-rounded diagonal (tests/test_parity_regime.py)"}

    peak, peak_kind = measured_peak()
    # dominant kernel, timed live with CUDA events on the library's stream (hierarchy of the last substep)
    roof = None
    try:
        k_ms, k_bytes = sim.time_kernel("gmg_sweep_l1", 40)
        a_ms, a_bytes = sim.time_kernel("visc_apply", 40)
        roof = {"bound": "hbm", "kernel": "k_gmg_sweep_tma<1> on multigrid level 1 (Jacobi sweep over explicit Galerkin rows, one warp per row, "
                          "coefficient rows staged by cp.async.bulk)",
                "achieved": k_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_kind": peak_kind,
                "ms_per_launch": k_ms, "algorithmic_bytes_per_launch": k_bytes, "traffic": ncu_traffic("gmg_sweep_l1"),
                "other_kernels": {"k_visc_apply": {"ms_per_launch": a_ms, "algorithmic_bytes_per_launch": a_bytes,
                                                    "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "traffic": ncu_traffic("visc_apply")}}}
        roof["frac"] = roof["achieved"] / peak
    except Exception as e:   # diagonal preconditioner selected / no viscosity: fall back to the CG-iteration figure
        vis_dom = iters["v_ms"] >= iters["p_ms"]
        alg = (ALG_BYTES_VISC * iters["v_unk"]) if vis_dom else (ALG_BYTES_PRES * iters["p_unk"])
        solve_ms = iters["v_ms"] if vis_dom else iters["p_ms"]
        achieved = alg / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "CG iteration (stencil apply + update + direction)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind, "traffic": None, "note": str(e)}

    def gbs(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None

    K = args.steps
    stages = {   # SURVEY.md 8(d) units, whole job (all ranks), stage device time of rank 0
        "p2g": {"alg_bytes_per_step": ALG_BYTES_P2G * len(p), "ms": stage[1] / K, "note": "P2G + masks + extrapolation + save"},
        "g2p_advect": {"alg_bytes_per_step": ALG_BYTES_G2P * len(p), "ms": stage[6] / K},
        "pressure_pcg": {"alg_bytes_per_step": ALG_BYTES_PRES * iters["p_unk"] / K, "ms": iters["p_ms"] / K},
        "viscosity_pcg": {"alg_bytes_per_step": ALG_BYTES_VISC * iters["v_unk"] / K, "ms": iters["v_ms"] / K},
    }
    for v in stages.values():
        v["achieved_gbs"] = gbs(v["alg_bytes_per_step"], v["ms"])
        v["frac"] = v["achieved_gbs"] / peak if v["achieved_gbs"] else None

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                c = json.load(open(ref_cache_path(args)))
                cpu = {"value": c["value"], "unit": "substeps/s", "cores": 1, "kind": "reference",
                       "sample": "from `bench.py --impl reference` on this box: " + c["sample"], "host_cores_available": os.cpu_count()}
            except Exception:
                small = argparse.Namespace(**vars(args))
                small.size = max(32, args.size // 2)
                r = time_reference(small, 1, 0, 60.0)
                if r is not None:
                    cpu = {"value": r["steps"] / r["seconds"], "unit": "substeps/s", "cores": 1, "kind": "reference",
                           "sample": "first substep of the same scene at %d^3 (%d particles, 1/8 of the cells: the %d^3 substep "
                                     "takes 80-150 s on one core; `--impl reference` times the full workload)"
                                     % (small.size, r["particles"], args.size),
                           "host_cores_available": os.cpu_count()}
        if world == 1:
            par = "x1: one GPU runs the whole substep"
        else:
            par = ("x%d, strong scaling of ONE %d^3 substep: every rank holds the whole (replicated, bit-identical) state in its "
                   "180 GB; the work is cut into k-slabs balanced by liquid cells; slab results and ghost planes travel by "
                   "peer-memory stores over NVLink with flag hand-shakes inside the kernels (no NCCL on the substep path)"
                   % (world, args.size))
        line = {
            "metric": "substeps_per_second", "value": value, "unit": "substeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 fields / f64 CG vectors", "data": "synthetic",
            "config": workload_config(args, len(p), par),
            "e2e": {"value": 1.0 / e2e_s, "unit": "substeps/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(host.nbytes), "steps": args.steps, "same_window_as_value": True,
                    "final_particles_identical_to_resident_run": same_window},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "stage_ms": (stage / args.steps).tolist(),
            "stages": stages,
            "substep_dts": dts,
            "scene_build_seconds": scene_s,
            "strict_parity_mode": strict,
            "converged": {"pressure": all(f[0] for f in flags), "viscosity": all(f[1] for f in flags),
                          "viscosity_applied": all(f[2] for f in flags),
                          "max_pressure_residual": max(f[3] for f in flags), "max_viscosity_residual": max(f[4] for f in flags)},
            "pressure": {"iterations_per_step": iters["p"] / args.steps, "solve_ms_per_step": iters["p_ms"] / args.steps,
                         "unknowns": st1["pressure_unknowns"],
                         "cell_updates_per_s": iters["p_unk"] / (iters["p_ms"] * 1e-3) if iters["p_ms"] > 0 else None},
            "viscosity": {"iterations_per_step": iters["v"] / args.steps, "solve_ms_per_step": iters["v_ms"] / args.steps,
                          "setup_ms_per_step": iters["v_setup_ms"] / args.steps,
                          "unknowns": st1["viscosity_unknowns"],
                          "cell_updates_per_s": iters["v_unk"] / (iters["v_ms"] * 1e-3) if iters["v_ms"] > 0 else None},
            "p2g_particles_per_s": len(p) / (stage[1] / args.steps * 1e-3) if stage[1] > 0 else None,
        }
        if check is not None:
            line["multi_gpu_check"] = check
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="bunny", choices=sorted(SCENES))
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--viscosity", type=float, default=5.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict", action="store_true", help="skip the strict-parity-mode side measurement")
    ap.add_argument("--host-scene", action="store_true", help="build the scene with the host C++ layer instead of on the device")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: do not map peer memory (replicas only)")
    ap.add_argument("--no-self-check", action="store_true", help="multi-GPU: skip the start-up sharded-vs-single check")
    ap.add_argument("--precond", type=int, default=2, help="viscosity preconditioner: 2 Galerkin multigrid (default), 0 diagonal")
    ap.add_argument("--param", action="append", help="name=value passed to flip_set_param (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
