"""The sharded multi-GPU substep, end to end, in the GPU-less container: R ranks = R host threads of one process on the
CPU-emulated kernels (tests/cpu_emu), each with its own handle; "peer memory" is the other handles' heaps (the emulator's
IPC handle carries the pointer), and the hand-shakes of xch.h run for real between the threads.  Checked: the ranks'
results are bit-identical to each other and agree with a single-rank run (reductions are summed slab by slab, so the CG
scalars differ in the last bits).  The same protocol on real GPUs: tests/mgpu_worker.py (2 GPUs) and bench.py --gpus N
start-up check."""
import threading

import numpy as np
import pytest

import common
from __graft_entry__ import _analytic_scene
from flipviscosity3d_b200 import FlipSim


def run_ranks(lib, nranks, n, setup, work, timeout_s=10.0):
    """setup(sim) loads the scene; work(sim) steps it and returns a result.  Returns the per-rank results."""
    sims = [None] * nranks
    blobs = [None] * nranks
    results = [None] * nranks
    errors = []
    bar = threading.Barrier(nranks)

    def rank_main(r):
        try:
            sim = FlipSim(n, n, n, 1.0 / n, lib=lib)
            sims[r] = sim
            setup(sim)
            sim.set_param("xch_timeout_s", timeout_s)
            sim.dist_init(r, nranks, sim.dist_unique_id())
            blobs[r] = sim.dist_p2p_export()
            bar.wait()
            sim.dist_p2p_import(blobs)
            bar.wait()
            results[r] = work(sim)
            bar.wait()
        except Exception as e:  # noqa: BLE001
            errors.append((r, repr(e)))
            try:
                bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for sim in sims:
        if sim is not None:
            sim.close()
    assert not errors, errors
    return results


def _scene_setup(n, viscosity):
    phi, p = _analytic_scene(n)

    def setup(sim):
        sim.set_solid_sdf(phi); sim.set_particles(p); sim.set_viscosity(viscosity)
    return setup


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_sharded_frames_match_single_rank(emu_lib, nranks):
    n = 16
    setup = _scene_setup(n, 2.0)

    def work(sim):
        if nranks == 3:
            sim.set_param("shard_min_unknowns", 0)  # 3 ranks: pressure solve sharded as well; 2 ranks: the default policy
        for _ in range(2):
            sim.advance(0.01)
        st = sim.stats()
        return sim.get_particles().copy(), st

    single = FlipSim(n, n, n, 1.0 / n, lib=emu_lib)
    setup(single)
    ref, st1 = work(single)
    single.close()
    res = run_ranks(emu_lib, nranks, n, setup, work)
    for out, st in res:
        assert np.array_equal(out, res[0][0])                       # replicas: bit-identical
        assert st["viscosity_converged"] == 1 and st["pressure_converged"] == 1
        assert st["viscosity_unknowns"] == st1["viscosity_unknowns"] and st["pressure_unknowns"] == st1["pressure_unknowns"]
    assert common.maxdiff(res[0][0][:, :3], ref[:, :3]) <= 1e-5
    assert common.maxdiff(res[0][0][:, 3:], ref[:, 3:]) <= 1e-3


@pytest.mark.parametrize("precond", [2, 0], ids=["multigrid", "diagonal"])
def test_sharded_viscosity_solve_tight(emu_lib, precond):
    """one viscosity + pressure stage at a tight tolerance: sharded and single-rank solutions agree to 1e-7"""
    n = 24
    setup = _scene_setup(n, 3.0)

    def work(sim):
        sim.set_param("viscosity_precond", precond)
        sim.set_param("viscosity_tol", 1e-10)
        sim.set_param("shard_min_unknowns", 0)      # cut the (small) pressure system into slabs too
        sim.update_liquid_sdf(); sim.advect_velocity_field(); sim.add_body_force(0.01)
        sim.apply_viscosity(0.01)
        a = [x.copy() for x in sim.get_mac()]
        st = sim.stats()
        sim.solve_pressure(0.01)
        from flipviscosity3d_b200 import fields as F
        return a, sim.get_field(F.F_PRESSURE).copy(), st, sim.stats()

    single = FlipSim(n, n, n, 1.0 / n, lib=emu_lib)
    setup(single)
    a1, p1, sv1, sp1 = work(single)
    single.close()
    res = run_ranks(emu_lib, 2, n, setup, work)
    for a, p, sv, sp in res:
        assert sv["viscosity_converged"] == 1 and sp["pressure_converged"] == 1
        for x, y in zip(a, res[0][0]):
            assert np.array_equal(x, y)
        assert np.array_equal(p, res[0][1])
    for x, y in zip(res[0][0], a1):
        assert common.maxdiff(x, y) <= 1e-7
    assert common.maxdiff(res[0][1], p1) <= 1e-6 * max(1.0, float(np.abs(p1).max()))


def test_missing_rank_times_out_cleanly(emu_lib):
    """a rank that never shows up: the others give up after the time-out, report an error and leave the field untouched"""
    n = 16
    setup = _scene_setup(n, 2.0)
    out = {}

    def work(sim):
        if sim is not None and out.get("skip") is sim:
            return None
        return None

    sims, blobs = [None, None], [None, None]
    bar = threading.Barrier(2)
    result = {}

    def rank_main(r):
        sim = FlipSim(n, n, n, 1.0 / n, lib=emu_lib)
        sims[r] = sim
        setup(sim)
        sim.set_param("xch_timeout_s", 0.5)
        sim.dist_init(r, 2, sim.dist_unique_id())
        blobs[r] = sim.dist_p2p_export()
        bar.wait()
        sim.dist_p2p_import(blobs)
        bar.wait()
        if r == 0:
            sim.update_liquid_sdf(); sim.advect_velocity_field(); sim.add_body_force(0.01)
            before = [x.copy() for x in sim.get_mac()]
            try:
                sim.apply_viscosity(0.01)
                result["error"] = None
            except Exception as e:  # noqa: BLE001
                result["error"] = str(e)
            result["untouched"] = all(np.array_equal(x, y) for x, y in zip(sim.get_mac(), before))
        bar.wait()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for sim in sims:
        sim.close()
    assert result["error"] and "timed out" in result["error"]
    assert result["untouched"]
