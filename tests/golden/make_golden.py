"""Generates tests/golden/stage_golden_n16.npz from the compiled reference (oracle/_ref).

Run where /root/reference exists:  python tests/golden/make_golden.py
The fixture holds, for one small scene (bunny in inverted sphere, 16^3, viscosity 5), the
reference's inputs and outputs of every stage of one substep (teacher forcing), so the parity
tests can run where only the repository travels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
from oracle import refsim  # noqa: E402

N, VISC, DT = 16, 5.0, 0.01


def main():
    assert refsim.build(), "needs the reference sources"
    ref = common.make_ref_scene(N, viscosity=VISC)
    g = {"n": N, "viscosity": VISC, "dt": DT}
    g["solid_sdf"] = ref.get_solid_sdf()
    p = ref.get_particles()
    rng = np.random.default_rng(0)
    p[:, 3:] = (rng.standard_normal((len(p), 3)) * 0.5).astype(np.float32)
    ref.set_particles(p)
    g["particles0"] = p
    ref.compute_weights()
    g["weight_u"], g["weight_v"], g["weight_w"] = ref.get_weights()
    ref.update_liquid_sdf()
    g["liquid_sdf"] = ref.get_liquid_sdf()
    ref.advect_velocity_field()
    g["p2g_u"], g["p2g_v"], g["p2g_w"] = ref.get_mac()
    g["p2g_valid_u"], g["p2g_valid_v"], g["p2g_valid_w"] = ref.get_valid()
    ref.add_body_force(DT)
    g["force_u"], g["force_v"], g["force_w"] = ref.get_mac()
    vols = ref.viscosity_volumes()
    for name, a in zip(["c", "u", "v", "w", "eu", "ev", "ew"], vols):
        g["vol_" + name] = a
    info = ref.apply_viscosity(DT, tol=1e-10, maxit=20000)
    assert info["wrote"] == 1
    g["visc_u"], g["visc_v"], g["visc_w"] = ref.get_mac()
    g["pressure"] = ref.solve_pressure(DT)
    ref.apply_pressure(DT, g["pressure"])
    g["proj_valid_u"], g["proj_valid_v"], g["proj_valid_w"] = ref.get_valid()
    ref.extrapolate()
    ref.constrain()
    g["final_u"], g["final_v"], g["final_w"] = ref.get_mac()
    g["saved_u"], g["saved_v"], g["saved_w"] = ref.get_saved_mac()
    g["cfl"] = np.float32(ref.cfl())
    ref.advect_particles(DT)
    g["particles1"] = ref.get_particles()
    out = os.path.join(HERE, "stage_golden_n16.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
