#!/usr/bin/env python
"""Where do the CTAs' different finish times inside a colour phase come from -- the SM a CTA sits on, or the rows it was given?

Three traced launches (UFM_SOR_TRACE: %globaltimer per CTA and colour phase of the fourth iteration, plus %smid); for every colour the
time from the phase start to the CTA's last warp.  If the pattern repeats from launch to launch (correlation by blockIdx ~ 1) a static
per-CTA share of the slices could level it; if it follows the SM rather than the block, or does not repeat, it cannot."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UFM_SOR_TRACE"] = "1"


def main():
    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU

    c = S.CONFIG3
    m = M.square_mesh_with_nv(c["half_width"], 1000000)
    st = S.state_ssa_icestream(m, Hb=c["Hb"], H_shelf=c["H_shelf"])
    g = IceModelGPU(m, benchmark=st["benchmark"], device=0, use_analytical_GL_flux=1, exact_xy=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    g.update_general_ice_model_data(0.0)
    g.ssa_prepare(); g.ssa_viscosity(); g.ssa_sliding_and_setup()
    g.ssa_sor(max_inner=5, force_iters=True)
    runs, smids = [], []
    for _ in range(3):
        g.ssa_sor(max_inner=8, force_iters=True)
        t = g.sor_trace().astype(np.int64)
        start = t[:, :5, 0].min(axis=0)
        runs.append((t[:, :5, 2] - start[None, :]) / 1e3)       # (cta, colour): us from the phase start to the CTA's last warp
        smids.append(t[:, 5, 0].copy())
    runs = np.array(runs)                                        # (launch, cta, colour)
    out = {"n_ctas": int(runs.shape[1]), "same_sm_every_launch": bool(all(np.array_equal(smids[0], s) for s in smids)),
           "last_warp_done_us_min_med_max": [[float(np.min(runs[0][:, k])), float(np.median(runs[0][:, k])), float(np.max(runs[0][:, k]))] for k in range(5)]}
    cc = lambda a, b: float(np.corrcoef(a, b)[0, 1])
    out["corr_between_launches_per_colour"] = [[cc(runs[0][:, k], runs[1][:, k]), cc(runs[1][:, k], runs[2][:, k])] for k in range(5)]
    out["corr_between_colours_same_launch"] = [cc(runs[0][:, 0], runs[0][:, k]) for k in range(1, 5)]
    # if every CTA's share were scaled by (median time / its time) of launch 0, what spread would launch 1 have had?
    pred = []
    for k in range(5):
        f = np.median(runs[0][:, k]) / runs[0][:, k]
        t1 = runs[1][:, k] * f
        pred.append({"spread_us_as_measured": float(runs[1][:, k].max() - np.median(runs[1][:, k])), "spread_us_with_shares_from_launch_0": float(t1.max() - np.median(t1))})
    out["what_static_shares_would_do"] = pred
    order = np.argsort(smids[0])
    out["time_by_sm_colour0_first_32"] = [[int(smids[0][i]), round(float(runs[0][i, 0]), 2)] for i in order[:32]]
    np.save("gpurun_out/sor_imbalance_runs.npy", runs); np.save("gpurun_out/sor_imbalance_smid.npy", np.array(smids))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
