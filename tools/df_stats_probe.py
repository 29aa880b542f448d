#!/usr/bin/env python
"""Tuning aid for the dataflow SOR kernel (library built with -DDF_STATS, tools/build_variant.py): slack histogram of need[], how often
the wait's slow path runs, how often it blocks and for how long."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UFM_SOR_TRACE"] = "1"
from ufemism_b200 import mesh as M, scenarios as S
from ufemism_b200.capi import IceModelGPU
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
c = S.CONFIG3
m = M.square_mesh_with_nv(c["half_width"], nv)
st = S.state_ssa_icestream(m, Hb=c["Hb"], H_shelf=c["H_shelf"])
g = IceModelGPU(m, benchmark=st["benchmark"], use_analytical_GL_flux=1)
for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
    g.upload(k, st[k])
g.update_general_ice_model_data(0.0); g.ssa_prepare(); g.ssa_viscosity(); g.ssa_sliding_and_setup()
g.ssa_sor(max_inner=3, force_iters=True)
t0 = g.sor_trace(raw=True).astype(np.int64).ravel().copy()
g.reset_counters(); g.ssa_sor(max_inner=20, force_iters=True)
cn = g.counters()
t = g.sor_trace(raw=True).astype(np.int64).ravel() - t0
per = t[:148 * 8].reshape(148, 8)
hist = g.sor_trace(raw=True).astype(np.int64).ravel()[4096 * 12: 4096 * 12 + 64]
out = {"us_per_iteration": cn.sor_ms * 1e3 / cn.sor_iterations, "slack_hist": {int(i): int(v) for i, v in enumerate(hist) if v},
       "slow_path_per_warp_per_iteration": float(per[:, 0].sum() / (148 * 32 * 20)), "blocked_fraction_of_slow": float(per[:, 1].sum() / max(per[:, 0].sum(), 1)),
       "mean_us_in_slow_path": float(per[:, 2].sum() / max(per[:, 0].sum(), 1) / 1e3), "mean_lookahead_run": float(per[:, 3].sum() / max(per[:, 0].sum(), 1)),
       "us_in_slow_path_per_warp_per_iteration": float(per[:, 2].sum() / (148 * 32 * 20) / 1e3)}
print(json.dumps(out))
