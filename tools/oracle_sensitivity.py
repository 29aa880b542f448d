#!/usr/bin/env python
"""How sensitive is the bench trajectory itself?  The CPU restatement is run twice on the same mesh: once as is, once with the initial
thickness of ONE vertex changed by one unit in the last place.  Per step: relative L2 difference of Hi / U_SSA / V_SSA between the two runs.
(CPU only.)  python tools/oracle_sensitivity.py --nv 250000 --steps 4"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle.oracle import Oracle

ap = argparse.ArgumentParser()
ap.add_argument("--nv", type=int, default=250000)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
m, st = bench.build_workload(a.nv)
runs = []
for pert in (False, True):
    o = Oracle(m, benchmark=st["benchmark"], nthreads=os.cpu_count(), use_analytical_GL_flux=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    if pert:
        i = int(np.argmin(np.abs(np.hypot(m.V[:, 0], m.V[:, 1]) - 900e3)))     # a grounded vertex near the grounding line
        o["Hi"][i] = np.nextafter(o["Hi"][i], np.inf)
    runs.append((o, o.region(0.0)))
rel = lambda x, y: float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))
out = []
for k in range(a.steps):
    cnt = []
    for o, r in runs:
        b = (r.n_outer_total, r.n_sor_total)
        o.run_model(r, 1e12, max_steps=1)
        cnt.append([int(r.n_outer_total - b[0]), int(r.n_sor_total - b[1])])
    row = {"step": k, "counts": cnt, **{"rel_" + f: rel(runs[1][0][f], runs[0][0][f]) for f in ("Hi", "U_SSA", "V_SSA")}}
    out.append(row); print(json.dumps(row), flush=True)
