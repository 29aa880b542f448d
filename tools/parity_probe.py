#!/usr/bin/env python
"""Step-by-step GPU vs CPU-restatement comparison on the bench workload: per step the iteration counts, dt and the relative L2 difference of
Hi / U_SSA / V_SSA.  python tools/parity_probe.py --nv 1000000 --steps 5"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--nv", type=int, default=1000000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--threads", type=int, default=os.cpu_count())
ap.add_argument("--max-outer", type=int, default=0)
a = ap.parse_args()
from oracle.oracle import Oracle
from ufemism_b200.capi import IceModelGPU
m, st = bench.build_workload(a.nv)
o = Oracle(m, benchmark=st["benchmark"], nthreads=a.threads, use_analytical_GL_flux=1)
g = IceModelGPU(m, benchmark=st["benchmark"], use_analytical_GL_flux=1)
if a.max_outer:
    o.cfg.SSA_max_outer_loops = a.max_outer; g.set_params(SSA_max_outer_loops=a.max_outer)
for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
    o[k][:] = st[k]; g.upload(k, st[k])
ro, rg = o.region(0.0), g.region(0.0)
rel = lambda x, y: float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))
for k in range(a.steps):
    ao, ag = (ro.n_outer_total, ro.n_sor_total), (rg.n_outer_total, rg.n_sor_total)
    t = time.time(); o.run_model(ro, 1e12, max_steps=1); tc = time.time() - t
    g.run_model(rg, 1e12, max_steps=1)
    row = {"step": k, "cpu_s": round(tc, 2), "cpu": [ro.n_outer_total - ao[0], ro.n_sor_total - ao[1], ro.dt], "gpu": [rg.n_outer_total - ag[0], rg.n_sor_total - ag[1], rg.dt],
           "dt_crit_cpu": list(ro.dt_crit_last), "dt_crit_gpu": list(rg.dt_crit_last)}
    for f in ("Hi", "U_SSA", "V_SSA", "U_SIA", "Hs"):
        x, y = g.download(f), o[f]
        row["rel_" + f] = rel(x, y); row["nbits_" + f] = int(np.sum(x != y))
    print(json.dumps(row), flush=True)
