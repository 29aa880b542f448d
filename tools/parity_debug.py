#!/usr/bin/env python
"""Where do GPU and CPU part ways?  Runs both in lockstep with run_model for --lock steps, then does the next step piece by piece on both
sides and prints the relative L2 difference of the main fields after every piece."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--nv", type=int, default=250000)
ap.add_argument("--lock", type=int, default=2)
ap.add_argument("--outer", type=int, default=50)
a = ap.parse_args()
from oracle.oracle import Oracle
from ufemism_b200.capi import IceModelGPU
m, st = bench.build_workload(a.nv)
o = Oracle(m, benchmark=st["benchmark"], nthreads=os.cpu_count(), use_analytical_GL_flux=1)
g = IceModelGPU(m, benchmark=st["benchmark"], use_analytical_GL_flux=1)
for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
    o[k][:] = st[k]; g.upload(k, st[k])
ro, rg = o.region(0.0), g.region(0.0)
o.run_model(ro, 1e12, max_steps=a.lock); g.run_model(rg, 1e12, max_steps=a.lock)
rel = lambda x, y: float(np.linalg.norm(np.asarray(x, float) - np.asarray(y, float)) / max(np.linalg.norm(np.asarray(y, float)), 1e-300))
def cmp(tag, fields):
    print(tag, {f: (rel(g.download(f), o[f]), int(np.sum(g.download(f) != o[f]))) for f in fields}, flush=True)
print("locked", ro.time, rg.time, ro.dt, rg.dt, list(ro.do_), list(rg.do_))
cmp("start", ["Hi", "U_SSA", "V_SSA", "Up_SSA_Ac", "Up_SIA_Ac"])
o.calculate_ice_thickness_change(ro.dt); g.calculate_ice_thickness_change(ro.dt)
cmp("thk", ["Hi", "dHi_dt"])
o.update_general_ice_model_data(ro.time); g.update_general_ice_model_data(ro.time)
cmp("geom", ["Hs", "dHs_dx", "dHs_dx_shelf", "mask", "mask_gl", "mask_Ac", "mask_gl_Ac", "mask_shelf_Ac", "Hi_Ac", "dHs_dx_shelf_Ac", "dHi_dx_Ac"])
o.solve_SIA(); g.solve_SIA()
cmp("sia", ["D_SIA_Ac", "Up_SIA_Ac"])
o.basal_yield_stress(); o.calculate_GL_flux(); o.SSA_gather_AaAc()
g.ssa_prepare()
cmp("prepare", ["tau_c_AaAc", "phi_fric_AaAc", "U_SSA_AaAc", "V_SSA_AaAc", "Qabs_GL_Ac", "Qp_GL_Ac", "Ux_SSA_Ac"])
for k in range(a.outer):
    o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_viscosity(); g.ssa_sliding_and_setup()
    if k < 3 or k % 10 == 0: cmp(f"outer {k} visc", ["eta_AaAc", "N_AaAc", "S_AaAc", "RHSx_AaAc", "eu_i_AaAc"])
    n, res, rs, warn = o.solve_SSA_linearised()
    s = g.ssa_sor()
    if k < 3 or k % 10 == 0 or n != s.n_inner_last:
        print("   sor", n, s.n_inner_last, res, s.last_max_residual)
        cmp(f"outer {k} sor", ["U_SSA_AaAc", "V_SSA_AaAc"])
