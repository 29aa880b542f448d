"""Generate the committed golden vectors (run from the repo root: python -m tests.golden.make_golden).

The Fortran reference cannot be run in this image, so these are NOT reference outputs: they pin the CPU
restatement's own outputs on a 600-vertex mesh so that later edits to oracle/ or to the mesh substrate
cannot silently change what the GPU path is compared against."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def run_case(m):
    from oracle.oracle import Oracle
    from ufemism_b200 import scenarios as S

    out = {}
    st = S.state_ssa_icestream(m, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0)
    o = Oracle(m, benchmark=st["benchmark"], nthreads=1, use_analytical_GL_flux=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    r = o.region(0.0)
    o.run_model(r, 1e12, max_steps=3)
    for k in ("Hi", "Hs", "dHi_dt", "U_SSA", "V_SSA", "Up_SSA_Ac", "U_SIA", "D_SIA_Ac", "mask", "mask_Ac", "tau_c_AaAc", "eta_AaAc", "dHs_dx_shelf_Ac"):
        out["ssa_" + k] = o[k].copy()
    out["ssa_counts"] = np.array([r.n_steps, r.n_ssa, r.n_outer_total, r.n_sor_total], dtype=np.int64)
    out["ssa_time"] = np.array([r.time, r.dt])
    st = S.state_halfar(m)
    o = Oracle(m, benchmark="Halfar", nthreads=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    r = o.region(0.0)
    o.run_model(r, 2.0)
    out["halfar_Hi"] = o["Hi"].copy()
    out["halfar_Up_SIA_Ac"] = o["Up_SIA_Ac"].copy()
    out["halfar_steps"] = np.array([r.n_steps], dtype=np.int64)
    return out


def run_case_thermo(m):
    """update_ice_temperature on the thermo dome: EISMINT ice properties (no libm call except pow in U_3D) and the
    temperature-dependent properties with sliding (exp, pow); three steps each."""
    from oracle.oracle import Oracle
    from ufemism_b200 import scenarios as S

    out = {}
    for bm in ("EISMINT_1", "none"):
        st = S.state_thermo_dome(m, benchmark=bm)
        o = Oracle(m, benchmark=bm, nthreads=1)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            o[k][:] = st[k]
        o.update_general_ice_model_data(0.0)
        if bm == "none":
            o.solve_SSA()
        for _ in range(3):
            rc, nu = o.update_ice_temperature()
            assert rc == 0 and nu == 0
        for k in ("Ti", "W_3D", "U_3D", "frictional_heating"):
            out[f"{bm}_{k}"] = o[k].copy()
    return out


if __name__ == "__main__":
    from ufemism_b200 import mesh as M

    m = M.square_mesh_with_nv(750e3, 600, seed=11)
    m.save(os.path.join(HERE, "mesh_600.npz"))
    np.savez_compressed(os.path.join(HERE, "oracle_600.npz"), **run_case(m))
    np.savez_compressed(os.path.join(HERE, "oracle_thermo_600.npz"), **run_case_thermo(m))
    print("written", os.listdir(HERE))
