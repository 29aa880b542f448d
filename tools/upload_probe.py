#!/usr/bin/env python
"""Mesh-update / restart data path at full size (rows N3, N4): time the device upload from primary mesh data
(ufm_mesh_upload_primary; UFM_UPLOAD_TIMING=1 prints the library's own phase breakdown on stderr), the upload from host-built
secondary data, and one restart / help_fields time frame written from the device.  Host wall clock: these are host-side paths."""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=1000000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import numpy as np

    from ufemism_b200 import mesh as M
    from ufemism_b200 import restart as R
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU

    c = S.CONFIG3
    t0 = time.time()
    m = M.square_mesh_with_nv(c["half_width"], a.nv, order="random")
    out = {"nV": m.nV, "nAc": m.nAc, "nVAaAc": m.nVAaAc, "host_cores": os.cpu_count(), "mesh_substrate_s": round(time.time() - t0, 2)}
    st = S.state_ssa_icestream(m, Hb=c["Hb"], H_shelf=c["H_shelf"])
    os.environ["UFM_UPLOAD_TIMING"] = "1"

    def timed(label, fn, reps=2):
        ts = []
        for k in range(reps):
            sys.stderr.write(f"--- {label} #{k + 1}\n"); sys.stderr.flush()
            t = time.time(); fn(); ts.append(round(time.time() - t, 3))
        out[label + "_s"] = ts

    g = IceModelGPU(m, benchmark=st["benchmark"], use_analytical_GL_flux=1, primary_only=True)   # first upload: allocates the arena
    timed("reupload_from_primary", lambda: g.upload_mesh(m))
    g.primary_only = False; g.derive_nf = True
    timed("reupload_secondary_from_host_nf_on_device", lambda: g.upload_mesh(m))
    g.derive_nf = False
    timed("reupload_everything_from_host", lambda: g.upload_mesh(m), reps=1)
    g.primary_only = True
    g.upload_mesh(m)
    os.environ.pop("UFM_UPLOAD_TIMING")
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    r = g.region(0.0)
    g.run_model(r, 1e12, max_steps=2)
    d = tempfile.mkdtemp()
    zeta = [g.P.zeta[k] for k in range(g.P.nZ)]
    fn, hf = os.path.join(d, "restart_ANT_00001.nc"), os.path.join(d, "help_fields_ANT_00001.nc")
    names = ["Hi", "Hb", "Hs", "SL", "U_SSA", "V_SSA", "U_SIA", "V_SIA", "mask", "dHs_dx", "dHs_dy", "D_SIA"]
    t = time.time(); R.create_restart(fn, m, zeta, {"TriC": m.TriC}); out["restart_create_s"] = round(time.time() - t, 3)
    t = time.time(); R.create_help_fields(hf, m, zeta, names, {"TriC": m.TriC}); out["help_fields_create_s"] = round(time.time() - t, 3)
    ts, th = [], []
    for k in range(3):
        t = time.time(); g.write_restart(fn, r.time + k); ts.append(round(time.time() - t, 4))
        t = time.time(); g.write_help_fields(hf, r.time + k, names); th.append(round(time.time() - t, 4))
    out["restart_frame_s"], out["help_fields_frame_s"] = ts, th
    out["restart_bytes"], out["help_fields_bytes"] = os.path.getsize(fn), os.path.getsize(hf)
    t = time.time(); prim = R.read_restart_mesh(fn); out["restart_read_mesh_s"] = round(time.time() - t, 3)
    g2 = None
    t = time.time(); ti = g.load_restart(fn, r.time + 2); out["restart_load_s"] = round(time.time() - t, 4); out["frame_loaded"] = ti
    assert np.array_equal(prim["C"], m.C) and np.array_equal(prim["V"], m.V)
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
