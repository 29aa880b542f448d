#!/usr/bin/env python
"""Isolated SOR-kernel probe on the bench workload: forced iteration counts, CUDA-event timing from the
library's own counters.  Used for kernel tuning and as the short command wrapped by ncu
(B200_PROFILING.md): a number printed under ncu is never a bench value."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=1000000)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--exact-xy", dest="exact_xy", type=int, default=1)
    ap.add_argument("--order", default="random")
    ap.add_argument("--others", action="store_true", help="also time the per-step kernels")
    ap.add_argument("--thermo", action="store_true", help="time update_ice_temperature (row N2) on the same mesh, EISMINT ice properties")
    ap.add_argument("--checksum", action="store_true", help="add a sha256 of the downloaded (U, V) after the timed iterations (row-order experiments must not change it)")
    ap.add_argument("--variants", default="", help="semicolon-separated env settings to compare in one process, e.g. "
                    "'UFM_SOR_CHUNK=1;UFM_SOR_CHUNK=1,UFM_SOR_BAR=1' (the library re-reads them on every SOR launch)")
    a = ap.parse_args()
    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU

    c = S.CONFIG3
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = M.square_mesh_with_nv(c["half_width"], a.nv, order=a.order)
    st = S.state_ssa_icestream(m, Hb=c["Hb"], H_shelf=c["H_shelf"])
    g = IceModelGPU(m, benchmark=st["benchmark"], device=local, rank=rank, nranks=world, use_analytical_GL_flux=1, exact_xy=a.exact_xy)
    if world > 1:
        g.connect(dist, device=torch.device("cuda", local))
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    g.update_general_ice_model_data(0.0)
    g.ssa_prepare(); g.ssa_viscosity(); g.ssa_sliding_and_setup()
    g.ssa_sor(max_inner=5, force_iters=True)  # warm-up
    res = []
    for _ in range(a.reps):
        g.reset_counters()
        g.ssa_sor(max_inner=a.iters, force_iters=True)
        cn = g.counters()
        res.append(cn.sor_ms * 1e3 / cn.sor_iterations)
    cn = g.counters()
    out = {"ranks": world, "nV": m.nV, "M": m.nVAaAc, "exact_xy": a.exact_xy, "iters": a.iters, "us_per_iteration": res,
           "algorithmic_GB": cn.sor_bytes_per_iteration / 1e9, "achieved_GBps_best": cn.sor_bytes_per_iteration / (min(res) * 1e-6) / 1e9}
    if a.checksum:
        import hashlib

        import numpy as np
        out["row_order"] = os.environ.get("UFM_ROW_ORDER", "default")
        out["uv_sha256"] = hashlib.sha256(np.ascontiguousarray(g.download("U_SSA_AaAc")).tobytes() + np.ascontiguousarray(g.download("V_SSA_AaAc")).tobytes()).hexdigest()
    if a.variants:
        import numpy as np
        base = None
        out["variants"] = {}
        for spec in ["UFM_SOR_CHUNK=0,UFM_SOR_FUSE_BC=0,UFM_SOR_BAR=0"] + a.variants.split(";"):
            for kv in ("UFM_SOR_CHUNK=0,UFM_SOR_FUSE_BC=0,UFM_SOR_BAR=0," + spec).split(","):
                k_, v_ = kv.split("=")
                os.environ[k_] = v_
            g.upload("U_SSA_AaAc", np.zeros(m.nVAaAc)); g.upload("V_SSA_AaAc", np.zeros(m.nVAaAc))
            g.ssa_sor(max_inner=20, force_iters=True)
            uv = g.download("U_SSA_AaAc")
            if base is None:
                base = uv
            ts = []
            for _ in range(a.reps):
                g.reset_counters()
                g.ssa_sor(max_inner=a.iters, force_iters=True)
                cn = g.counters()
                ts.append(round(cn.sor_ms * 1e3 / cn.sor_iterations, 2))
            out["variants"][spec] = {"us_per_iteration": ts, "same_bits_as_baseline_after_20": bool(np.array_equal(uv, base))}
    if os.environ.get("UFM_SOR_TRACE"):
        import numpy as np
        for kv in os.environ.get("UFM_SOR_TRACE_VARIANT", "UFM_SOR_CHUNK=0,UFM_SOR_FUSE_BC=0,UFM_SOR_BAR=0").split(","):
            k_, v_ = kv.split("=")
            os.environ[k_] = v_
        g.ssa_sor(max_inner=8, force_iters=True)
        t = g.sor_trace().astype(np.int64)[:, :5, :]          # (cta, colour, {start, first done, last done, left})
        t0 = t[:, :, 0].min(axis=0)                            # earliest phase start over the grid, per colour
        rel = t - t0[None, :, None]
        out["trace_us"] = {
            "phase_len": ((t[:, :, 3].max(axis=0) - t0) / 1e3).round(2).tolist(),
            "start_spread": ((t[:, :, 0].max(axis=0) - t0) / 1e3).round(2).tolist(),
            "first_warp_done_min_med_max": [[round(float(x) / 1e3, 2) for x in (rel[:, c, 1].min(), np.median(rel[:, c, 1]), rel[:, c, 1].max())] for c in range(5)],
            "last_warp_done_min_med_max": [[round(float(x) / 1e3, 2) for x in (rel[:, c, 2].min(), np.median(rel[:, c, 2]), rel[:, c, 2].max())] for c in range(5)],
            "left_barrier_min_med_max": [[round(float(x) / 1e3, 2) for x in (rel[:, c, 3].min(), np.median(rel[:, c, 3]), rel[:, c, 3].max())] for c in range(5)],
            "next_start_minus_last_done_max": [round(float(t[:, (c + 1) % 5, 0].min() - t[:, c, 2].max()) / 1e3, 2) for c in range(4)],
        }
        np.save("gpurun_out/sor_trace.npy", t)
    if a.others:
        import torch
        def tm(fn, n=5):
            fn(); g.synchronize()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            g.synchronize()
            return (time.perf_counter() - t) / n * 1e3
        out["ms"] = {"geom": tm(lambda: g.update_general_ice_model_data(0.0)), "sia": tm(g.solve_SIA), "thk": tm(lambda: g.calculate_ice_thickness_change(1e-3)),
                     "cfl": tm(g.determine_timesteps), "prepare": tm(g.ssa_prepare), "visc": tm(g.ssa_viscosity), "setup": tm(g.ssa_sliding_and_setup),
                     "sor_1iter_launch": tm(lambda: g.ssa_sor(max_inner=1, force_iters=True)), "finish": tm(g.ssa_finish)}
    if a.thermo and world == 1:
        import numpy as np
        from oracle.oracle import Oracle
        st = S.state_thermo_dome(m, benchmark="EISMINT_1", H0=3000.0, R0=0.7 * c["half_width"])
        gt = IceModelGPU(m, benchmark="EISMINT_1", device=local, thermo=True)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            gt.upload(k, st[k])
        gt.update_general_ice_model_data(0.0)
        def tmt(fn, n=5):
            fn(); gt.synchronize()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            gt.synchronize()
            return (time.perf_counter() - t) / n * 1e3
        out["thermo_ms"] = {"update_ice_temperature": tmt(gt.update_ice_temperature), "solve_SIA_3D_uv": tmt(gt.solve_SIA_3D),
                            "w3d": tmt(gt.thermo_w3d), "heat": tmt(gt.thermo_heat)}
        o = Oracle(m, benchmark="EISMINT_1", nthreads=os.cpu_count())
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            o[k][:] = st[k]
        o.update_general_ice_model_data(0.0)
        t = time.perf_counter(); o.update_ice_temperature(); out["thermo_ms"]["cpu_oracle_update_ice_temperature"] = (time.perf_counter() - t) * 1e3
        out["thermo_ms"]["cpu_threads"] = os.cpu_count()
        nz = 15
        out["thermo_algorithmic_GB"] = {"w3d": m.nV * (3 * nz * 8 + 60) / 1e9, "heat": m.nV * ((2 + 3) * nz * 8 + 12 * 8 + 100 + 2 * nz * 8) / 1e9}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
