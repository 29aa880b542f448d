#!/usr/bin/env python
"""Can half the SMs absorb the bandwidth the other half leaves unused at its grid barriers?

Two handles on ONE GPU, each sweeping the same 1 M-vertex system with a 74-CTA grid (UFM_SOR_GRID), launched concurrently from two host
threads on their own streams: independent sweeps drift out of step by themselves, so whenever one is at a barrier / ramping the other is
mid-phase.  Compared with one handle on all 148 SMs doing the same total work back to back.  If the pair is not faster than that, a
second CTA group half a phase out of step cannot be either."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def prepared(m, st, **kw):
    from ufemism_b200.capi import IceModelGPU

    g = IceModelGPU(m, benchmark=st["benchmark"], device=0, use_analytical_GL_flux=1, exact_xy=1, **kw)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    g.update_general_ice_model_data(0.0)
    g.ssa_prepare(); g.ssa_viscosity(); g.ssa_sliding_and_setup()
    g.ssa_sor(max_inner=5, force_iters=True)
    return g


def main():
    import torch

    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S

    iters = int(os.environ.get("PAIR_ITERS", "100"))
    c = S.CONFIG3
    m = M.square_mesh_with_nv(c["half_width"], int(os.environ.get("PAIR_NV", "1000000")))
    st = S.state_ssa_icestream(m, Hb=c["Hb"], H_shelf=c["H_shelf"])
    out = {"nV": m.nV, "iters": iters}
    full = prepared(m, st)

    def one(g):
        g.reset_counters(); g.ssa_sor(max_inner=iters, force_iters=True); cn = g.counters()
        return cn.sor_ms * 1e3 / cn.sor_iterations
    out["one_handle_148_ctas_us_per_iteration"] = [one(full) for _ in range(3)]
    for grid in (74, 111):
        os.environ["UFM_SOR_GRID"] = str(grid)
        a = prepared(m, st)
        out[f"one_handle_{grid}_ctas_alone_us_per_iteration"] = [one(a) for _ in range(2)]
        if grid == 74:
            b = prepared(m, st)
            res = []
            for rep in range(4):
                t = [None, None]
                bar = threading.Barrier(2)

                def run(i, g):
                    bar.wait()
                    t[i] = one(g)
                th = [threading.Thread(target=run, args=(i, g)) for i, g in enumerate((a, b))]
                torch.cuda.synchronize()
                w0 = time.perf_counter()
                for x in th: x.start()
                for x in th: x.join()
                torch.cuda.synchronize()
                wall = (time.perf_counter() - w0) * 1e6 / iters
                res.append({"each_us_per_iteration": t, "wall_us_per_pair_of_iterations": wall})
            out["two_handles_74_ctas_each_concurrently"] = res
            b.close()
        a.close()
    del os.environ["UFM_SOR_GRID"]
    s = out["one_handle_148_ctas_us_per_iteration"]
    out["reading"] = ("two sweeps back to back on the full grid = 2 x %.1f us; a concurrent pair that takes less per pair of iterations shows the headroom staggering could use"
                      % min(s))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
