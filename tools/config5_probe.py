#!/usr/bin/env python
"""BASELINE configs[4]: ~4 M-vertex mesh (16 M combined-mesh rows), hybrid SIA + SSA with the analytical grounding-line flux, vertex-partitioned
over the GPUs of one node (x-strips, NVLink halo exchange).  Run alone (1 GPU) or under torchrun (2/4/8 GPUs):

    python tools/config5_probe.py --out profiles/config5_4M_1gpu_r02.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/config5_probe.py

The mesh goes to the library as PRIMARY data (ufm_mesh_upload_primary derives the rest), so no 11 GB of neighbour functions are built on the host."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(nv=4000000, warmup=2, steps=4, sor_iters=40, dist=None, torch=None, rank=0, world=1, local=0, log=lambda *a: None):
    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU

    c = S.CONFIG3
    t = time.time()
    m = M.primary_mesh_with_nv(c["half_width"], nv)
    st = S.state_ssa_icestream(m, scale=1.0, Hb=c["Hb"], H_shelf=c["H_shelf"])
    t_mesh = time.time() - t
    t = time.time()
    g = IceModelGPU(m, benchmark=st["benchmark"], device=local, rank=rank, nranks=world, primary_only=True, use_analytical_GL_flux=1)
    dev = torch.device("cuda", local) if torch is not None else None
    if world > 1:
        g.connect(dist, device=dev)
    t_up = time.time() - t
    log(f"[config5] rank {rank}: mesh nV={m.nV} built in {t_mesh:.1f}s, upload from primary data {t_up:.1f}s")
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    r = g.region(0.0)
    g.run_model(r, 1e12, max_steps=warmup)
    g.synchronize()
    if world > 1:
        dist.barrier()
    g.reset_counters()
    t0, tm0 = time.perf_counter(), r.time
    g.run_model(r, 1e12, max_steps=steps)
    g.synchronize()
    wall = time.perf_counter() - t0
    cn = g.counters()
    if world > 1:
        tt = torch.tensor([wall], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt.item())
    us_solve = cn.sor_ms * 1e3 / max(cn.sor_iterations, 1)
    g.ssa_sor(max_inner=5, force_iters=True)
    g.reset_counters()
    g.ssa_sor(max_inner=sor_iters, force_iters=True)
    cf = g.counters()
    us_forced = cf.sor_ms * 1e3 / max(cf.sor_iterations, 1)
    if world > 1:
        tt = torch.tensor([us_forced, us_solve], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        us_forced, us_solve = (float(x) for x in tt.tolist())
    _, part = g.owners()
    out = {"workload": f"config5_ssa_icestream_nV{m.nV}_AaAc{m.nVAaAc}_MISMIP_mod_GLflux", "n_gpus": world, "per_step_kernels_partitioned": bool(part),
           "steps": steps, "warmup": warmup, "ms_per_step": wall / steps * 1e3, "model_yr_per_wall_hr": (r.time - tm0) / wall * 3600.0,
           "n_sor": int(cn.sor_iterations), "sor_us_per_iteration_in_solve": us_solve, "sor_us_per_iteration_forced": us_forced,
           "sor_algorithmic_GB_per_iteration": cf.sor_bytes_per_iteration / 1e9, "sor_aggregate_GBps": cf.sor_bytes_per_iteration / (us_forced * 1e-6) / 1e9,
           "mesh_build_s_host": t_mesh, "upload_from_primary_s": t_up, "model_time": float(r.time)}
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=4000000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(a.nv, dist=dist, torch=torch, rank=rank, world=world, local=local, log=lambda *x: print(*x, file=sys.stderr, flush=True))
    if rank == 0:
        print(json.dumps(out), flush=True)
        if a.out:
            json.dump(out, open(a.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
