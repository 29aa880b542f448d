"""Partitioned (multi-GPU) SSA solve vs the single-GPU solve: run under torchrun, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Storage is replicated, work is partitioned: every rank computes its own x-strip -- the SSA solve (rows a neighbour strip reads are pushed by
the sweep kernel over NVLink) and the per-step kernels (thickness update, geometry, SIA, yield stress, critical time steps; thickness,
out-flux factors and edge velocities of the strip boundary exchanged once per kernel).  Colour sweeps are order-free, reductions use a
rank-independent fixed tree or are exact (min, integer sum) -> results must be BIT-IDENTICAL to the single-GPU run, with identical
iteration counts and time steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    nv = int(os.environ.get("UFM_CHECK_NV", "10000"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = M.square_mesh_with_nv(750e3, nv)
    st = S.state_ssa_icestream(m, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0)
    ok = True
    for gl in (1, 0):
        g = IceModelGPU(m, benchmark=st["benchmark"], device=local, rank=rank, nranks=world, use_analytical_GL_flux=gl)
        g.connect(dist, device=torch.device("cuda", local))
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
            g.upload(k, st[k])
        r = g.region(0.0)
        g.run_model(r, 1e12, max_steps=3)
        dev = torch.device("cuda", local)
        # partitioned per-step kernels: a rank's download is valid for the elements it owns; the global field is put together by owner
        res = {f: g.download_global(dist, f, device=dev) for f in ("Hi", "U_SSA", "V_SSA", "Up_SSA_Ac", "U_SIA", "Hs", "dHs_dx", "mask", "mask_gl_Ac", "D_SIA", "dHi_dt")}
        counts = (r.n_steps, r.n_ssa, r.n_outer_total, r.n_sor_total, r.time)
        if rank == 0:
            print(f"gl={gl}: per-step kernels partitioned: {g.owners()[1]}", flush=True)
        # every rank must arrive at the same complete answer
        for f, a in res.items():
            t = torch.from_numpy(a.copy()).cuda()
            ref = t.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(t, ref):
                ok = False
                print(f"[rank {rank}] gl={gl} field {f} differs from rank 0", flush=True)
        dist.barrier()
        if rank == 0:
            g1 = IceModelGPU(m, benchmark=st["benchmark"], device=local, use_analytical_GL_flux=gl)
            for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
                g1.upload(k, st[k])
            r1 = g1.region(0.0)
            g1.run_model(r1, 1e12, max_steps=3)
            c1 = (r1.n_steps, r1.n_ssa, r1.n_outer_total, r1.n_sor_total, r1.time)
            if c1 != counts:
                ok = False
                print(f"gl={gl} counts differ: partitioned {counts} vs single {c1}", flush=True)
            for f, a in res.items():
                b = g1.download(f)
                if not np.array_equal(a, b):
                    ok = False
                    print(f"gl={gl} field {f}: partitioned != single GPU, max abs diff {np.abs(a - b).max()}", flush=True)
            print(f"gl={gl}: {world} ranks, counts {counts}, |U|max {np.abs(res['U_SSA']).max():.3f}", flush=True)
            g1.close()
        dist.barrier()
        g.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTI_GPU_CHECK " + ("PASS" if flag.item() == 0 else "FAIL"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
