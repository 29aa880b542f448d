"""N > 1 host-side logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from ufemism_b200 import capi

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # 1. the IPC-blob exchange used by IceModelGPU.connect: every rank ends up with all blobs in rank order
        blob = bytes([rank + 1]) * capi.IceModelGPU.COMM_BLOB_BYTES
        blobs = capi.exchange_blobs(dist, blob, world)
        assert [b[0] for b in blobs] == [r + 1 for r in range(world)] and all(len(b) == 256 for b in blobs)
        # 2. max-over-ranks timing
        assert capi.max_over_ranks(dist, 10.0 * (rank + 1)) == 10.0 * world
        # 3. every rank computes the same partition plan from the same mesh (deterministic, no communication)
        from ufemism_b200 import mesh as M
        m = M.square_mesh_with_nv(750e3, 1500, seed=3)
        own = capi.partition_owners(m, world)
        import torch
        t = torch.from_numpy(own.copy())
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        q.put((rank, "ok", int((own == rank).sum())))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"fail: {e!r}", 0))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    counts = sorted(r[2] for r in res)
    assert abs(counts[0] - counts[1]) <= 1  # strips balanced by row count


@pytest.mark.parametrize("P", [2, 4, 8])
def test_partition_owners_are_balanced_x_strips(P):
    sys.path.insert(0, ROOT)
    from ufemism_b200 import capi
    from ufemism_b200 import mesh as M

    m = M.square_mesh_with_nv(750e3, 3000, seed=5)
    own = capi.partition_owners(m, P)
    cnt = np.bincount(own, minlength=P)
    assert cnt.max() - cnt.min() <= 1 and cnt.sum() == m.nVAaAc
    x = m.VAaAc[:, 0]
    for r in range(P - 1):  # strips are ordered in x
        assert x[own == r].max() <= x[own == r + 1].min()
    # the halo is thin: only a small fraction of rows has a neighbour in another strip
    nb = m.CAaAc - 1
    cross = 0
    for ai in range(m.nVAaAc):
        n = m.nCAaAc[ai]
        if (own[nb[ai, :n]] != own[ai]).any():
            cross += 1
    assert cross < 0.35 * m.nVAaAc * (P / 8) + 0.1 * m.nVAaAc


@pytest.mark.parametrize("P", [2, 3, 8])
def test_halo_plan_of_the_partitioned_per_step_kernels(P):
    """ufm_partition_halo_counts (host only) = the lists ufm_mesh_upload builds for the partitioned per-step kernels: rank s sends rank q
    every own Aa vertex that q reads (a neighbour of one of q's vertices, or one of the four vertices of one of q's staggered vertices)
    and every own staggered vertex on a connection of one of q's vertices.  Checked against a numpy restatement from C / iAci / Aci;
    x-strips only talk to their neighbour strips, and the traffic is a boundary effect (O(sqrt N))."""
    import numpy as np

    from tests.conftest import get_mesh
    from ufemism_b200 import capi

    m = get_mesh(10000)
    own = capi.partition_owners(m, P).astype(np.int64)
    oa, oc = own[: m.nV], own[m.nV:]
    ca, cc = capi.partition_halo_counts(m, P)
    rd_aa = np.zeros((m.nV, P), bool); rd_ac = np.zeros((m.nAc, P), bool)
    for c in range(m.nC_mem):
        has = c < m.nC
        u, a = m.C[has, c] - 1, m.iAci[has, c] - 1
        rd_aa[u, oa[has]] = True
        rd_ac[a, oa[has]] = True
    for k in range(4):
        rd_aa[m.Aci[:, k] - 1, oc] = True
    want_a, want_c = np.zeros((P, P), np.int64), np.zeros((P, P), np.int64)
    for s in range(P):
        for q in range(P):
            if s != q:
                want_a[s, q] = np.sum(rd_aa[oa == s, q]); want_c[s, q] = np.sum(rd_ac[oc == s, q])
    assert np.array_equal(ca, want_a) and np.array_equal(cc, want_c)
    far = np.abs(np.subtract.outer(np.arange(P), np.arange(P))) > 1
    # strips exchange with their neighbour strips only -- except vertex 1, which every domain-boundary staggered vertex names as its
    # fourth vertex with a zero coefficient (src/mesh_ArakawaC_module.f90:165-167)
    assert (ca[far] <= 1).all() and not cc[far].any()
    assert 0 < ca.sum() < 0.1 * m.nV * P and 0 < cc.sum() < 0.1 * m.nAc * P   # a boundary effect
