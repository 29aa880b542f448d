#!/usr/bin/env python
"""CPU model of the dataflow SOR sweep's stage plan (k_ssa_sor_df / k_sor_need in csrc/ufm_ssa.cu): device row order through the
library's own ufm_plan_row_order, the stage of every slice for a grid of `--warps` warps, need[] per slice, and how far ahead of a
slice's own position its dependencies lie (slack, in stages).  No GPU needed.

python tools/sor_dataflow_plan.py --nv 250000 [--bands 64] [--warps 4736]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def morton16(x, y):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        for sh, msk in ((8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
            v = (v | (v << np.uint64(sh))) & np.uint64(msk)
        return v
    fx = np.clip((x - x.min()) * (65535.0 / (x.max() - x.min())), 0, 65535)
    fy = np.clip((y - y.min()) * (65535.0 / (y.max() - y.min())), 0, 65535)
    return (spread(fx) | (spread(fy) << np.uint64(1))).astype(np.uint32)


def plan(m, n_bands=64, window=4096, warps=4736):
    from ufemism_b200 import capi

    L = capi.load_library()
    M = m.nVAaAc
    XY = np.asarray(m.VAaAc)
    is_edge = np.concatenate([m.edge_index, m.edge_index_Ac]) > 0
    block = np.where(is_edge, 6, np.asarray(m.colour)).astype(np.uint8)
    zeros, ones = np.zeros(M, np.uint8), np.ones(M, np.uint8)
    deg = np.asarray(m.nCAaAc).astype(np.uint8)
    mort = morton16(XY[:, 0], XY[:, 1])
    X = np.ascontiguousarray(XY[:, 0])
    order = np.empty(M, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.ufm_plan_row_order.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    assert L.ufm_plan_row_order(M, p(block), p(zeros), p(zeros), p(ones), p(deg), p(mort), p(X), n_bands, window, p(order)) == 0
    # device positions: every block padded to 256 rows
    pos = np.full(M, -1, np.int64)
    rng, cur = [], 0
    for b in range(1, 7):
        rows = order[block[order] == b]
        pos[rows] = cur + np.arange(len(rows))
        rng.append((cur // 32, (cur + len(rows) + 31) // 32))
        cur = (cur + len(rows) + 255) // 256 * 256
    st_base, st_K, st_act, base = [], [], [], 0
    for c in range(5):
        n = rng[c][1] - rng[c][0]
        K = max(1, -(-n // warps)); act = max(1, -(-n // K))
        st_base.append(base); st_K.append(K); st_act.append(act); base += K
    stage_of_pos = np.full(cur, -1, np.int64)
    for c in range(5):
        s = np.arange(rng[c][0], rng[c][1])
        st = st_base[c] + (s - rng[c][0]) // st_act[c]
        stage_of_pos[rng[c][0] * 32: rng[c][1] * 32] = np.repeat(st, 32)
    colour = np.asarray(m.colour)
    nb = np.asarray(m.CAaAc) - 1
    swept = np.flatnonzero(~is_edge)
    out = {"n_stages": base, "K": st_K, "act": st_act, "slices_per_colour": [r[1] - r[0] for r in rng[:5]]}
    my_stage = stage_of_pos[pos[swept]]
    need_row = np.full(len(swept), -1, np.int64)
    for c in range(nb.shape[1]):
        valid = c < deg[swept]
        j = nb[swept, c]
        ok = valid & (~is_edge[np.where(valid, j, 0)]) & (colour[np.where(valid, j, 0)] < colour[swept])
        need_row = np.maximum(need_row, np.where(ok, stage_of_pos[pos[np.where(valid, j, 0)]], -1))
    sl = pos[swept] // 32
    n_slices = cur // 32
    need = np.full(n_slices, -1, np.int64)
    np.maximum.at(need, sl, need_row)
    own = np.full(n_slices, -1, np.int64)
    own[sl] = my_stage
    has = (own >= 0) & (need >= 0)
    slack = own[has] - need[has]
    out["slack_stages"] = {"min": int(slack.min()), "p1": float(np.percentile(slack, 1)), "p10": float(np.percentile(slack, 10)), "median": float(np.median(slack)),
                           "max": int(slack.max()), "frac_slack_le_1": float(np.mean(slack <= 1)), "frac_slack_le_2": float(np.mean(slack <= 2))}
    for c in range(1, 5):
        selc = has & (own >= st_base[c]) & (own < st_base[c] + st_K[c])
        out[f"colour{c + 1}_slack_min_med"] = [int((own[selc] - need[selc]).min()), float(np.median(own[selc] - need[selc]))]
    return out, dict(pos=pos, need=need, own=own, rng=rng, st_base=st_base, st_act=st_act)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=250000)
    ap.add_argument("--bands", type=int, default=64)
    ap.add_argument("--window", type=int, default=4096)
    ap.add_argument("--warps", type=int, default=148 * 32)
    a = ap.parse_args()
    import bench

    m, _ = bench.build_workload(a.nv)
    out, _ = plan(m, a.bands, a.window, a.warps)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
