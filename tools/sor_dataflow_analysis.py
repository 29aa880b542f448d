"""Design study for a barrier-free SOR sweep; CPU only, no GPU needed.

For the device row order of the AaAc mesh (colour-major; inside a colour either the current (degree, Morton) order, Morton-major with
the degree sorted only inside windows of `--window` rows, or x-bands with Morton order inside a band) it reports

* the sliced-ELL padding (stored entries / real entries) -- what a layout costs in coefficient bytes;
* for chunks of `--chunk` rows, how many chunks of lower colours a chunk depends on, and the pipeline lag: how far (as a fraction of
  its colour block) colour c must have progressed before the chunk at fraction f of colour c+1 can start.  lag <= small means the
  ramp-down of colour c can overlap the ramp-up of colour c+1; lag ~ 1 means a barrier in all but name.

python tools/sor_dataflow_analysis.py --nv 250000
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def morton(x, y):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    fx = np.clip((x - x.min()) / (x.max() - x.min()) * 65535.0, 0, 65535)
    fy = np.clip((y - y.min()) / (y.max() - y.min()) * 65535.0, 0, 65535)
    return spread(fx) | (spread(fy) << np.uint64(1))


def analyse(m, layout, window, chunk):
    M = m.nVAaAc
    XY = np.asarray(m.VAaAc)
    mort = morton(XY[:, 0], XY[:, 1])
    deg = np.asarray(m.nCAaAc).astype(np.int64)
    colour = np.asarray(m.colour).astype(np.int64)
    is_edge = np.concatenate([m.edge_index, m.edge_index_Ac]) > 0
    rows = np.flatnonzero(~is_edge)
    if layout == "degree_morton":                       # what ufm_mesh_upload does today
        order = rows[np.lexsort((mort[rows], deg[rows], colour[rows]))]
    else:                                               # (x-band,) Morton-major, degree sorted inside windows of `window` rows
        band = np.zeros(M, np.int64)
        if layout.startswith("banded"):
            nb_ = int(layout.split("_")[1])
            band = np.minimum((XY[:, 0] - XY[:, 0].min()) / (XY[:, 0].max() - XY[:, 0].min()) * nb_, nb_ - 1).astype(np.int64)
        o1 = rows[np.lexsort((mort[rows], band[rows], colour[rows]))]
        pos_in_colour = np.zeros(len(o1), np.int64)
        for c in range(1, 6):
            sel = colour[o1] == c
            pos_in_colour[sel] = np.arange(sel.sum())
        order = o1[np.lexsort((deg[o1], pos_in_colour // window, colour[o1]))]
    pos = np.full(M, -1, np.int64)
    pos[order] = np.arange(len(order))
    # colour blocks start on slice boundaries in the device layout; model that
    starts, out, p = {}, {}, 0
    newpos = np.full(M, -1, np.int64)
    for c in range(1, 6):
        blk = order[colour[order] == c]
        starts[c] = p
        newpos[blk] = p + np.arange(len(blk))
        p += (len(blk) + 31) // 32 * 32
    n_slots = p
    # padding of sliced ELL: slice width = max degree in the slice
    d_slot = np.zeros(n_slots, np.int64)
    d_slot[newpos[order]] = deg[order]
    w = d_slot.reshape(-1, 32).max(axis=1)
    out["ell_padding"] = float((w * 32).sum() / deg[order].sum())
    # chunk dependencies
    chunk_of = newpos // chunk
    size_c = {c: int((colour[order] == c).sum()) for c in range(1, 6)}
    frac = np.zeros(M)
    for c in range(1, 6):
        blk = order[colour[order] == c]
        frac[blk] = (newpos[blk] - starts[c]) / max(size_c[c], 1)
    nb = np.asarray(m.CAaAc) - 1
    n_dep, lag = [], []
    for c in range(2, 6):
        blk = order[colour[order] == c]
        for k in range(0, len(blk), chunk):
            r = blk[k:k + chunk]
            nbs = nb[r]
            valid = (np.arange(nb.shape[1])[None, :] < deg[r][:, None])
            j = nbs[valid]
            j = j[(~is_edge[j]) & (colour[j] == c - 1)]           # the binding dependency: the colour just before
            if len(j) == 0:
                continue
            n_dep.append(len(np.unique(chunk_of[j])))
            lag.append(float(frac[j].max() - frac[r].min()))      # progress of colour c-1 needed beyond this chunk's own position
    out.update(chunk_rows=chunk, n_chunks=len(n_dep), deps_mean=float(np.mean(n_dep)), deps_max=int(np.max(n_dep)),
               lag_median=float(np.median(lag)), lag_p95=float(np.percentile(lag, 95)), lag_max=float(np.max(lag)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=250000)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--window", type=int, default=4096)
    args = ap.parse_args()
    import bench

    m, _ = bench.build_workload(args.nv)
    res = {"nV": int(m.nV), "nVAaAc": int(m.nVAaAc)}
    res["degree_morton (current)"] = analyse(m, "degree_morton", args.window, args.chunk)
    for wdw in (1024, args.window, 16384):
        res[f"morton_major_window_{wdw}"] = analyse(m, "morton", wdw, args.chunk)
    for nb_ in (16, 64, 256):
        res[f"x_bands_{nb_}_morton_inside_window_{args.window}"] = analyse(m, f"banded_{nb_}", args.window, args.chunk)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
