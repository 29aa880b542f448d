#!/usr/bin/env python
"""Host-only half of ufm_mesh_upload_primary (no GPU needed): time ufm_mesh_derive_secondary on a 1 M-vertex mesh, phase by phase
(UFM_UPLOAD_TIMING=1), and check the derived arrays against the mesh substrate's own."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("UFM_UPLOAD_TIMING", "1")


def main():
    from ufemism_b200 import mesh as M
    from ufemism_b200 import capi

    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    t = time.time()
    m = M.square_mesh_with_nv(750e3, nv)
    print(f"mesh {m.nV} vertices built in {time.time() - t:.1f}s", file=sys.stderr)
    t = time.perf_counter()
    d = capi.derive_secondary(m, repeat=int(os.environ.get("DERIVE_REPEAT", "5")))   # the first into a fresh object, the others into the same one
    print(f"--- derive_secondary x repeat: {time.perf_counter() - t:.3f} s", file=sys.stderr)
    ok = True
    for name, ref in (("Aci", m.Aci), ("iAci", m.iAci), ("VAc", m.VAc), ("edge_index_Ac", m.edge_index_Ac), ("nCAaAc", m.nCAaAc), ("CAaAc", m.CAaAc), ("A", m.A), ("Cw", m.Cw), ("colour_vi", m.colour_vi), ("colour_nV", m.colour_nV)):
        if name in d:
            a, b = np.asarray(d[name]), np.asarray(ref)
            same = a.shape == b.shape and np.array_equal(a, b)
            ok &= bool(same)
            print(name, "identical" if same else f"DIFFERENT {a.shape} {b.shape}", file=sys.stderr)
    print("ALL IDENTICAL" if ok else "MISMATCH", file=sys.stderr)


if __name__ == "__main__":
    main()
