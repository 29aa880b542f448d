"""Golden vectors produced by the REFERENCE'S OWN SOURCE (run from the repo root, where /root/reference is mounted:
python -m tests.golden.make_reference_source_golden).

The Fortran cannot be compiled in this image; oracle/f90py.py translates the text of the hot-path routines of /root/reference/src
statement by statement into Python and tests/ref_source.py runs them on the golden mesh.  What is stored are the outputs of those
translated routines -- not of the oracle.  tests/test_reference_source.py compares the oracle, the mesh substrate and (on the B200)
the CUDA path against them, and, where the reference is mounted, re-runs the translation live."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def digest(a):
    a = np.ascontiguousarray(np.asarray(a))
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    from tests import ref_cases as RC
    from tests.test_reference_source import run_reference_source

    out = run_reference_source(RC.golden_mesh())
    store = {}
    for k, a in out.items():
        a = np.asarray(a)
        store["sha256__" + k] = np.array(digest(a))
        if a.size <= 40000:
            store[k] = a
    path = os.path.join(HERE, "reference_source_600.npz")
    np.savez_compressed(path, **store)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
