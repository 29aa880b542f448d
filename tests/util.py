"""Shared helpers for the parity tests: build an oracle state and a GPU state from one scenario."""
import numpy as np

INPUTS = ("Hi", "Hb", "SL", "SMB_year", "BMB")


def make_oracle(mesh, state, nthreads=1, **cfg):
    from oracle.oracle import Oracle

    o = Oracle(mesh, benchmark=state["benchmark"], nthreads=nthreads, **cfg)
    for k in INPUTS:
        o[k][:] = state[k]
    return o


def make_gpu(mesh, state, **params):
    from ufemism_b200.capi import IceModelGPU

    g = IceModelGPU(mesh, benchmark=state["benchmark"], **params)
    for k in INPUTS:
        g.upload(k, state[k])
    return g


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))


def assert_bits_equal(a, b, name=""):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        bad = np.flatnonzero(~same.ravel())
        i = bad[0]
        raise AssertionError(f"{name}: {bad.size} of {a.size} elements differ; first at {i}: {a.ravel()[i]!r} vs {b.ravel()[i]!r}")
