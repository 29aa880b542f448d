import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """Devices the driver API sees (no torch import, no context created); 0 when there is no driver at all."""
    import ctypes

    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing in ufm_create.  When the gpu tests were asked
    for by name (`-m gpu`) nothing is skipped: on a GPU box a missing device must fail loudly, not pass as 'skipped'."""
    expr = config.getoption("markexpr", "") or ""
    if "gpu" in expr and "not gpu" not in expr:
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu tests run on the B200 box with -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_MESH_CACHE = {}


def get_mesh(nv, half_width=750e3, seed=20211103, order="random"):
    from ufemism_b200 import mesh as M

    key = (nv, half_width, seed, order)
    if key not in _MESH_CACHE:
        _MESH_CACHE[key] = M.square_mesh_with_nv(half_width, nv, seed=seed, order=order)
    return _MESH_CACHE[key]


@pytest.fixture(scope="session")
def mesh_2k():
    return get_mesh(2000)


@pytest.fixture(scope="session")
def mesh_10k():
    return get_mesh(10000)


def fan_mesh(nv=2000, hw=750e3, fans=((-200e3, 100e3, 10), (250e3, -150e3, 14), (0.0, 300e3, 16))):
    """A lattice mesh with three 'fans': vertices of degree 10, 14 and 16 (= nC_mem, the most the reference's arrays hold).
    Rows wider than 8 take the generic (non-unrolled) paths of the kernels, which ordinary Delaunay meshes never reach."""
    import numpy as np

    from ufemism_b200 import mesh as M

    area = (2 * hw) ** 2
    h = np.sqrt(2 * area / (np.sqrt(3) * nv))
    pts = M.make_points(-hw, hw, -hw, hw, h, seed=7)
    corners, rest = pts[:4], pts[4:]
    extra = []
    for cx, cy, k in fans:
        r0 = 1.6 * h
        rest = rest[np.hypot(rest[:, 0] - cx, rest[:, 1] - cy) > r0]
        ang = 2 * np.pi * np.arange(k) / k
        extra.append(np.array([[cx, cy]]))
        extra.append(np.stack([cx + 0.45 * r0 * np.cos(ang), cy + 0.45 * r0 * np.sin(ang)], 1))
        extra.append(np.stack([cx + 0.8 * r0 * np.cos(ang + np.pi / k), cy + 0.8 * r0 * np.sin(ang + np.pi / k)], 1))
    return M.build_mesh(np.concatenate([corners, rest] + extra), -hw, hw, -hw, hw)


@pytest.fixture(scope="session")
def mesh_fan():
    return fan_mesh()
