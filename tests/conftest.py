import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
