"""Build the in-tree native artefacts (explicit nvcc / gcc, no JIT cache):

* ``libufemism_b200.so`` -- CUDA kernels + C ABI, sm_100a only, ``-fmad=false`` (bit-level parity contract)
* ``libufm_mesh.so``     -- CPU mesh substrate (mesh creation stays on the CPU per north_star)
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libufemism_b200.so")
CU_SOURCES = ["ufm_api.cu", "ufm_upload.cu", "ufm_ssa.cu", "ufm_geom.cu", "ufm_thermo.cu"]
# host-only sources of the same library (compiled by nvcc's host compiler): restart / help_fields files, secondary mesh data
HOST_SOURCES = ["ufm_netcdf.cpp", "ufm_mesh_primary.cpp", "ufm_pow_host.cpp", "mesh_host.c"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fopenmp", "-ccbin", "/usr/bin/g++",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES + HOST_SOURCES]
    deps = srcs + [os.path.join(CSRC, "ufm_internal.cuh"), os.path.join(CSRC, "ufm_pow.cuh"), os.path.join(HERE, "..", "include", "ufemism_b200.h")]
    if force or _stale(LIB, deps):
        objs = []
        for s in srcs:
            o = os.path.splitext(s)[0] + ".o"
            if force or _stale(o, [s] + deps[len(srcs):]):
                cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
                subprocess.run(cmd, check=True)
            objs.append(o)
        # the arch is given at link time too: without it nvcc adds an empty default-arch (sm_52) fatbin to the library
        subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp"] + objs, check=True)
    return LIB


def build_all(force=False, verbose=False):
    from . import mesh

    mesh.build_mesh_lib(force=force)
    return build_cuda(force=force, verbose=verbose)


if __name__ == "__main__":
    import sys

    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
