#!/usr/bin/env python
"""Build a tuning variant of libufemism_b200.so with extra -D flags (only the listed sources are recompiled; the other objects of
the regular build are reused) into ufemism_b200/variants/.  Select it at run time with UFM_B200_LIB=<path> (capi.py)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufemism_b200 import build as B


def main():
    name, flags = sys.argv[1], [f for f in sys.argv[2:] if f.startswith("-D")]
    srcs = [f for f in sys.argv[2:] if not f.startswith("-D")] or ["ufm_ssa.cu"]
    B.build_all()
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for s in B.CU_SOURCES + B.HOST_SOURCES:
        o = os.path.join(B.CSRC, os.path.splitext(s)[0] + ".o")
        if s in srcs:
            o = os.path.join(out_dir, f"{os.path.splitext(s)[0]}_{name}.o")
            subprocess.run([B.NVCC] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, s), "-o", o], check=True)
        objs.append(o)
    lib = os.path.join(out_dir, f"libufemism_b200_{name}.so")
    subprocess.run([B.NVCC, "-shared", "-o", lib, "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp"] + objs, check=True)
    print(lib)


if __name__ == "__main__":
    main()
