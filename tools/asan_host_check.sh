#!/bin/bash
# Host-side C / C++ of the product library under AddressSanitizer (+ UBSan for the C file): the mesh substrate (make_Ac_mesh,
# make_combined_AaAc_mesh, Voronoi geometry, five-colouring with its scratch block) on four meshes incl. one with degree-16 vertices,
# the derivation of the secondary mesh data with reuse of the previous mesh's buffers, the NetCDF restart / help_fields writer and reader,
# the host twin of the device pow / tan and the row-order planner (tests/test_restart_files.py, tests/test_abi.py).  No GPU needed.
set -eu
cd "$(dirname "$0")/.."
OUT=gpurun_out/asan; mkdir -p $OUT
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -ffp-contract=off -fPIC -shared -o $OUT/libufm_mesh.so ufemism_b200/csrc/mesh_host.c -lm
python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
from ufemism_b200 import build as B
B.build_all()
out = "gpurun_out/asan"
objs = []
for s in B.CU_SOURCES + B.HOST_SOURCES:
    o = os.path.join(B.CSRC, os.path.splitext(s)[0] + ".o")
    if s in ("ufm_mesh_primary.cpp", "mesh_host.c", "ufm_netcdf.cpp", "ufm_pow_host.cpp"):
        o = os.path.join(out, os.path.splitext(s)[0] + "_asan.o")
        flags = list(B.NVCC_FLAGS)
        i = flags.index("-Xcompiler"); flags[i + 1] += ",-fsanitize=address,-fno-omit-frame-pointer,-g"
        subprocess.run([B.NVCC] + [f if f != "-O3" else "-O1" for f in flags] + ["-c", os.path.join(B.CSRC, s), "-o", o], check=True)
    objs.append(o)
subprocess.run([B.NVCC, "-shared", "-o", os.path.join(out, "libufemism_b200_asan.so"), "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp", "-Wno-deprecated-gpu-targets"] + objs, check=True)
PY
cat > $OUT/substrate.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import ufemism_b200.mesh as M
M._LIB_PATH = os.path.join(os.getcwd(), "gpurun_out/asan/libufm_mesh.so")
M.build_mesh_lib = lambda force=False: M._LIB_PATH
for nv, seed in ((600, 1), (3000, 2), (20000, 3)):
    m = M.square_mesh_with_nv(750e3, nv, seed=seed)
    print("mesh", m.nV, m.nAc, int(m.colour_nV.sum()), flush=True)
from conftest import fan_mesh
m = fan_mesh()
print("fan mesh", m.nV, "max degree", int(m.nC.max()))
PY
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
LD_PRELOAD=$ASAN:$UBSAN ASAN_OPTIONS=detect_leaks=0 python $OUT/substrate.py
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0 UFM_B200_LIB=$PWD/$OUT/libufemism_b200_asan.so python -m pytest tests/test_restart_files.py tests/test_abi.py -q -x -k "not live and not cpp_host and not build_flags and not kernel_resources and not citations" -p no:cacheprovider
# the CPU oracle itself (test infrastructure, but every parity claim leans on it): tests/test_oracle.py against an ASan / UBSan build
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -ffp-contract=off -fno-fast-math -fPIC -shared -o $OUT/libufm_oracle.so oracle/ufm_oracle.c -lm
cat > $OUT/run_oracle.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import oracle.oracle as O
O._LIB = os.path.join(os.getcwd(), "gpurun_out/asan/libufm_oracle.so")
O.build = lambda force=False: O._LIB
import pytest
sys.exit(pytest.main(["tests/test_oracle.py", "-q", "-x", "-p", "no:cacheprovider", "-k", "not gcc_code_generation"]))
PY
LD_PRELOAD=$ASAN:$UBSAN ASAN_OPTIONS=detect_leaks=0 python $OUT/run_oracle.py
echo "asan_host_check: clean"
