#!/bin/bash
# barrier kernel: Neumann rows on idle warps (base), + header prefetch (variant); phase trace; then the GPU suite
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for v in base prefetch base; do
  lib=ufemism_b200/libufemism_b200.so; [ "$v" != base ] && lib=ufemism_b200/variants/libufemism_b200_$v.so
  UFM_B200_LIB=$PWD/$lib timeout 200 python tools/sor_probe.py --iters 100 --reps 3 --checksum --others > $OUT/r02g_$v.json 2> $OUT/r02g_$v.err
  echo "variant=$v rc=$?"; cut -c1-900 $OUT/r02g_$v.json
done
UFM_SOR_TRACE=1 UFM_SOR_TRACE_VARIANT=UFM_SOR_CHUNK=1,UFM_SOR_FUSE_BC=1,UFM_SOR_BAR=1 timeout 300 python tools/sor_probe.py --iters 50 --reps 1 > $OUT/r02g_trace.json 2> $OUT/r02g_trace.err
echo "trace rc=$?"; cut -c1-1500 $OUT/r02g_trace.json
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r02g_gpu_suite.log 2>&1
echo "suite rc=$?"; tail -15 $OUT/r02g_gpu_suite.log
