#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r02h_gpu_suite.log 2>&1
echo "suite rc=$?"; tail -14 $OUT/r02h_gpu_suite.log
timeout 1200 python bench.py --steps 8 --warmup 3 > $OUT/r02h_bench_1gpu.json 2> $OUT/r02h_bench_1gpu.err
echo "bench rc=$?"; tail -12 $OUT/r02h_bench_1gpu.err; cut -c1-600 $OUT/r02h_bench_1gpu.json
