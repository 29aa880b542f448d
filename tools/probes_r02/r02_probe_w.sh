#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02w; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_suite.log 2>&1; tail -3 $OUT/${TAG}_gpu_suite.log
UFM_UPLOAD_TIMING=1 timeout 900 python tools/upload_probe.py --out $OUT/upload_probe_1M_${TAG}.json 2> $OUT/upload_probe_1M_${TAG}_phases.log | cut -c1-600
timeout 300 python tools/sor_probe.py --iters 5 --reps 1 --others > $OUT/${TAG}_others.json 2> $OUT/${TAG}_others.err; cut -c1-900 $OUT/${TAG}_others.json
