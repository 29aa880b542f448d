#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02up2; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_restart.py -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; tail -2 $OUT/${TAG}_tests.log
UFM_UPLOAD_TIMING=1 timeout 900 python tools/upload_probe.py --out $OUT/upload_probe_1M_${TAG}.json 2> $OUT/upload_probe_1M_${TAG}_phases.log | cut -c1-400
awk '/reupload_from_primary #2/{f=1} /reupload_secondary/{f=0} f' $OUT/upload_probe_1M_${TAG}_phases.log
