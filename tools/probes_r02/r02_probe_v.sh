#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02v; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thickness or SIA or run_model or reference_source or device_loop" > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
for e in 1 0; do
  UFM_THK_EDGE=$e timeout 300 python tools/sor_probe.py --iters 5 --reps 1 --others > $OUT/${TAG}_others_edge$e.json 2> $OUT/${TAG}_others_edge$e.err; cut -c1-900 $OUT/${TAG}_others_edge$e.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_edge1.csv python tools/sor_probe.py --iters 5 --reps 1 --others > /dev/null 2>&1
UFM_THK_EDGE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_edge0.csv python tools/sor_probe.py --iters 5 --reps 1 --others > /dev/null 2>&1
ls $OUT | grep $TAG
