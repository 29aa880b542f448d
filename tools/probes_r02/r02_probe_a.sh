#!/bin/bash
# round 2, first GPU call: what the x-band row orders do to the CURRENT (barrier) SOR kernel, per-step kernel times, phase trace
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for ORDER in default bands:64 bands:16 bands:256; do
  if [ "$ORDER" = default ]; then unset UFM_ROW_ORDER; else export UFM_ROW_ORDER=$ORDER; fi
  timeout 300 python tools/sor_probe.py --iters 100 --reps 2 --checksum --others > $OUT/r02a_order_${ORDER/:/}.json 2> $OUT/r02a_order_${ORDER/:/}.err
  echo "$ORDER rc=$?"; cut -c1-900 $OUT/r02a_order_${ORDER/:/}.json
done
export UFM_ROW_ORDER=bands:64
UFM_SOR_TRACE=1 UFM_SOR_TRACE_VARIANT=UFM_SOR_CHUNK=1,UFM_SOR_FUSE_BC=1,UFM_SOR_BAR=1 timeout 300 python tools/sor_probe.py --iters 50 --reps 1 > $OUT/r02a_trace_bands64.json 2> $OUT/r02a_trace_bands64.err
echo "trace rc=$?"; cut -c1-1500 $OUT/r02a_trace_bands64.json
