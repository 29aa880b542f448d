#!/bin/bash
# occupancy / batch-size variants of k_sia_aa and k_thk (tools/build_variant.py): routine totals per variant
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for v in base aa4 aa5 aa4c2 aa6c2 aac8 thk5 thk6 base; do
  lib=ufemism_b200/libufemism_b200.so; [ "$v" != base ] && lib=ufemism_b200/variants/libufemism_b200_$v.so
  UFM_B200_LIB=$PWD/$lib timeout 200 python tools/sor_probe.py --iters 5 --reps 1 --others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'sia', round(d['ms']['sia'],4), 'thk', round(d['ms']['thk'],4))"
done | tee $OUT/r02_minb_variants.log
