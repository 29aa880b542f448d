#!/bin/bash
# occupancy variants of k_geom_ac / k_geom_aa2 / k_sia_ac (tools/build_variant.py): routine totals per variant
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for v in gac4 gac5 gac6 gac8 gaa8 gaa4 sac4 sac6 sac8 gac4; do
  lib=ufemism_b200/variants/libufemism_b200_$v.so
  UFM_B200_LIB=$PWD/$lib timeout 200 python tools/sor_probe.py --iters 5 --reps 1 --others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'geom', round(d['ms']['geom'],4), 'sia', round(d['ms']['sia'],4))"
done | tee $OUT/r02_minb2_variants.log
