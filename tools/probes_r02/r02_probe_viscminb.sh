#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for v in 4 3 5 6 4; do
  UFM_VISC_MINB=$v timeout 200 python tools/sor_probe.py --iters 5 --reps 1 --others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('visc_minb=$v', 'visc', round(d['ms']['visc'],4), 'prepare', round(d['ms']['prepare'],4), 'geom', round(d['ms']['geom'],4))"
done | tee $OUT/r02_viscminb.log
