#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02last; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_suite.log 2>&1; tail -2 $OUT/${TAG}_gpu_suite.log
timeout 300 python tools/sor_probe.py --iters 5 --reps 1 --others > $OUT/${TAG}_others.json 2> $OUT/${TAG}_others.err; cut -c1-900 $OUT/${TAG}_others.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_others.csv python tools/sor_probe.py --iters 5 --reps 1 --others > /dev/null 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['frac'],3) for k,v in d['roofline_kernels'].items() if isinstance(v,dict)}, d['ssa']['ms_per_step_without_ssa_solve'])
PY
