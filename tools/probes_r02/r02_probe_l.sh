#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > $OUT/r02l_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -12 $OUT/r02l_gpu_suite.log
timeout 1500 python bench.py --steps 8 --warmup 3 > $OUT/r02l_bench_1gpu.json 2> $OUT/r02l_bench_1gpu.err
echo "bench rc=$?"; tail -8 $OUT/r02l_bench_1gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_iteration'])
print('parity', d.get('parity'))
print('cpu', d.get('cpu_baseline',{}).get('value'))
oc=d.get('other_configs',{})
for k,v in oc.items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','wall_s','model_yr_per_wall_hr','parity','error','mesh_update','vs_Halfar_solution','steps')})
PY
