#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
N=${1:-8}
for NV in 10000 100000; do
UFM_CHECK_NV=$NV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/r02q_check_${N}gpu_$NV.log 2>&1
echo "check nv=$NV rc=$?"; grep -E "MULTI_GPU_CHECK|differs|gl=|Error|error" $OUT/r02q_check_${N}gpu_$NV.log | tail -8
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02q_bench_${N}gpu.json 2> $OUT/r02q_bench_${N}gpu.err
echo "bench rc=$?"; tail -4 $OUT/r02q_bench_${N}gpu.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/r02q_bench_${N}gpu.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','partitioned_bit_identical')}, d['e2e']['value'], d['roofline']['us_per_iteration'], d['ssa'].get('ms_per_step_without_ssa_solve'))
    print(d.get('config5_4M'))
    print(d.get('independent_regions_mode'))
except Exception as e: print('no line', e)
PY
