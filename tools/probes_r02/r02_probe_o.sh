#!/bin/bash
# N GPUs: partitioned per-step kernels -- bit identity with the single-GPU run (check script at 10 k and 100 k vertices), then the bench line
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
N=${1:-2}
for NV in 10000 100000; do
UFM_CHECK_NV=$NV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/r02o_check_${N}gpu_$NV.log 2>&1
echo "check nv=$NV rc=$?"; grep -E "MULTI_GPU_CHECK|differs|gl=|Error|error" $OUT/r02o_check_${N}gpu_$NV.log | tail -12
done
for PS in 1 0; do
UFM_PARTITION_STEP=$PS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > $OUT/r02o_bench_${N}gpu_ps$PS.json 2> $OUT/r02o_bench_${N}gpu_ps$PS.err
echo "bench part_step=$PS rc=$?"; tail -3 $OUT/r02o_bench_${N}gpu_ps$PS.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/r02o_bench_${N}gpu_ps$PS.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','partitioned_bit_identical')}, d['e2e']['value'], d['roofline']['us_per_iteration'], d['ssa'].get('ms_per_step_without_ssa_solve'))
except Exception as e: print('no line', e)
PY
done
