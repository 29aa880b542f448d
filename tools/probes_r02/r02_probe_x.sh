#!/bin/bash
# N GPUs, final code of the round: bit identity with the single-GPU run (check script at 10 k and 100 k vertices), then the bench line as the driver runs it
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
N=${1:-2}; TAG=${2:-r02x}; STEPS=${3:-20}; WARM=${4:-5}
for NV in 10000 100000; do
UFM_CHECK_NV=$NV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/${TAG}_check_${N}gpu_$NV.log 2>&1
echo "check nv=$NV rc=$?"; grep -E "MULTI_GPU_CHECK|differs|gl=|Error|error" $OUT/${TAG}_check_${N}gpu_$NV.log | tail -8
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup $WARM ${5:-} > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -4 $OUT/${TAG}_bench_${N}gpu.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/${TAG}_bench_${N}gpu.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','partitioned_bit_identical')}, d['e2e']['value'], d['roofline']['us_per_iteration'], d['ssa'].get('ms_per_step_without_ssa_solve'))
    print(d.get('config5_4M'))
except Exception as e: print('no line', e)
PY
