#!/bin/bash
# 4 GPUs: partitioned run must stay bit-identical with neighbour-only signalling; bench line at N = 4 with both signalling modes
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/r02n_check_${N}gpu.log 2>&1
echo "check rc=$?"; grep -E "MULTI_GPU_CHECK|differs|gl=" $OUT/r02n_check_${N}gpu.log | tail -6
for ALL in 0 1; do
UFM_PEER_ALL=$ALL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > $OUT/r02n_bench_${N}gpu_all$ALL.json 2> $OUT/r02n_bench_${N}gpu_all$ALL.err
echo "bench all=$ALL rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/r02n_bench_${N}gpu_all$ALL.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','partitioned_bit_identical')}, d['e2e']['value'], d['roofline']['us_per_iteration'])
except Exception as e: print('no line', e)
PY
done
