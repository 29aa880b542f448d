#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02g; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_restart.py tests/test_gpu_parity.py -m gpu -x -q -k "host or run_model or hybrid or cpp_host" > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/${TAG}_bench_$i.json 2> $OUT/${TAG}_bench_$i.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench_$i.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'])
PY
done
UFM_XFER_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/${TAG}_bench_nooverlap.json 2> $OUT/${TAG}_bench_nooverlap.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench_nooverlap.json') if l.startswith('{')][-1])
print('no overlap', d['value'], d['ms_per_step'], d['e2e'])
PY
