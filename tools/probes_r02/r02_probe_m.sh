#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > $OUT/r02m_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -12 $OUT/r02m_gpu_suite.log
python - <<'PY'
import sys, time, json
sys.path.insert(0, '.')
import torch, bench
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for f in (bench.leg_config1_eismint, bench.leg_config2_halfar):
    d = f(torch, s, 0, 16)
    print(json.dumps({k: d[k] for k in d if k in ("workload", "steps", "wall_s", "ms_per_step", "device_ms_per_step", "gpu_launches", "gpu_launches_per_step", "parity", "model_yr_per_wall_hr", "cpu_restatement")}))
PY
