#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 600 python tools/parity_probe.py --nv 250000 --steps 4 > $OUT/r02i_parity_250k.log 2>&1; tail -5 $OUT/r02i_parity_250k.log | cut -c1-900
timeout 900 python tools/parity_probe.py --nv 1000000 --steps 4 > $OUT/r02i_parity_1M.log 2>&1; tail -5 $OUT/r02i_parity_1M.log | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "run_model" > $OUT/r02i_tests.log 2>&1; tail -4 $OUT/r02i_tests.log
