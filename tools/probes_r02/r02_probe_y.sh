#!/bin/bash
# compute-sanitizer (memcheck) over the tests that exercise the kernels touched in round 2 (small meshes)
set -u
OUT=gpurun_out; TAG=r02y; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thickness or SIA or run_model_halfar or device_loop or update_general or solve_SSA_full or sor_schedule" > $OUT/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/${TAG}_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thickness or SIA or run_model_halfar" > $OUT/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/${TAG}_racecheck.log | tail -6
