#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02k_tests.log 2>&1; echo "suite rc=$?"; tail -6 $OUT/r02k_tests.log
timeout 600 python tools/parity_probe.py --nv 250000 --steps 6 > $OUT/r02k_parity_250k.log 2>&1; tail -6 $OUT/r02k_parity_250k.log | cut -c1-700
