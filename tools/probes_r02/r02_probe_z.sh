#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 600 python tools/sor_imbalance_probe.py > $OUT/r02z_sor_imbalance.json 2> $OUT/r02z_sor_imbalance.err; echo rc=$?; cat $OUT/r02z_sor_imbalance.json; tail -3 $OUT/r02z_sor_imbalance.err
