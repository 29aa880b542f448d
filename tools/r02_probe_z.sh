#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
PAIR_NV=500000 timeout 600 python tools/sor_pair_probe.py > $OUT/r02z_sor_pair_500k.json 2> $OUT/r02z_sor_pair.err; echo rc=$?; cat $OUT/r02z_sor_pair_500k.json; tail -3 $OUT/r02z_sor_pair.err
