#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
UFM_B200_LIB=$PWD/ufemism_b200/variants/libufemism_b200_stats.so timeout 300 python tools/df_stats_probe.py 1000000 > $OUT/r02e_stats_1M.json 2> $OUT/r02e_stats_1M.err
echo "rc=$?"; cat $OUT/r02e_stats_1M.json; tail -3 $OUT/r02e_stats_1M.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sor or dataflow" > $OUT/r02e_tests.log 2>&1
echo "tests rc=$?"; tail -5 $OUT/r02e_tests.log
