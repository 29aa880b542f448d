#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for v in sleep200 sleep1000; do
UFM_SOR_DATAFLOW=1 UFM_B200_LIB=$PWD/ufemism_b200/variants/libufemism_b200_$v.so timeout 300 python tools/df_stats_probe.py 1000000 > $OUT/r02f_$v.json 2> $OUT/r02f_$v.err
echo "$v rc=$?"; cat $OUT/r02f_$v.json; tail -2 $OUT/r02f_$v.err
done
