#!/bin/bash
# A/B of tuning builds (tools/build_variant.py) in one GPU call: tools/probe_variants.sh <tag> <variant names...>
tag=$1; shift
mkdir -p gpurun_out
for v in base "$@" base; do
  lib=ufemism_b200/libufemism_b200.so; [ "$v" != base ] && lib=ufemism_b200/variants/libufemism_b200_$v.so
  echo "variant=$v" >> gpurun_out/${tag}.log
  UFM_B200_LIB=$PWD/$lib timeout 200 python tools/sor_probe.py --iters 100 --reps 3 --others >> gpurun_out/${tag}.log 2>> gpurun_out/${tag}.err
done
cat gpurun_out/${tag}.log; tail -2 gpurun_out/${tag}.err
