#!/bin/bash
# round 2: where does the dataflow kernel's time go?  timing-only variants (NOT correct): no acquire (no L1 invalidation), no release fence, no per-slice reduce
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sor or dataflow" > $OUT/r02c_tests.log 2>&1
echo "tests rc=$?"; tail -8 $OUT/r02c_tests.log
for v in base noacq norel noacqrel noall; do
  lib=ufemism_b200/libufemism_b200.so; [ "$v" != base ] && lib=ufemism_b200/variants/libufemism_b200_$v.so
  UFM_B200_LIB=$PWD/$lib timeout 200 python tools/sor_probe.py --iters 100 --reps 2 --checksum > $OUT/r02c_$v.json 2> $OUT/r02c_$v.err
  echo "variant=$v rc=$?"; cut -c1-420 $OUT/r02c_$v.json
done
