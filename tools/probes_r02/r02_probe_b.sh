#!/bin/bash
# round 2: dataflow SOR kernel -- parity tests first, then speed against the barrier kernel
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sor or solve_SSA or run_model or dataflow" --durations=5 > $OUT/r02b_tests.log 2>&1
echo "tests rc=$?"; tail -15 $OUT/r02b_tests.log
for DF in 1 0; do
  UFM_SOR_DATAFLOW=$DF timeout 300 python tools/sor_probe.py --iters 100 --reps 3 --checksum > $OUT/r02b_df$DF.json 2> $OUT/r02b_df$DF.err
  echo "dataflow=$DF rc=$?"; cut -c1-700 $OUT/r02b_df$DF.json; tail -3 $OUT/r02b_df$DF.err
done
UFM_SOR_DATAFLOW=1 timeout 300 python tools/sor_probe.py --iters 3 --reps 20 > $OUT/r02b_df1_3iters.json 2>&1; cut -c1-500 $OUT/r02b_df1_3iters.json
UFM_SOR_DATAFLOW=0 timeout 300 python tools/sor_probe.py --iters 3 --reps 20 > $OUT/r02b_df0_3iters.json 2>&1; cut -c1-500 $OUT/r02b_df0_3iters.json
