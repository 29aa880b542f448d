#!/bin/bash
# round 2: ncu --set full of the dataflow kernel and of the barrier kernel on the same mesh (10 forced iterations), raw metrics side by side
set -u
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
for DF in 1 0; do
  UFM_SOR_DATAFLOW=$DF timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ssa_sor -s 1 -c 1 -f -o $OUT/r02d_sor_df$DF \
    python tools/sor_probe.py --iters 10 --reps 1 > $OUT/r02d_ncu_df$DF.log 2>&1
  echo "df=$DF rc=$?"
  ncu -i $OUT/r02d_sor_df$DF.ncu-rep --page raw --csv > $OUT/r02d_sor_df${DF}_raw.csv 2>/dev/null
done
ls -la $OUT | grep r02d
