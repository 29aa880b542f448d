mkdir -p gpurun_out
for c in -1 0 25 -1 0; do
  echo "carveout=$c" >> gpurun_out/r01i_l1.log
  UFM_L1_CARVEOUT=$c timeout 200 python tools/sor_probe.py --iters 100 --reps 3 --others >> gpurun_out/r01i_l1.log 2>> gpurun_out/r01i_l1.err
done
cat gpurun_out/r01i_l1.log
tail -2 gpurun_out/r01i_l1.err
