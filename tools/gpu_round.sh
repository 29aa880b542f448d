#!/bin/bash
# One gpurun call = the whole evidence set of a round, named per round, so that no GPU-minute goes to box start-up twice:
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r02a'            # suite + bench + launch list + SOR / viscosity ncu captures
#   gpurun --timeout 300 -- 'bash tools/gpu_round.sh r02a quick'      # suite + bench only
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/ and describe it in profiles/README.md.
# Numbers printed by the runs under ncu are never bench values.
set -u
TAG=${1:?tag, e.g. r02a}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1

step() { echo "[gpu_round] $(date +%T) $*"; }

step "GPU test suite"
timeout 600 python -m pytest tests -m gpu -x -q --durations=10 > $OUT/${TAG}_gpu_suite.log 2>&1
echo "rc=$?" >> $OUT/${TAG}_gpu_suite.log
tail -3 $OUT/${TAG}_gpu_suite.log

step "bench, 1 GPU (CUDA events; the round's bench line)"
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err
echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_1gpu.json

[ "$MODE" = quick ] && exit 0

step "launch list of the bench command (per-kernel shares; cold-cache, serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_${TAG}_bench_steps2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
echo "rc=$?"

step "ncu --set full: SOR sweep (10 forced iterations in one launch, 1 M vertices)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ssa_sor -s 1 -c 1 -f -o $OUT/sor_${TAG} \
  python tools/sor_probe.py --iters 10 --reps 1 > $OUT/${TAG}_ncu_sor.log 2>&1
echo "rc=$?"
ncu -i $OUT/sor_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python - "$OUT/sor_${TAG}_raw.txt" <<'EOF'
import csv, sys
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active", "smsp__average_warp", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__throughput", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active")
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(sys.argv[1], "w") as f:
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name",) or any(h.startswith(k) for k in keep):
                f.write(f"{h} [{u}] = {v}\n")
EOF

step "experimental tests (skipped in the default suite): x-band row order must be bit-identical"
UFM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k experimental > $OUT/${TAG}_gpu_experimental.log 2>&1
echo "rc=$?"; tail -3 $OUT/${TAG}_gpu_experimental.log

step "row-order probe: default (degree, Morton) vs x-bands (UFM_ROW_ORDER), 100 forced SOR iterations, checksum of U,V must agree"
for ORDER in default bands:16 bands:64 bands:256; do
  if [ "$ORDER" = default ]; then unset UFM_ROW_ORDER; else export UFM_ROW_ORDER=$ORDER; fi
  timeout 300 python tools/sor_probe.py --iters 100 --reps 2 --checksum > $OUT/${TAG}_sor_probe_order_${ORDER/:/}.json 2> $OUT/${TAG}_sor_probe_order_${ORDER/:/}.err
  echo "$ORDER rc=$?"; cut -c1-400 $OUT/${TAG}_sor_probe_order_${ORDER/:/}.json
done
unset UFM_ROW_ORDER

step "ncu --set full: fused viscosity + setup kernel (inside a real solve)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ssa_viscosity -s 2 -c 1 -f -o $OUT/visc_${TAG} \
  python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_visc.log 2>&1
echo "rc=$?"
ls -la $OUT | grep "$TAG"
