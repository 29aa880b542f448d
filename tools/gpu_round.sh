#!/bin/bash
# One gpurun call = the whole 1-GPU evidence set of a round, named per round:
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02z'            # suite + both bench arms + launch list + ncu captures + upload probe
#   gpurun --timeout 900  -- 'bash tools/gpu_round.sh r02z quick'      # suite + bench only
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/ and describe it in profiles/README.md.
# Numbers printed by the runs under ncu are never bench values.
set -u
TAG=${1:?tag, e.g. r02z}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
step() { echo "[gpu_round] $(date +%T) $*"; }

step "GPU test suite"
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > $OUT/${TAG}_gpu_suite.log 2>&1
echo "rc=$?" >> $OUT/${TAG}_gpu_suite.log
tail -3 $OUT/${TAG}_gpu_suite.log

step "smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

step "bench, reference arm (the CPU restatement run for real, as the driver runs it: --steps 20 --warmup 5)"
( time timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ) 2>&1 | grep real
cut -c1-400 $OUT/${TAG}_bench_reference.json

step "bench, 1 GPU (--steps 20 --warmup 5; CUDA events; the round's bench line)"
( time timeout 1700 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err ) 2>&1 | grep real
echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_1gpu.json; tail -8 $OUT/${TAG}_bench_1gpu.err

[ "$MODE" = quick ] && exit 0

step "launch list of the bench command (per-kernel shares; cold-cache, serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_${TAG}_bench_steps2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_ncu_bench.log 2>&1
echo "rc=$?"

capture() {  # name, kernel regex (demangled name incl. template arguments), skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 -f -o $OUT/${name}_${TAG} "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_raw_summary.py $OUT/${name}_${TAG}.ncu-rep > $OUT/${name}_${TAG}_raw.txt 2>/dev/null
}
step "ncu --set full: SOR sweep (10 forced iterations in one launch, 1 M vertices)"
capture sor ".*k_ssa_sor.*" 1 python tools/sor_probe.py --iters 10 --reps 1
step "ncu --set full: per-step kernels (one launch each, 1 M vertices, warm state)"
capture visc ".*k_ssa_viscosity.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture geom_ac ".*k_geom_ac.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture geom_aa2 ".*k_geom_aa2.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture sia_ac ".*k_sia_ac\\(.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture sia_aa ".*k_sia_aa.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture thk_flux ".*k_thk_flux.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture thk1 ".*k_thk<\\(int\\)1, \\(bool\\)1>.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others
capture thk2 ".*k_thk<\\(int\\)2, \\(bool\\)1>.*" 2 python tools/sor_probe.py --iters 5 --reps 1 --others

step "upload probe (mesh update / restart data path, 1 M vertices)"
timeout 900 python tools/upload_probe.py --out $OUT/upload_probe_1M_${TAG}.json 2> $OUT/upload_probe_1M_${TAG}_phases.log | cut -c1-600
ls -la $OUT | grep "$TAG" | head -60
