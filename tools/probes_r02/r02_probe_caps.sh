#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02j; mkdir -p $OUT; export PYTHONUNBUFFERED=1
capture() {
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 -f -o $OUT/${name}_${TAG} "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "$name rc=$?"
  python tools/ncu_raw_summary.py $OUT/${name}_${TAG}.ncu-rep > $OUT/${name}_${TAG}_raw.txt 2>/dev/null
}
P="python tools/sor_probe.py --iters 5 --reps 1 --others"
capture sia_aa ".*k_sia_aa.*" 2 $P
capture thk1 ".*k_thk<\(int\)1, \(bool\)1>.*" 2 $P
capture thk2 ".*k_thk<\(int\)2, \(bool\)1>.*" 2 $P
grep -E "gpu__time_duration|registers_per_thread \[|warps_active.avg.pct" $OUT/*_${TAG}_raw.txt
