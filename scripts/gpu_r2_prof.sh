#!/bin/bash
# final profiling pass of round 2: launch list of one bench step + ncu --set full of the dominant conv and the HBM kernels.
# raw / source CSV exports come back in gpurun_out/ (the .ncu-rep files stay on the box).
set -u
mkdir -p gpurun_out
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv python bench.py $BARGS > gpurun_out/r02z_launches.log 2>&1
echo "launch list exit=$?"
cap() {  # name, regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > gpurun_out/r02z_ncu_$name.log 2>&1
  echo "$name exit=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/r02z_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/r02z_${name}_src.csv 2>/dev/null
}
cap conv1 "kdpair_kernel<2, 64, 0, 42, 1>" 3 1 python bench.py $BARGS
cap cv_split "cv_split_bf16_kernel" 6 2 python bench.py $BARGS
cap lift "lift_fast_bf16_kernel" 3 1 python bench.py $BARGS
cap roi "roi_sample_fast_bf16_kernel" 2 1 python scripts/bench_instance.py 8 2
for f in gpurun_out/r02z_*_raw.csv; do python scripts/ncu_summary.py $f 2>/dev/null | cut -c1-220; done
python scripts/sass_summary.py > gpurun_out/r02z_sass_summary.txt 2>&1; tail -5 gpurun_out/r02z_sass_summary.txt
