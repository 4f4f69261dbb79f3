#!/bin/bash
# ncu --set full of the trunk stragglers: conv6 (deconv 64->32), dres1.conv2 (CTA-pair kernel with residual), hg.conv3 (per-tap)
set -u
mkdir -p gpurun_out
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals"
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > gpurun_out/r02y_ncu_$name.log 2>&1
  echo "$name exit=$?"
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/r02y_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/r02y_${name}_src.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r02y_${name}_raw.csv 2>/dev/null | cut -c1-230
}
cap deconv "conv3d_deconv_kernel" 9 3 python bench.py $BARGS
cap resid 'kdpair_kernel<\(int\)2, \(int\)64, \(bool\)1' 3 1 python bench.py $BARGS
cap pertap "conv3d_tcgen05_kernel" 3 1 python bench.py $BARGS
