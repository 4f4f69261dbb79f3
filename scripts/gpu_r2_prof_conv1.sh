#!/bin/bash
set -u
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kdpair_kernel<\(int\)2, \(int\)64, \(bool\)0, \(int\)42, \(bool\)1>' -s 3 -c 1 -f -o /tmp/prof_conv1 python bench.py $BARGS > gpurun_out/r02z_ncu_conv1.log 2>&1
echo "conv1 exit=$?"
ncu -i /tmp/prof_conv1.ncu-rep --page raw --csv > gpurun_out/r02z_conv1_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv1.ncu-rep --page source --csv > gpurun_out/r02z_conv1_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02z_conv1_raw.csv | cut -c1-230
