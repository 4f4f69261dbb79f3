#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== staged"; timeout 300 python scripts/layer_times.py 2>&1 | tail -18
echo "=== direct"; SNVC_CONV_STORE=direct timeout 300 python scripts/layer_times.py 2>&1 | tail -18
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:kdfuse" -s 5 -c 1 -f -o gpurun_out/prof_kd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kd.log 2>&1; tail -2 gpurun_out/ncu_kd.log
