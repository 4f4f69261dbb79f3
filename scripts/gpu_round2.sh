#!/bin/bash
# instance bench + ncu launch list of the graph-replayed bench step
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bench_instance.py 8 5 > gpurun_out/instance.log 2> gpurun_out/instance_layers.log; tail -c 900 gpurun_out/instance.log | cut -c1-700; cat gpurun_out/instance_layers.log | tail -22
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -c 300 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
