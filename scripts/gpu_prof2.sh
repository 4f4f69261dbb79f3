#!/bin/bash
# tests + bench + ncu --set full of selected kernels
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run tests_gpu 1200 python -m pytest tests -q -m gpu ${PYTEST_ARGS:--x}
run bench 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
TAILN=2 run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
TAILN=3 run ncu_full 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-lift_|deconv}" -s ${SKIP:-12} -c ${COUNT:-4} -f -o gpurun_out/prof_sel python bench.py --steps 1 --warmup 3 --no-cpu-baseline
