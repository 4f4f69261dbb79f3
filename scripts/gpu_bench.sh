#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run tests_gpu 900 python -m pytest tests -q -m gpu ${PYTEST_ARGS:-}
run bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=2 run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
