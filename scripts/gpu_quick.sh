#!/bin/bash
# quick GPU loop: selected tests + bench (+ optional extra command)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run tests_gpu 1200 python -m pytest tests -q -m gpu ${PYTEST_ARGS:--x}
if [ -n "${EXTRA:-}" ]; then TAILN=40 run extra 600 bash -c "$EXTRA"; fi
run bench 600 python bench.py --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS:-}
if [ "${LAUNCHES:-1}" = "1" ]; then TAILN=2 run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline; fi
