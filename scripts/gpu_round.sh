#!/bin/bash
# Full GPU check: parity tests, smoke, bench (both arms), ncu launch list.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run tests_gpu 1500 python -m pytest tests -q -m gpu -x
run smoke 300 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 10 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
TAILN=5 run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
