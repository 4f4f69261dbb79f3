#!/bin/bash
# Full GPU check: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the dominant conv kernels.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run tests_gpu 1500 python -m pytest tests -q -m gpu -x
run smoke 300 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 10 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
TAILN=3 run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
TAILN=3 run ncu_pair 900 ncu --set full --clock-control none --import-source on -k regex:kdpair -s 8 -c 4 -f -o gpurun_out/prof_kdpair python bench.py --steps 1 --warmup 3 --no-cpu-baseline
