#!/bin/bash
# tests + bench + launch list + ncu --set full capture of ONE whole step (every hot-path kernel once)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; lscpu | head -20 >> gpurun_out/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then run tests_gpu 1200 python -m pytest tests -q -m gpu -x --durations=8; fi
run bench 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-}
TAILN=2 run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
KREGEX=${KREGEX:-"conv3d_|lift_|cv_nd"}
TAILN=3 run ncu_full 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s ${SKIP:-45} -c ${COUNT:-15} -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls -la gpurun_out/
