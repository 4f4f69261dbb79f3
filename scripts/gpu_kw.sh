#!/bin/bash
# kw+kd-fused conv kernel: parity tests, A/B layer times against the kd-fused kernel, bench
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run conv_tests 900 python -m pytest tests/test_gpu_conv3d.py -q -m gpu -x
run model_tests 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parallel.py -q -m gpu -x
TAILN=20 run layers_kw 300 python scripts/layer_times.py
SNVC_CONV_MODE=kd TAILN=20 run layers_kd 300 python scripts/layer_times.py
run bench 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline


