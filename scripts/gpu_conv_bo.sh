#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
export SNVC_CONV_BO=1
run conv_bo1 600 python -m pytest tests/test_gpu_conv3d.py -q -m gpu
export SNVC_CONV_BO=0
run conv_bo0 600 python -m pytest tests/test_gpu_conv3d.py -q -m gpu
