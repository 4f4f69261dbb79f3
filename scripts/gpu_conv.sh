#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 40 gpurun_out/$name.log; }
run conv_tiny 300 python -m pytest tests/test_gpu_conv3d.py -q -m gpu -x -k tiny
run conv_all 900 python -m pytest tests/test_gpu_conv3d.py -q -m gpu
