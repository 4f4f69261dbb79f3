#!/bin/bash
# Run on the B200 box (through gpurun).  Every phase under its own timeout and in its own process,
# so a trapped kernel (sticky CUDA error) cannot take the other phases down.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run hbm_tests 600 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_voxel_sample.py -q -m gpu -x
run conv_tiny 300 python -m pytest tests/test_gpu_conv3d.py -q -m gpu -x -k tiny
run conv_all 900 python -m pytest tests/test_gpu_conv3d.py -q -m gpu
