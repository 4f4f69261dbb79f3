#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv2d.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r2b_conv2d.log; tail -25 gpurun_out/r2b_conv2d.log
timeout 900 python -m pytest tests/test_gpu_ref_pin.py tests/test_gpu_host_return.py tests/test_gpu_fullsize.py tests/test_gpu_cost_volume.py tests/test_gpu_models.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2b_tests.log; tail -25 gpurun_out/r2b_tests.log
