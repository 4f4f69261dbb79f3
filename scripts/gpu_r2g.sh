#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/cv_sweep.py > gpurun_out/r2g_cv_sweep.log 2>&1; cat gpurun_out/r2g_cv_sweep.log | tail -10
timeout 600 python -m pytest tests/test_gpu_nms.py tests/test_gpu_conv2d.py tests/test_gpu_ref_pin.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python scripts/rpn_times.py > gpurun_out/r2g_rpn_times.log 2>&1; grep "graph replay\|kept\|avgpool\|nms_" gpurun_out/r2g_rpn_times.log | cut -c1-200
