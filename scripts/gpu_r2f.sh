#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/cv_sweep.py > gpurun_out/r2f_cv_sweep.log 2>&1; cat gpurun_out/r2f_cv_sweep.log | tail -12
timeout 600 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_ref_pin.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python scripts/rpn_times.py > gpurun_out/r2f_rpn_times.log 2>&1; tail -32 gpurun_out/r2f_rpn_times.log
timeout 600 python bench.py --no-instance --no-stress --no-gpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
