#!/bin/bash
# round 2, call A: new parity pins + host-return probe + same-box GPU baselines + bench
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python scripts/d2h_probe.py > gpurun_out/r2_d2h_probe_n1.log 2>&1; tail -12 gpurun_out/r2_d2h_probe_n1.log
timeout 900 python -m pytest tests/test_gpu_ref_pin.py tests/test_gpu_host_return.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2_new_tests.log; tail -30 gpurun_out/r2_new_tests.log
timeout 900 python -m pytest tests -q -m gpu --ignore=tests/test_gpu_ref_pin.py --ignore=tests/test_gpu_host_return.py --ignore=tests/test_gpu_fullsize.py 2>&1 | tail -40 > gpurun_out/r2_all_tests.log; tail -15 gpurun_out/r2_all_tests.log
timeout 600 python scripts/gpu_baselines.py > gpurun_out/r2_gpu_baselines.log 2>&1; tail -3 gpurun_out/r2_gpu_baselines.log
timeout 300 python bench.py > gpurun_out/r2_bench_a.log 2>&1; tail -2 gpurun_out/r2_bench_a.log
