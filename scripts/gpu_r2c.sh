#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2c_tests.log; tail -12 gpurun_out/r2c_tests.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 6000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
