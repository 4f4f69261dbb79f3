#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2e_tests.log; tail -6 gpurun_out/r2e_tests.log
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1500 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
