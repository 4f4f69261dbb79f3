#!/bin/bash
# full validation of the tree: all GPU tests, smoke, both bench arms
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_tests.log 2>&1; tail -5 gpurun_out/r2h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 600 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
