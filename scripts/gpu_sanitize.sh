#!/bin/bash
# compute-sanitizer passes over the GPU tests of the hand-written HBM kernels (memcheck) and racecheck (shared-memory hazards) on
# the gather kernels that hand their set-up over through shared memory
set -u
mkdir -p gpurun_out
T="tests/test_gpu_cost_volume.py tests/test_gpu_voxel_sample.py tests/test_gpu_nms.py tests/test_gpu_depth_head.py tests/test_gpu_host_return.py tests/test_gpu_grid_proj.py"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest $T -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_tests.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck_tests.log | tail -3
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_voxel_sample.py tests/test_gpu_cost_volume.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck_tests.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/r02_sanitizer_racecheck_tests.log | tail -3
