#!/bin/bash
# residual-tile experiment: cold-start repeats of the conv tests (fresh process each), then layer times / bench
set -u
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_gpu_conv3d.py -q -m gpu -x -k "staged_tma_store or cta_pair or kitti_level or residual_modes" 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_models.py tests/test_gpu_parallel.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python scripts/layer_times.py 2>&1 | tail -12
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pairs/s %.0f  ms/step %.3f  trunk %.3f' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step']))"; done
