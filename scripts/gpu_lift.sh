#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_voxel_sample.py tests/test_gpu_parallel.py tests/test_gpu_fullsize.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -4
BARGS="--steps 10 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals"
for m in fast coop; do if [ $m = fast ]; then unset SNVC_LIFT_MODE; else export SNVC_LIFT_MODE=$m; fi; timeout 200 python bench.py $BARGS 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; st=dict(r['stages']); print('$m', round(d['value'],1), round(d['ms_per_step'],4), 'lift', round(st['lift']['ms_per_step'],4), round(st['lift']['frac'],3), 'cv', round(st['cost_volume']['frac'],3), 'e2e', round(d['e2e']['value'],1))"; done
