#!/bin/bash
set -u
timeout 400 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_ref_pin.py tests/test_gpu_fullsize.py tests/test_gpu_models.py tests/test_gpu_parallel.py -q -m gpu -x 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; st=dict(r['stages']); print(round(d['value'],1), round(d['ms_per_step'],4), 'cv', round(st['cost_volume']['ms_per_step'],4), round(st['cost_volume']['frac'],3), 'lift', round(st['lift']['frac'],3))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals > /dev/null 2>&1
grep "cv_" gpurun_out/r02w_launches.csv | tail -2 | cut -d, -f5,15 | cut -c1-120
