#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxel_sample.py -q -m gpu -x 2>&1 | tail -4
for m in v3 coop1; do SNVC_ROI_MODE=$m timeout 300 python scripts/bench_instance.py 8 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ('layers','workload')})"; done
