#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxel_sample.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -4
for m in fast fast32 v3; do if [ $m = fast ]; then unset SNVC_ROI_MODE; else export SNVC_ROI_MODE=$m; fi; timeout 300 python scripts/bench_instance.py 8 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ('layers','workload')})"; done
unset SNVC_ROI_MODE
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_sample_fast -s 2 -c 1 -f -o /tmp/prof_roi python scripts/bench_instance.py 8 2 > gpurun_out/r02b_ncu_roi.log 2>&1
ncu -i /tmp/prof_roi.ncu-rep --page raw --csv > gpurun_out/r02b_roi_raw.csv 2>/dev/null
ncu -i /tmp/prof_roi.ncu-rep --page source --csv > gpurun_out/r02b_roi_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02b_roi_raw.csv
