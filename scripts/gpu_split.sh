#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run split_tests 600 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_conv3d.py tests/test_gpu_models.py tests/test_gpu_parallel.py -q -m gpu -x
show() { python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/$1.log') if l.startswith('{')][0])
print('$1', 'pairs/s %.0f  ms/step %.3f  trunk %.3f ms (%.0f TF exec)  conv1 %.0f TF %.3f ms  cv %.3f (%.0f GB/s) lift %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step'], d['stages']['trunk']['achieved_tflops'], d['roofline']['achieved'], d['roofline']['share_of_step']*d['ms_per_step'], d['stages']['cost_volume']['ms_per_step'], d['stages']['cost_volume']['achieved_gbs'], d['stages']['lift']['ms_per_step'], d['e2e']['value']))
PY
}
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sp_split_$i.log 2>&1; show sp_split_$i
  SNVC_SPLIT_CV=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sp_full_$i.log 2>&1; show sp_full_$i
done
