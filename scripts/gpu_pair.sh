#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run pair_test 300 python -m pytest tests/test_gpu_conv3d.py -q -m gpu -x -k "tiny or cta_pair"
if grep -q "passed" gpurun_out/pair_test.log && ! grep -q "failed" gpurun_out/pair_test.log; then
  run conv_tests 600 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_models.py -q -m gpu -x
  TAILN=12 run layers_pair 300 python scripts/layer_times.py
  SNVC_CONV_MODE=kw TAILN=12 run layers_kw 300 python scripts/layer_times.py
  for i in 1 2; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/it_graph_$i.log 2>&1
    python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/it_graph_$i.log') if l.startswith('{')][0])
print('pair $i', 'pairs/s %.0f  ms/step %.3f  trunk %.3f ms  conv1 %.0f TF  cv %.3f lift %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step'], d['roofline']['achieved'], d['stages']['cost_volume']['ms_per_step'], d['stages']['lift']['ms_per_step'], d['e2e']['value']))
PY
  done
fi
