#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run model_tests 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x
show() { python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/$1.log') if l.startswith('{')][0])
print('$1', 'pairs/s %.0f  ms/step %.3f  trunk %.3f ms  conv1 %.0f TF  cv %.3f lift %.3f  e2e %.0f launches %d' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step'], d['roofline']['achieved'], d['stages']['cost_volume']['ms_per_step'], d['stages']['lift']['ms_per_step'], d['e2e']['value'], d['gpu_launches']), d['clocks'])
PY
}
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_graph_$i.log 2>&1; show g_graph_$i
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eager > gpurun_out/g_eager_$i.log 2>&1; show g_eager_$i
  SNVC_CONV_MODE=kd timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_graphkd_$i.log 2>&1; show g_graphkd_$i
done
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/g_graph_40.log 2>&1; show g_graph_40
SNVC_CONV_MODE=kd timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/g_graphkd_40.log 2>&1; show g_graphkd_40
tail -3 gpurun_out/g_graph_1.log | cut -c1-600
