#!/bin/bash
# iteration loop: conv + model parity tests, per-layer times, graph-replay bench (x2)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run conv_tests 900 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_models.py tests/test_gpu_parallel.py ${MORE_TESTS:-} -q -m gpu -x
TAILN=12 run layers 300 python scripts/layer_times.py
show() { python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/$1.log') if l.startswith('{')][0])
print('$1', 'pairs/s %.0f  ms/step %.3f  trunk %.3f ms (%.0f TF exec)  conv1 %.0f TF  cv %.3f (%.0f GB/s) lift %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step'], d['stages']['trunk']['achieved_tflops'], d['roofline']['achieved'], d['stages']['cost_volume']['ms_per_step'], d['stages']['cost_volume']['achieved_gbs'], d['stages']['lift']['ms_per_step'], d['e2e']['value']))
PY
}
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/it_graph_$i.log 2>&1; show it_graph_$i
done
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
