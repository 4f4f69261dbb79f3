#!/bin/bash
# same-box A/B of the bench step: kw+kd-fused (default) vs kd-fused (SNVC_CONV_MODE=kd), alternating
set -u
mkdir -p gpurun_out
for i in 1 2; do
  for m in kw kd; do
    if [ $m = kd ]; then export SNVC_CONV_MODE=kd; else unset SNVC_CONV_MODE; fi
    timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/ab_${m}_$i.log 2>&1
    python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/ab_${m}_$i.log') if l.startswith('{')][0])
print('$m', $i, 'pairs/s %.0f  ms/step %.3f  trunk %.3f ms  conv1 %.0f TF  cv %.3f lift %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['stages']['trunk']['ms_per_step'], d['roofline']['achieved'], d['stages']['cost_volume']['ms_per_step'], d['stages']['lift']['ms_per_step'], d['e2e']['value']), d['clocks']['reasons'])
PY
  done
done
