#!/bin/bash
# weak-scaling point: bench.py under torchrun on N GPUs of one box (N = $1)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.log 2>&1
grep '^{' gpurun_out/scale_$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=%d pairs/s %.0f  ms/step %.3f  e2e %.0f  conv1 %.0f TF  trunk %.3f ms' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['stages']['trunk']['ms_per_step']), d['clocks']['reasons'])"
