#!/bin/bash
# slab path: parity tests, then the bench's stress leg on N GPUs (graph replay and eager)
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x 2>&1 | tail -4
BARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-instance --no-gpu-baseline --no-proposals"
for G in 1 0; do
  SNVC_STRESS_GRAPH=$G timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$G bench.py --gpus $N $BARGS > gpurun_out/r2_stress_n${N}_g$G.json 2> gpurun_out/r2_stress_n${N}_g$G.err
  echo "graph=$G rc=$?"; grep '^{' gpurun_out/r2_stress_n${N}_g$G.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stress']); print('value', d['value'], 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/r2_stress_n${N}_g$G.err | cut -c1-300
done
