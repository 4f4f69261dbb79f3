#!/bin/bash
# peer-memory halo path: parity test, then the stress leg with peer push vs NCCL (A/B), tight timeouts
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x -k "peer or world2" 2>&1 | tail -6
BARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-instance --no-gpu-baseline --no-proposals"
for HM in peer nccl; do
  SNVC_STRESS_HALO=$HM timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N $BARGS > gpurun_out/r2_peer_n${N}_$HM.json 2> gpurun_out/r2_peer_n${N}_$HM.err
  echo "halo=$HM rc=$?"; grep '^{' gpurun_out/r2_peer_n${N}_$HM.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['stress']; print(s['ms_per_volume'], s['launch'][:20], s['parallelism'][:90])"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_peer_n${N}_$HM.err | tail -4 | cut -c1-300
done
