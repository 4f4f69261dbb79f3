#!/bin/bash
# N-GPU bench (N = number of visible GPUs)
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r2_n${N}_topo.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err; tail -c 1500 gpurun_out/r2_n${N}_bench.json; tail -3 gpurun_out/r2_n${N}_bench.err
