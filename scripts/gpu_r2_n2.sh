#!/bin/bash
# 2-GPU call: slab parity over NCCL (HaloComm), bench at N=2 (stress + instance legs), D2H probe with / without NUMA binding
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_n2_tests.log; tail -4 gpurun_out/r2_n2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; tail -c 2500 gpurun_out/r2_n2_bench.json; tail -3 gpurun_out/r2_n2_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/d2h_probe.py > gpurun_out/r2_n2_d2h_bind.log 2>&1; grep "rank" gpurun_out/r2_n2_d2h_bind.log | tail -12
SNVC_NUMA_BIND=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/d2h_probe.py > gpurun_out/r2_n2_d2h_nobind.log 2>&1; grep "rank" gpurun_out/r2_n2_d2h_nobind.log | tail -12
