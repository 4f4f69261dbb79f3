#!/bin/bash
# final validation of the tree on N GPUs: (N = 1) all GPU tests + smoke; both bench arms
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2z_tests.log 2>&1; tail -3 gpurun_out/r2z_tests.log
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
  timeout 300 python bench.py --impl reference > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err; tail -c 400 gpurun_out/r2z_bench_reference_arm.json
  timeout 600 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 300 gpurun_out/r2z_bench_n1.json; tail -2 gpurun_out/r2z_bench_n1.err
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/r2z_bench_n${N}.json 2> gpurun_out/r2z_bench_n${N}.err; tail -c 300 gpurun_out/r2z_bench_n${N}.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2z_bench_n${N}.err | tail -3
fi
