#!/bin/bash
# ncu --set full capture of selected kernels (1 GPU, short command)
set -u
mkdir -p gpurun_out
KREGEX=${KREGEX:-halo}
SKIP=${SKIP:-5}
COUNT=${COUNT:-2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -f -o gpurun_out/prof_$KREGEX python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$KREGEX.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_$KREGEX.log; ls -la gpurun_out/*.ncu-rep
