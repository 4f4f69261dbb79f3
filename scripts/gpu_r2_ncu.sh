#!/bin/bash
# round 2 profiling call: launch list of one bench step + ncu --set full of the HBM kernels and the dominant conv.
# .ncu-rep files stay on the box (size); raw / source CSV exports come back in gpurun_out/.
set -u
mkdir -p gpurun_out
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-instance --no-stress --no-gpu-baseline --no-proposals"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $BARGS > gpurun_out/r02_launches.log 2>&1
echo "launch list exit=$?"
for K in cv_split lift_ndhwc kdpair; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-6} -c 2 -f -o /tmp/prof_$K python bench.py $BARGS > gpurun_out/r02_ncu_$K.log 2>&1
  echo "$K exit=$?"
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/r02_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$K.ncu-rep --page source --csv > gpurun_out/r02_${K}_src.csv 2>/dev/null
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_sample -s 2 -c 2 -f -o /tmp/prof_roi python scripts/bench_instance.py 8 2 > gpurun_out/r02_ncu_roi.log 2>&1
echo "roi exit=$?"
ncu -i /tmp/prof_roi.ncu-rep --page raw --csv > gpurun_out/r02_roi_raw.csv 2>/dev/null
ncu -i /tmp/prof_roi.ncu-rep --page source --csv > gpurun_out/r02_roi_src.csv 2>/dev/null
ls -la gpurun_out/r02_* | head -30
for f in gpurun_out/r02_*_raw.csv; do python scripts/ncu_summary.py $f 2>/dev/null | head -8; done
timeout 300 python scripts/rpn_times.py > gpurun_out/r02_rpn_times.log 2>&1; tail -40 gpurun_out/r02_rpn_times.log
