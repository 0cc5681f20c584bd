#!/bin/bash
mkdir -p gpurun_out
./scripts/micro/fp64_peak | tee gpurun_out/fp64_peak.log
python scripts/prof_fit.py --n 4096 > gpurun_out/prof_fit4096_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit4096.csv python scripts/prof_fit.py --n 4096 --reps 1 > gpurun_out/ncu_fit.log 2>&1
echo "fit ncu rc=$?"; cat gpurun_out/prof_fit4096_plain.log
