#!/bin/bash
mkdir -p gpurun_out
P="python scripts/prof_fit.py --n 16384 --side 64 --reps 1"
$P > gpurun_out/p16k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit16384.csv $P > gpurun_out/ncu_fit16k.log 2>&1
echo "rc=$?"; cat gpurun_out/p16k_plain.log
