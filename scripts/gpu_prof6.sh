#!/bin/bash
# per-launch device times of the INT8-digit fit+predict pipeline at N=16384 (128x128 grid so ncu's serialised
# replays stay short), plain run first
mkdir -p gpurun_out
P="python scripts/prof_fit.py --n 16384 --side 128 --reps 1 --i8 --zorder"
$P > gpurun_out/p16k_i8_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit16384_i8.csv $P > gpurun_out/ncu_fit16k_i8.log 2>&1
echo "rc=$?"; cat gpurun_out/p16k_i8_plain.log
