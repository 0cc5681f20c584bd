#!/bin/bash
mkdir -p gpurun_out
P="python scripts/prof_fit.py --n 640 --side 32 --reps 1"
$P > gpurun_out/p640_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potf2inv -s 2 -c 1 -o gpurun_out/prof_potf2b $P > gpurun_out/ncu_potf2b.log 2>&1
echo "rc=$?"
