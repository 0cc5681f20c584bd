#!/bin/bash
# ncu capture of the INT8 digit TRMM (plain run first, as the recipe requires)
mkdir -p gpurun_out
CMD="python scripts/prof_i8.py 8192 16384 ${1:-8} nocheck ${2:-}"
timeout 120 $CMD > gpurun_out/i8_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_i8 -s 1 -c 1 -f -o gpurun_out/prof_i8${2:-} $CMD > gpurun_out/ncu_i8.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/i8_plain.log; tail -3 gpurun_out/ncu_i8.log
