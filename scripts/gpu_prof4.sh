#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_trmm.py > gpurun_out/trmm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 1 -c 1 -o gpurun_out/prof_trmm python scripts/prof_trmm.py > gpurun_out/ncu_trmm.log 2>&1
echo "rc=$?"; cat gpurun_out/trmm_plain.log
