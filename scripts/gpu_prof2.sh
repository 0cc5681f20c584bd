#!/bin/bash
# ncu evidence, one capture per kernel family (each only after the same command exited 0 without ncu)
mkdir -p gpurun_out
P="python scripts/prof_fit.py --n 4096 --side 128 --reps 1"
$P > gpurun_out/p_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 70 -c 3 -o gpurun_out/prof_gemm $P > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
$P > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potf2inv -s 4 -c 1 -o gpurun_out/prof_potf2 $P > gpurun_out/ncu_potf2.log 2>&1
echo "potf2 rc=$?"
$P > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kbuild -c 2 -o gpurun_out/prof_kbuild2 $P > gpurun_out/ncu_kbuild2.log 2>&1
echo "kbuild rc=$?"
$P --tf32 > gpurun_out/p_tf32_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"trmm_tf32x3|split_tf32" -c 3 -o gpurun_out/prof_tf32 $P --tf32 > gpurun_out/ncu_tf32.log 2>&1
echo "tf32 rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu --skip-large > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-large > gpurun_out/ncu_bench.log 2>&1
echo "bench launches rc=$?"
cat gpurun_out/p_plain.log gpurun_out/p_tf32_plain.log
