#!/bin/bash
# ncu evidence at the benchmark shapes (run on the GPU box through scripts/gpu.sh; every program first runs plain).
# Reports land in gpurun_out/; scripts/ncu_summary.py turns them into the CSVs under profiles/.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. the headline scoring kernel, configs[2], single launch
python scripts/prof_score.py --mode stream --reps 3 > $O/ev_score_plain.log 2>&1 &&
  $NCU -k regex:score_sets_k8 -s 2 -c 1 -o $O/r02_prof_score python scripts/prof_score.py --mode stream --reps 3 > $O/ev_score_ncu.log 2>&1
# 2. kernel build (16384^2 symmetric, 65536 x 16384 with fused mean) and the two triangular gemv solves, N = 16384
python scripts/prof_fit.py --n 16384 --side 256 --reps 1 --i8 --zorder > $O/ev_fit_plain.log 2>&1 &&
  $NCU -k regex:"kbuild|gemv_lower" -c 8 -o $O/r02_prof_kbuild_solve python scripts/prof_fit.py --n 16384 --side 256 --reps 1 --i8 --zorder > $O/ev_fit_ncu.log 2>&1
# 3. MLL gradient pass, N = 4096
python scripts/prof_mll.py > $O/ev_mll_plain.log 2>&1 &&
  $NCU -k regex:mll_grad -c 1 -o $O/r02_prof_mll python scripts/prof_mll.py > $O/ev_mll_ncu.log 2>&1
# 4. per-launch list of a short bench run (kernel shares of the step)
python bench.py --steps 2 --warmup 1 --skip-large --no-cpu > $O/ev_bench_plain.json 2> $O/ev_bench_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-large --no-cpu > $O/ev_bench_ncu.json 2> $O/ev_bench_ncu.err
ls -la $O/*.ncu-rep $O/r02_launches_bench.csv
