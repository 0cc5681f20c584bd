#!/bin/bash
# full GPU test suite + smoke, then the ncu evidence of the INT8 digit pipeline (launch list of the factorisation,
# one --set full capture of the masked variance GEMM); plain runs first, as the profiling recipe requires
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/pytest_all.log; tail -1 gpurun_out/smoke.log
timeout 200 python scripts/prof_i8chol.py 16384 2048 8 morton > gpurun_out/i8chol_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit16384_i8.csv \
  python scripts/prof_i8chol.py 16384 2048 8 morton > gpurun_out/ncu_i8chol.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/i8chol_plain.log
CMD="python scripts/prof_i8.py 8192 16384 7 nocheck morton"
timeout 120 $CMD > gpurun_out/i8_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_i8 -s 1 -c 1 -f -o gpurun_out/prof_i8_uniform $CMD > gpurun_out/ncu_i8.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/i8_plain.log; tail -3 gpurun_out/ncu_i8.log
