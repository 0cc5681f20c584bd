#!/bin/bash
# what the driver runs at round end (tests, smoke, both bench arms), then the ncu evidence for this round's new kernels:
# launch list of the bench command and one --set full capture of score_cov_k8_kernel (plain runs first)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -4 gpurun_out/pytest_all.log; tail -1 gpurun_out/smoke.log; tail -4 gpurun_out/bench_default.err; tail -4 gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.log
python bench.py --steps 3 --warmup 3 --no-cpu --skip-large > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-large > gpurun_out/ncu_bench.log 2>&1
echo "bench launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_cov_k8 -s 2 -c 1 -f -o gpurun_out/prof_score_cov python bench.py --steps 3 --warmup 3 --no-cpu --skip-large > gpurun_out/ncu_cov.log 2>&1
echo "ncu cov rc=$?"; tail -2 gpurun_out/ncu_cov.log
