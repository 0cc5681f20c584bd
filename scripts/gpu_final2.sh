#!/bin/bash
# round-end check without the ncu passes: tests, smoke, both bench arms
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
grep "passed\|failed" gpurun_out/pytest_all.log; tail -1 gpurun_out/smoke.log; tail -3 gpurun_out/bench_default.err; cut -c1-200 gpurun_out/bench_ref.log
