#!/bin/bash
# One GPU-box visit: parity tests (two processes so a CUDA fault in one file does not mask the other), then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/rc.log
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_api.log 2>&1
echo "api rc=$?" | tee -a gpurun_out/rc.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/rc.log
timeout 900 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/rc.log
tail -5 gpurun_out/pytest_kernels.log; tail -5 gpurun_out/pytest_api.log; tail -2 gpurun_out/smoke.log; tail -c 600 gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log
