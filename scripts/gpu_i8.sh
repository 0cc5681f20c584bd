#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "i8" --timeout 30 -p no:cacheprovider > gpurun_out/pytest_i8.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_i8.log
timeout 120 python scripts/prof_i8.py 4096 4096 7,8 > gpurun_out/i8_4096.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/i8_4096.log
