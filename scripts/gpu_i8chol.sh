#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "i8" -p no:cacheprovider > gpurun_out/pytest_i8.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_i8.log
timeout 300 python scripts/prof_i8chol.py 8192 1024,2048 8 > gpurun_out/i8chol_8192.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/i8chol_8192.log
timeout 300 python scripts/prof_i8chol.py 16384 1024,2048,4096 7,8 > gpurun_out/i8chol_16384.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/i8chol_16384.log
