#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_api.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_long.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_long.log
