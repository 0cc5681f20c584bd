#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "potf2_rank or potrf" -p no:cacheprovider > gpurun_out/pytest_potf2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_potf2.log
python scripts/prof_potf2.py
for r in 1 2 4; do echo "rank $r"; timeout 120 python scripts/prof_fit.py --n 4096 --side 64 --reps 4 --rank $r | tail -2; done
