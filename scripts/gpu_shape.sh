#!/bin/bash
mkdir -p gpurun_out
for sh in 0 1; do
  echo "== ALGP_GEMM_SHAPE=$sh"
  ALGP_GEMM_SHAPE=$sh timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "potrf or gemm or whiten or solve" 2>&1 | tail -2
  ALGP_GEMM_SHAPE=$sh python scripts/prof_fit.py --n 4096 --reps 3 2>&1 | tail -1
  ALGP_GEMM_SHAPE=$sh python scripts/prof_fit.py --n 16384 --side 256 --reps 3 2>&1 | tail -1
done | tee gpurun_out/gemm_shape.log
