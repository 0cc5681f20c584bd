#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -x -q -m gpu -k "append_block or episode" -p no:cacheprovider > gpurun_out/pytest_blk.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_blk.log
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
import torch, numpy as np
import bench
from algp_b200 import engine
print(json.dumps(bench.episode_bench(torch, engine)))
PY
