#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "i8" -p no:cacheprovider > gpurun_out/pytest_i8.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_i8.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-large --no-cpu > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_small.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['setup_ms_factor_and_W'], d['setup_ms_factor_and_W_i8'], d['setup_i8_max_abs_dW'])
PY
tail -3 gpurun_out/bench_small.err
