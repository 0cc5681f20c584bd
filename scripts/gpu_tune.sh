#!/bin/bash
# compare score-kernel tuning variants and measure pure-store / pure-read HBM bandwidth
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do
  echo "variant $v: $(ALGP_SCORE_VARIANT=$v python scripts/prof_score.py 2>&1 | tail -2 | tr '\n' ' ')"
done | tee gpurun_out/tune_score.log
python - <<'PY' | tee gpurun_out/hbm_store.log
import torch
n = 1 << 28                        # 2 GiB of fp64
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.fill_(1.5));   print("store-only fill 2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: x.zero_());      print("store-only memset 2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: y.copy_(x));     print("copy 2 GiB (r+w bytes): %.3f ms  %.0f GB/s" % (ms, 2 * n * 8 / ms / 1e6))
ms = t(lambda: x.sum());        print("read-only sum 2 GiB: %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
PY
