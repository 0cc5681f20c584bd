"""Scoring of large batches, configs[2]: the single launch, one launch per column chunk (accumulator fragments parked in
global memory), and the persistent launch with the partial Grams resident in shared memory (csrc/score.cu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from algp_b200 import _lib, engine  # noqa: E402

grid, y, base, idx, delta, hy = bench.workload()
hyper = engine.Hyper(np.log(hy["ls"]), np.log(hy["os"]), np.log(hy["noise"]), hy["kind"])
pi0 = np.zeros(len(grid))
pi0[base] = 1.0 / bench.STATIC_STD ** 2
state = engine.PosteriorState(hyper, engine.to_dev(grid), base, pi0, is_static=pi0 > 0, cov_mode="never")
H = state.H_base
idx_d, delta_d = engine.to_dev(idx, dtype=torch.int32), engine.to_dev(delta)
pair = torch.empty(2, dtype=torch.int64, device="cuda")
state.score_mode = "tiled"


def run(B, tile, resident=0, reps=12):
    _lib.lib.algp_set_score_tile_cols(tile)
    _lib.lib.algp_set_score_resident(resident)
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(3):
        state.score_sets(idx_d[:B], delta_d[:B], H_base=H, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        state.score_sets(idx_d[:B], delta_d[:B], H_base=H, out=out)
        state.argmax(out, 0, out=pair)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out.clone()


CHUNKS = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [512, 640, 768, 896, 1024]
for B in [int(b) for b in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["65536", "65536", "32768"])]:
    t0, ref = run(B, -1)
    t1, s1 = run(B, 1024, -1)
    t2, s2 = run(B, 0, 0)
    line = "B=%6d single launch %.4f ms | launches/1024 %.4f (%.0e) | default policy %.4f (%.0e)" % (
        B, t0, t1, float((s1 - ref).abs().max()), t2, float((s2 - ref).abs().max()))
    for ch in CHUNKS:
        t, s_ = run(B, ch, 1)
        line += " | resident/%d %.4f (%.0e)" % (ch, t, float((s_ - ref).abs().max()))
    print(line, flush=True)
_lib.lib.algp_set_score_tile_cols(0)
_lib.lib.algp_set_score_resident(0)
