"""Scoring-kernel variants against the batch size (strong scaling shards the 65536 sets of configs[2] into 32768 /
16384 / 8192 per GPU): single launch, one launch per column chunk, and the item mode (one launch over
(chunk, candidate) work items + an epilogue-only launch)."""
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


def run(B, mode, tile=0, env=None, reps=12):
    state.score_mode = mode
    _lib.lib.algp_set_score_tile_cols(tile)
    if env:
        os.environ["ALGP_SCORE_TILE_MODE"] = env
    else:
        os.environ.pop("ALGP_SCORE_TILE_MODE", None)
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(3):
        state.score_sets(idx_d[:B], delta_d[:B], H_base=H, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):                       # back to back, as in the bench step (score + argmax)
        state.score_sets(idx_d[:B], delta_d[:B], H_base=H, out=out)
        state.argmax(out, 0, out=pair)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out.clone()


for B in (256, 1000, 4096, 8192, 10000, 12000, 16384, 32768, 65536):
    os.environ.pop("ALGP_SCORE_PARTS", None)
    t0, ref = run(B, "tiled", -1)
    line = "B=%5d  one warp per candidate %.4f ms" % (B, t0)
    for parts in (2, 4):
        if B > 16384:
            continue
        os.environ["ALGP_SCORE_PARTS"] = str(parts)
        t, s_ = run(B, "tiled", 0)
        line += " | %d warps per candidate %.4f (%.0e)" % (parts, t, float((s_ - ref).abs().max()))
    os.environ.pop("ALGP_SCORE_PARTS", None)
    t, s_ = run(B, "tiled", 0)
    line += " | default policy %.4f (%.0e)" % (t, float((s_ - ref).abs().max()))
    print(line)
_lib.lib.algp_set_score_tile_cols(0)
