"""Where does the ~10.3 TB/s ceiling of the candidate-scoring kernels come from?  Config B (65536 sets of 8, N=4096)
with (a) the row pitch of Wt moved off the power of two, (b) the candidate rows confined to windows of the field that
fit the TLB reach / the L2, for the row-streaming and the L2-tiled kernel."""
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
Xd = engine.to_dev(grid)
delta_d = engine.to_dev(delta)


def run(state, idx_d, mode, tile=0, reps=5):
    state.score_mode = mode
    _lib.lib.algp_set_score_tile_cols(tile)
    H = state.H_base
    out = torch.empty(idx_d.shape[0], dtype=torch.float64, device=idx_d.device)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        state.score_sets(idx_d, delta_d, H_base=H, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    _lib.lib.algp_set_score_tile_cols(0)
    return min(ts)


for cap in (0, 32, 96, 160):
    st = engine.PosteriorState(hyper, Xd, base, pi0, is_static=pi0 > 0, capacity=cap, cov_mode="never")
    idx_d = engine.to_dev(idx, dtype=torch.int32)
    print("row pitch %5d doubles (%6d B): stream %.3f ms   tiled(1024) %.3f ms   tiled(512) %.3f ms" %
          (st.ldw, st.ldw * 8, run(st, idx_d, "stream"), run(st, idx_d, "tiled", 1024), run(st, idx_d, "tiled", 512)))
    if cap == 32:
        for rows in (16384, 12000, 8000, 6000, 4000, 3000, 2000, 1000):
            ix = engine.to_dev((idx % rows).astype(np.int32), dtype=torch.int32)
            t_s, t_t = run(st, ix, "stream"), run(st, ix, "tiled", 1024)
            print("   candidate rows confined to the first %5d (%4d MB of Wt): stream %.3f ms (%.1f TB/s)   tiled(1024) %.3f ms (%.1f TB/s)" %
                  (rows, rows * st.ldw * 8 >> 20, t_s, 17.18 / t_s, t_t, 17.18 / t_t))
    del st
