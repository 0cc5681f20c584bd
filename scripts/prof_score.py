"""Config B scoring (65536 sets of 8 vs the N=4096 factor) a few times; plain or under ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from algp_b200 import engine  # noqa: E402

grid, y, base, idx, delta, hy = bench.workload()
hyper = engine.Hyper(np.log(hy["ls"]), np.log(hy["os"]), np.log(hy["noise"]), hy["kind"])
pi0 = np.zeros(len(grid))
pi0[base] = 1.0 / bench.STATIC_STD ** 2
state = engine.PosteriorState(hyper, engine.to_dev(grid), base, pi0, is_static=pi0 > 0)
H = state.H_base
idx_d, delta_d = engine.to_dev(idx, dtype=torch.int32), engine.to_dev(delta)
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s = state.score_sets(idx_d, delta_d, H_base=H)
    e1.record()
    p = state.argmax(s)
    torch.cuda.synchronize()
    print("rep %d: score %.3f ms, best %d" % (rep, e0.elapsed_time(e1), int(p[1].item())))
