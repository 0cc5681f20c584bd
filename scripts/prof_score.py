"""Config B scoring (65536 sets of 8 vs the N=4096 factor): the row-streaming kernel against the L2-tiled one at
several chunk widths; plain or under ncu (`--mode tiled --tile 512 --reps 2` for one variant only)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from algp_b200 import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="sweep", choices=["sweep", "stream", "tiled"])
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--reps", type=int, default=6)
args = ap.parse_args()

grid, y, base, idx, delta, hy = bench.workload()
hyper = engine.Hyper(np.log(hy["ls"]), np.log(hy["os"]), np.log(hy["noise"]), hy["kind"])
pi0 = np.zeros(len(grid))
pi0[base] = 1.0 / bench.STATIC_STD ** 2
state = engine.PosteriorState(hyper, engine.to_dev(grid), base, pi0, is_static=pi0 > 0, cov_mode="never")
H = state.H_base
idx_d, delta_d = engine.to_dev(idx, dtype=torch.int32), engine.to_dev(delta)
out = torch.empty(len(idx), dtype=torch.float64, device=idx_d.device)


def run(mode, tile, reps):
    state.score_mode = mode
    _lib.lib.algp_set_score_tile_cols(tile)
    ts = []
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = state.score_sets(idx_d, delta_d, H_base=H, out=out)
        e1.record()
        p = state.argmax(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts, int(p[1].item()), s.clone()


if args.mode == "sweep":
    ts, best, ref = run("stream", 0, args.reps)
    print("stream          : min %.3f ms median %.3f ms  best %d" % (min(ts), float(np.median(ts)), best))
    for tile in (0, 256, 512, 768, 1024, 1280, 1536, 2048, 4096):
        ts, b, s = run("tiled", tile, args.reps)
        print("tiled chunk %4d: min %.3f ms median %.3f ms  best %d  max|dscore| %.2e" %
              (tile, min(ts), float(np.median(ts)), b, float((s - ref).abs().max().item())))
else:
    ts, best, _ = run(args.mode, args.tile, args.reps)
    print("%s tile %d: %s best %d" % (args.mode, args.tile, " ".join("%.3f" % t for t in ts), best))
_lib.lib.algp_set_score_tile_cols(0)
