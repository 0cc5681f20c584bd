"""Run the fit+predict pipeline (kbuild -> potrf -> trtri -> solve -> cross/mean -> variance) a few times.
Used plain (CUDA-event timings) and under ncu for the per-launch list."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algp_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--side", type=int, default=64)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tf32", action="store_true")
ap.add_argument("--i8", action="store_true")
ap.add_argument("--rank", type=int, default=2)
ap.add_argument("--zorder", action="store_true", help="sort train / test points along a Z curve as precision i8 does")
a = ap.parse_args()
from algp_b200._lib import call as _call
_call("algp_set_potf2_rank", a.rank)
rng = np.random.default_rng(1)
x = rng.uniform(0, a.side, size=(a.n, 2))
yy, xx = np.meshgrid(np.arange(a.side), np.arange(a.side), indexing="ij")
xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
y = np.sin(x[:, 0] / 9.0) + rng.normal(0, 0.1, a.n)
hy = engine.Hyper(np.log([a.side / 16.0] * 2), 0.0, np.log(1e-2), "rbf")
xd, xsd, y0 = engine.to_dev(x), engine.to_dev(xs), engine.to_dev(y - y.mean())
var = engine.to_dev(np.full(a.n, 0.01))
if a.zorder:
    perm, lo, hi = engine.morton_perm(xd)
    xd, y0 = xd[perm].contiguous(), y0[perm].contiguous()
    xsd = xsd[engine.morton_perm(xsd, lo, hi)[0]].contiguous()
for rep in range(a.reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    f = engine.GPFactor(hy, xd, diag_add=var, factor="auto" if a.i8 else "dmma")
    e[1].record()
    mu, v = f.mean_var(xsd, y0, float(y.mean()), precision="tf32" if a.tf32 else ("i8" if a.i8 else "fp64"))
    e[2].record()
    torch.cuda.synchronize()
    f.check()
    print("rep %d: factor(kbuild+potrf+trtri) %.3f ms, predict %.3f ms, var range %.3e..%.3e" %
          (rep, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), v.min().item(), v.max().item()))
