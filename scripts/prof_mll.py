"""One evaluation of the exact MLL and its gradient at N = 4096 (d = 2, RBF): target for ncu on mll_grad_kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algp_b200 import engine  # noqa: E402
from algp_b200.mll import MLLWorkspace  # noqa: E402

rng = np.random.default_rng(7)
n = 4096
x = rng.uniform(0, 64, size=(n, 2))
y = np.sin(x[:, 0] / 5.0) + np.cos(x[:, 1] / 7.0) + rng.normal(0, 0.1, n)
ws = MLLWorkspace(x, y - y.mean(), np.full(n, 0.01))
hy = engine.Hyper(np.log([4.0, 4.0]), 0.0, np.log(1e-2), "rbf")
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss, g = ws.loss_and_grad(hy)
    e1.record()
    torch.cuda.synchronize()
    print("rep %d: loss %.6f  |grad| %.3e  %.3f ms" % (rep, loss, float(np.abs(g).max()), e0.elapsed_time(e1)))
