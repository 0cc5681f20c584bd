import numpy as np
import torch

import oracle as O
from algp_b200 import engine


def hyper_pair(ls, os_, noise, kind):
    th = O.Theta.from_values(ls, os_, noise, kind)
    hy = engine.Hyper(th.log_lengthscale, th.log_outputscale, th.log_noise, kind)
    return th, hy


def dev(a, dtype=torch.float64):
    return engine.to_dev(np.asarray(a), dtype=dtype)


def field_problem(n_side_r, n_side_c, n_train, seed, d_extra=0):
    rng = np.random.default_rng(seed)
    grid, y = O.gaussian_mixture_field(n_side_r, n_side_c, seed=seed)
    X = grid
    if d_extra:
        X = np.hstack([grid, rng.integers(0, 2, size=(len(grid), d_extra)).astype(np.float64)])
    tr = rng.choice(len(X), n_train, replace=False)
    ytr = np.maximum(0, y[tr] + rng.normal(0, 0.1, n_train))
    return X, y, tr, ytr, rng
