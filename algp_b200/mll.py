"""Exact marginal log-likelihood and its gradient on the device (reference models.py:145-158).

loss = -(1/N) log N(y0; 0, A),  A = s^2 k(X,X) + diag(var) + sigma_n^2 I   (models.py:148; gpytorch's
ExactMarginalLogLikelihood divides by N), differentiated w.r.t. the raw log-parameters
(log lengthscale[d], log outputscale, log noise) that Adam updates in GPR.fit.
"""
import numpy as np
import torch

from . import _lib, engine
from ._lib import call, ptr, stream


def mll_loss_and_grad(hyper, x, y0, var):
    """Returns (loss, grad[d+2]) as host float / ndarray; one small D2H per call."""
    dev = engine.require_cuda()
    xd = engine.to_dev(np.asarray(x, dtype=np.float64), device=dev)
    yd = engine.to_dev(np.asarray(y0, dtype=np.float64), device=dev)
    vd = engine.to_dev(np.asarray(var, dtype=np.float64), device=dev)
    n, d = xd.shape
    f = engine.GPFactor(hyper, xd, diag_add=vd, diag_scalar=hyper.noise)
    alpha, beta = f.solve(yd)
    ldq = f.logdet_quad(beta)
    Ainv = torch.empty((f.Npad, f.Npad), dtype=torch.float64, device=dev)
    call("algp_potri_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(Ainv), f.Npad, stream())
    work = torch.empty(_lib.lib.algp_mll_grad_work_doubles(n), dtype=torch.float64, device=dev)
    out = torch.empty(d + 2 + 2, dtype=torch.float64, device=dev)
    ls, ls_p = _lib.host_f64(hyper.log_ls)
    call("algp_mll_grad", ptr(xd), n, d, ls_p, hyper.log_os, hyper.kind, hyper.noise, ptr(alpha), ptr(Ainv), f.Npad,
         ptr(work), ptr(out), stream())
    out[d + 2:] = ldq
    host = out.cpu().numpy()
    f.check()
    logdet, quad = host[d + 2], host[d + 3]
    ll = -0.5 * quad - 0.5 * logdet - 0.5 * n * np.log(2 * np.pi)
    return float(-ll / n), -host[:d + 2] / n
