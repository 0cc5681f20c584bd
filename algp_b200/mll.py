"""Exact marginal log-likelihood and its gradient on the device (reference models.py:145-158).

loss = -(1/N) log N(y0; 0, A),  A = s^2 k(X,X) + diag(var) + sigma_n^2 I   (models.py:148; gpytorch's
ExactMarginalLogLikelihood divides by N), differentiated w.r.t. the raw log-parameters
(log lengthscale[d], log outputscale, log noise) that Adam updates in GPR.fit.

``MLLWorkspace`` keeps the training data and every N^2 work buffer resident in HBM across the iterations of the
Adam loop (models.py:145-158 evaluates the same-sized problem max_iter times): an iteration uploads nothing but the
d + 2 hyper-parameters (as kernel arguments) and reads back d + 4 doubles.
"""
import numpy as np
import torch

from . import _lib, engine
from ._lib import call, ptr, stream


class MLLWorkspace(object):
    def __init__(self, x, y0, var):
        dev = engine.require_cuda()
        self.x = engine.to_dev(np.asarray(x, dtype=np.float64), device=dev)
        self.y0 = engine.to_dev(np.asarray(y0, dtype=np.float64), device=dev)
        self.var = engine.to_dev(np.asarray(var, dtype=np.float64), device=dev)
        self.n, self.d = self.x.shape
        npad = max(engine.BLK, engine.pad_to(self.n))
        f64 = dict(dtype=torch.float64, device=dev)
        self.buffers = {"L": torch.empty((npad, npad), **f64), "Linv": torch.empty((npad, npad), **f64),
                        "info": torch.zeros(1, dtype=torch.int32, device=dev),
                        "trtri": torch.empty(max(2, _lib.lib.algp_trtri_work_doubles(npad)), **f64)}
        self.Ainv = torch.empty((npad, npad), **f64)
        self.work = torch.empty(_lib.lib.algp_mll_grad_work_doubles(self.n), **f64)
        self.out = torch.empty(self.d + 4, **f64)
        self.host = torch.empty(self.d + 5, dtype=torch.float64).pin_memory()

    def loss_and_grad(self, hyper):
        """(loss, grad[d+2]) as host float / ndarray at the hyper-parameters `hyper`; one (d+5)-double D2H."""
        n, d = self.n, self.d
        f = engine.GPFactor(hyper, self.x, diag_add=self.var, diag_scalar=hyper.noise, buffers=self.buffers)
        alpha, beta = f.solve(self.y0)
        ldq = f.logdet_quad(beta)
        call("algp_potri_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(self.Ainv), f.Npad, stream())
        ls, ls_p = _lib.host_f64(hyper.log_ls)
        call("algp_mll_grad", ptr(self.x), n, d, ls_p, hyper.log_os, hyper.kind, hyper.noise, ptr(alpha), ptr(self.Ainv), f.Npad,
             ptr(self.work), ptr(self.out), stream())
        self.out[d + 2:] = ldq
        self.host[:d + 4].copy_(self.out, non_blocking=True)
        self.host[d + 4:].copy_(f.info.to(torch.float64), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = self.host.numpy()
        if int(host[d + 4]) != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite (leading minor of order %d)" % int(host[d + 4]))
        logdet, quad = host[d + 2], host[d + 3]
        ll = -0.5 * quad - 0.5 * logdet - 0.5 * n * np.log(2 * np.pi)
        return float(-ll / n), -host[:d + 2].copy() / n


def mll_loss_and_grad(hyper, x, y0, var):
    """One evaluation with a throw-away workspace (tests, GPR.loss_and_grad)."""
    return MLLWorkspace(x, y0, var).loss_and_grad(hyper)
