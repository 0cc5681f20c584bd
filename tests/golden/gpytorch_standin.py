"""Stand-in for the 2018 "Beta" gpytorch API that /root/reference/models.py
imports (README.md:9; un-pinned, absent from this image, incompatible with any
modern gpytorch).  TEST INFRASTRUCTURE used only by make_golden.py so that the
reference's own models.py / utils.py / agent.py can execute in this container.

It restates, densely and in the caller's dtype (float32, because the reference
feeds FloatTensors), exactly the pieces models.py touches:
ScaleKernel(RBFKernel | MaternKernel(nu=1.5), ard_num_dims=d), WhiteNoiseKernel,
ZeroMean, GaussianLikelihood (log_noise), MultivariateNormal with the old
callable ``mean()`` / ``covar()``, ExactGP (train: prior at the training inputs,
eval: exact posterior) and ExactMarginalLogLikelihood (log-prob / N).
Parameters are raw logs initialised at 0 like the reference-era library
(SURVEY.md 9.1).  The kernel closed forms are the published ones; this is the
"parity unpinned" part of the oracle (see oracle/oracle.py header).
"""
import math
import sys
import types

import torch
import torch.nn as nn


class _Lazy(object):
    def __init__(self, t):
        self._t = t

    def evaluate(self):
        return self._t

    def diag(self):
        return torch.diagonal(self._t)

    def cpu(self):
        return self

    def __add__(self, other):
        return _Lazy(self._t + (other._t if isinstance(other, _Lazy) else other))


class Kernel(nn.Module):
    def __call__(self, x1, x2=None):
        x2_ = x1 if x2 is None else x2
        if x1.dim() == 1:
            x1 = x1.unsqueeze(-1)
        if x2_.dim() == 1:
            x2_ = x2_.unsqueeze(-1)
        return _Lazy(self.forward(x1, x2_))

    def __add__(self, other):
        return AdditiveKernel(self, other)


class AdditiveKernel(Kernel):
    def __init__(self, k1, k2):
        super().__init__()
        self.k1 = k1
        self.k2 = k2

    def forward(self, x1, x2):
        return self.k1.forward(x1, x2) + self.k2.forward(x1, x2)


class _ARDKernel(Kernel):
    def __init__(self, ard_num_dims=None, **kw):
        super().__init__()
        d = 1 if ard_num_dims is None else ard_num_dims
        self.log_lengthscale = nn.Parameter(torch.zeros(1, 1, d))

    def _sqdist(self, x1, x2):
        ls = self.log_lengthscale.exp().view(1, -1)
        a = x1 / ls
        b = x2 / ls
        diff = a.unsqueeze(1) - b.unsqueeze(0)
        return (diff * diff).sum(-1)


class RBFKernel(_ARDKernel):
    def forward(self, x1, x2):
        return torch.exp(-0.5 * self._sqdist(x1, x2))


class MaternKernel(_ARDKernel):
    def __init__(self, nu=2.5, ard_num_dims=None, **kw):
        super().__init__(ard_num_dims=ard_num_dims)
        if nu != 1.5:
            raise NotImplementedError("stand-in covers nu=1.5 (models.py:222)")
        self.nu = nu

    def forward(self, x1, x2):
        r2 = self._sqdist(x1, x2)
        # d sqrt / d 0 is inf; fit() differentiates through here, so the r = 0
        # entries (the diagonal) are masked BEFORE the sqrt and get zero gradient
        pos = r2 > 0
        r = torch.where(pos, torch.sqrt(torch.where(pos, r2, torch.ones_like(r2))), torch.zeros_like(r2))
        s3 = math.sqrt(3.0)
        return (1.0 + s3 * r) * torch.exp(-s3 * r)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, **kw):
        super().__init__()
        self.base_kernel = base_kernel
        self.log_outputscale = nn.Parameter(torch.zeros(1))

    def forward(self, x1, x2):
        return self.log_outputscale.exp() * self.base_kernel.forward(x1, x2)


class WhiteNoiseKernel(Kernel):
    def __init__(self, variances):
        super().__init__()
        self.register_buffer("variances", variances.clone().view(-1))

    def forward(self, x1, x2):
        if x1.size(0) == x2.size(0) == self.variances.numel() and torch.equal(x1, x2):
            return torch.diag(self.variances.to(x1.dtype))
        return torch.zeros(x1.size(0), x2.size(0), dtype=x1.dtype)


class SpectralMixtureKernel(Kernel):
    def __init__(self, *a, **kw):
        raise NotImplementedError("spectral mixture is out of scope (SURVEY.md 2)")


class ZeroMean(nn.Module):
    def forward(self, x):
        return torch.zeros(x.size(0), dtype=x.dtype)


class MultivariateNormal(object):
    def __init__(self, mean, covar):
        self._mean = mean
        self._covar = covar if isinstance(covar, _Lazy) else _Lazy(covar)

    def mean(self):
        return self._mean

    def covar(self):
        return self._covar

    def log_prob(self, target):
        A = self._covar.evaluate()
        L = torch.linalg.cholesky(A)
        diff = (target - self._mean).unsqueeze(-1)
        beta = torch.linalg.solve_triangular(L, diff, upper=False)
        n = target.numel()
        return -0.5 * (beta * beta).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2 * math.pi)


class GaussianLikelihood(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.log_noise = nn.Parameter(torch.zeros(1, 1))

    def forward(self, dist):
        cov = dist.covar().evaluate()
        n = cov.size(0)
        return MultivariateNormal(dist.mean(), cov + self.log_noise.exp().view(()) * torch.eye(n, dtype=cov.dtype))


class ExactGP(nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            self.train_inputs = (inputs,)
        if targets is not None:
            self.train_targets = targets

    def __call__(self, x):
        if self.training:
            return self.forward(x)
        # exact posterior at x given (train_inputs, train_targets): SURVEY.md 9.2
        xt = self.train_inputs[0]
        prior = self.forward(xt)
        A = self.likelihood(prior).covar().evaluate()
        xl = self.latent_func(x)
        xtl = self.latent_func(xt)
        Ksx = self.kernel_covar_module(xl, xtl).evaluate()
        Kss = self.kernel_covar_module(xl).evaluate()
        L = torch.linalg.cholesky(A)
        V = torch.linalg.solve_triangular(L, Ksx.t(), upper=False)
        beta = torch.linalg.solve_triangular(L, self.train_targets.unsqueeze(-1), upper=False)
        mean = (V.t() @ beta).squeeze(-1)
        return MultivariateNormal(mean, Kss - V.t() @ V)


class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model

    def forward(self, output, target):
        return self.likelihood(output).log_prob(target) / target.numel()


def install():
    """Register the stand-in as ``gpytorch`` (+ the sub-modules models.py
    imports) and inert stubs for the plotting / debugging packages that
    utils.py / agent.py import at module scope but the hot path never calls."""
    from unittest import mock

    g = types.ModuleType("gpytorch")
    subs = {
        "kernels": dict(RBFKernel=RBFKernel, WhiteNoiseKernel=WhiteNoiseKernel, MaternKernel=MaternKernel,
                        SpectralMixtureKernel=SpectralMixtureKernel, ScaleKernel=ScaleKernel),
        "means": dict(ZeroMean=ZeroMean),
        "likelihoods": dict(GaussianLikelihood=GaussianLikelihood),
        "distributions": dict(MultivariateNormal=MultivariateNormal),
        "models": dict(ExactGP=ExactGP),
        "mlls": dict(ExactMarginalLogLikelihood=ExactMarginalLogLikelihood),
    }
    sys.modules["gpytorch"] = g
    for name, members in subs.items():
        m = types.ModuleType("gpytorch." + name)
        for k, v in members.items():
            setattr(m, k, v)
        setattr(g, name, m)
        sys.modules["gpytorch." + name] = m
    for missing in ("ipdb", "seaborn", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(missing)
        except Exception:
            sys.modules[missing] = mock.MagicMock(name=missing)
    # torch >= 2.7 dropped ReduceLROnPlateau(verbose=...) (models.py:124)
    import torch.optim.lr_scheduler as lrs
    _orig = lrs.ReduceLROnPlateau

    class _ReduceLROnPlateau(_orig):
        def __init__(self, *a, verbose=None, **kw):
            super().__init__(*a, **kw)

    lrs.ReduceLROnPlateau = _ReduceLROnPlateau
