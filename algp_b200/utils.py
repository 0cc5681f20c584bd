"""Drop-in for the hot-path functions of the reference's ``utils.py``:
``CONST``, ``to_torch``, ``to_numpy`` (utils.py:10-30), ``entropy_from_cov``
(utils.py:188-194) and ``predictive_distribution`` (utils.py:293-319), plus a
seeded ``generate_gaussian_data`` (utils.py:90-108) for the benchmarks.

Same signatures and flag-dependent return tuples; inputs and outputs are host
NumPy arrays; the arithmetic runs on the GPU in fp64 (Cholesky instead of the
reference's explicit float32 inverse / LU slogdet).
"""
import hashlib

import numpy as np
import torch

CONST = .5 * np.log(2 * np.pi * np.exp(1))


def to_torch(arr):
    # utils.py:13-20 forces float32; this build keeps float64 (the fp64 tier of the north star)
    if arr is None:
        return None
    if arr.__class__.__module__ == 'torch':
        return arr
    if arr.__class__.__module__ == 'numpy':
        return torch.from_numpy(np.asarray(arr, dtype=np.float64))
    return arr


def to_numpy(x):
    if x is None:
        return None
    if x.__class__.__module__ == 'torch':
        return x.detach().cpu().numpy()
    if x.__class__.__module__ == 'numpy':
        return x
    return np.array(x)


def generate_gaussian_data(num_rows, num_cols, k=5, min_var=10, max_var=100, algo='sum', seed=None):
    """Mixture-of-Gaussians field of utils.py:90-108; ``seed`` (not in the reference, which never
    seeds) selects numpy.random.default_rng(seed), None keeps the global RNG like the reference."""
    x, y = np.meshgrid(np.arange(num_cols), np.arange(num_rows))
    grid = np.vstack([y.flatten(), x.flatten()]).transpose()
    rng = np.random if seed is None else np.random.default_rng(seed)
    means_x = rng.uniform(0, num_rows, size=k)
    means_y = rng.uniform(0, num_cols, size=k)
    means = np.vstack([means_x, means_y]).transpose()
    variances = rng.uniform(min_var, max_var, size=k)
    y = np.zeros(num_rows * num_cols)
    for i in range(k):
        dist_sq = np.sum(np.square(grid - means[i].reshape(1, -1)), axis=1)
        tmp = np.exp(-dist_sq / variances[i])
        if algo == 'max':
            y = np.maximum(y, tmp)
        elif algo == 'sum':
            y += tmp
    return grid, y


def entropy_from_cov(cov, constant=CONST):
    """utils.py:188-194: k*CONST + 0.5*log det(cov), log det from a device Cholesky.
    Raises numpy.linalg.LinAlgError if cov is not positive definite (the reference's LU slogdet
    would return log|det| of an indefinite matrix; a covariance is never indefinite)."""
    from . import engine
    if constant is None:
        constant = CONST
    cov = np.asarray(to_numpy(cov), dtype=np.float64)
    k = cov.shape[0]
    if k == 0:
        return 0.0
    dev = engine.require_cuda()
    kpad = max(engine.BLK, engine.pad_to(k))
    A = torch.eye(kpad, dtype=torch.float64, device=dev)
    A[:k, :k] = engine.to_dev(cov, device=dev)
    return k * constant + 0.5 * engine.chol_logdet(A, k)


def _digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"none")
        else:
            a = np.ascontiguousarray(a)
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
    return h.digest()


def _factor_for(gp, hyper, train_x, train_var, reorder=False):
    """Device factor of cov_aa (utils.py:296), cached on the GPR until its data or theta change.
    reorder: factor the training set in Z-curve order (f.perm maps factor rows to the caller's rows); only
    order-independent consumers (posterior mean / variance) may ask for it."""
    from . import engine
    cache = getattr(gp, "_cache", None)
    prec = getattr(gp, "precision", "fp64")
    # precision "i8": factor through the recursive INT8 digit factorisation when N is large enough to pay
    key = ("factor", _digest(train_x, train_var), hyper.key(), engine.factor_plan(prec, np.shape(train_x)[0]), bool(reorder))
    if cache is not None and cache.get("factor_key") == key:
        return cache["factor"]
    dev = engine.require_cuda()
    x = engine.to_dev(train_x, device=dev)
    wn = None if train_var is None else engine.to_dev(np.asarray(train_var, dtype=np.float64), device=dev)
    perm = box = None
    if reorder:
        perm, lo, hi = engine.morton_perm(x)
        box = (lo, hi)
        x = x.index_select(0, perm).contiguous()
        wn = None if wn is None else wn.index_select(0, perm).contiguous()
    fkind, fslices = engine.factor_plan(prec, train_x.shape[0])
    f = engine.GPFactor(hyper, x, diag_add=wn, diag_scalar=hyper.noise, factor=fkind, factor_slices=fslices)
    f.perm, f.box = perm, box
    if cache is not None:
        cache.clear()
        cache["factor_key"] = key
        cache["factor"] = f
    return f


def _posterior(gp, hyper, train_x, train_y, test_x, train_var, test_var, want_var=False, want_cov=False,
               want_kxx=False):
    """mu, var, cov (and the padded device Kxx / cov buffers when want_kxx) of the latent posterior."""
    from . import engine
    train_x = np.asarray(to_numpy(train_x), dtype=np.float64)
    if train_x.ndim == 1:
        train_x = train_x[:, None]
    test_x = np.asarray(to_numpy(test_x), dtype=np.float64)
    if test_x.ndim == 1:
        test_x = test_x[:, None]
    train_y = np.asarray(to_numpy(train_y), dtype=np.float64).reshape(-1)
    ymean = float(np.mean(train_y))                                   # utils.py:294
    prec = getattr(gp, "precision", "fp64")
    # the mean / variance path does not depend on the order of the points: in INT8 digit mode both sets are sorted
    # along a Z curve so that far-apart (test tile, train chunk) pairs become all-zero digit tiles the GEMM skips
    reorder = (not want_cov) and engine.uses_digits(prec, len(train_y)) and len(train_y) >= engine.I8_REORDER_MIN
    f = _factor_for(gp, hyper, train_x, train_var, reorder=reorder)
    dev = f.L.device
    xs = engine.to_dev(test_x, device=dev)
    y0 = engine.to_dev(train_y - ymean, device=dev)
    tv = None if test_var is None else engine.to_dev(np.asarray(test_var, dtype=np.float64), device=dev)
    M = xs.shape[0]
    if not want_cov:
        tperm = None
        if reorder:
            y0 = y0.index_select(0, f.perm)
            tperm, _, _ = engine.morton_perm(xs, f.box[0], f.box[1])
            xs = xs.index_select(0, tperm).contiguous()
            tv = None if tv is None else tv.index_select(0, tperm).contiguous()
        mu, var = f.mean_var(xs, y0, ymean, tv, want_var=want_var, precision=prec)
        if tperm is not None:
            inv = torch.empty_like(tperm)
            inv[tperm] = torch.arange(M, device=dev)
            mu = mu.index_select(0, inv)
            var = var.index_select(0, inv) if want_var else None
        mu_h = mu.cpu().numpy()
        var_h = var.cpu().numpy() if want_var else None
        f.check()
        return mu_h, var_h, None
    alpha, _ = f.solve(y0)
    Ks, part = f.cross(xs, alpha)
    mu = engine.rowsum(part, 1.0, ymean, rows=M)
    V, _ = f.whiten(Ks, want_V=True, want_norm=False)
    Mpad = Ks.shape[0]
    Kxx, _ = engine.kbuild(hyper, xs, None, Mpad, Mpad, diag_add=tv, diag_scalar=0.0, pad_identity=True)   # utils.py:297
    cov = Kxx.clone() if want_kxx else Kxx
    engine.gemm_nt(V, V, cov, -1.0, 1.0)                               # utils.py:305
    mu_h = mu.cpu().numpy()
    f.check()
    if want_kxx:
        return mu_h, None, (Kxx, cov)
    cov_h = cov[:M, :M].cpu().numpy()
    return mu_h, (np.diag(cov_h).copy() if want_var else None), cov_h


def predictive_distribution(gp, train_x, train_y, test_x, train_var=None, test_var=None, return_var=False,
                            return_cov=False, return_mi=False):
    """utils.py:293-319 with the reference's return convention:
    mu | (mu, var) | (mu, cov) | (mu, mi) | (mu, cov, mi)."""
    from . import engine
    hyper = gp.hyper()
    if not (return_var or return_cov or return_mi):
        mu, _, _ = _posterior(gp, hyper, train_x, train_y, test_x, train_var, test_var)
        return mu
    if return_var and not (return_cov or return_mi):
        mu, var, _ = _posterior(gp, hyper, train_x, train_y, test_x, train_var, test_var, want_var=True)
        return mu, var
    M = len(test_x)
    mu, _, (Kxx, cov) = _posterior(gp, hyper, train_x, train_y, test_x, train_var, test_var, want_cov=True, want_kxx=True)
    res = None
    cov_h = None
    if return_cov:
        cov_h = cov[:M, :M].cpu().numpy()
        res = (mu, cov_h)
    if return_mi:
        # entropy_from_cov(cov_xx) - entropy_from_cov(cov): the M*CONST terms cancel (utils.py:314)
        mi = 0.5 * (engine.chol_logdet(Kxx, M) - engine.chol_logdet(cov, M))
        res = (mu, mi)
    if return_cov and return_mi:
        res = (mu, cov_h, mi)
    return res


def predictive_distribution_prefixes(gp, train_x, train_y, test_x, train_var, counts, test_var=None,
                                     return_var=False, return_cov=False, return_mi=False):
    """[predictive_distribution(gp, train_x[:c], train_y[:c], test_x, train_var[:c], ...) for c in counts]
    -- the loop of Agent.prediction_vs_distance (agent.py:497-518) -- from ONE factorisation: the Cholesky
    factor (and its inverse) of a leading sub-matrix is the leading block of the full factor, so every prefix
    posterior is a running sum over the columns of V = K(X*,X) Linv^T.  `counts` must be ascending."""
    from . import engine, _lib
    from ._lib import call, ptr, stream
    counts = [int(c) for c in counts]
    assert counts == sorted(counts) and counts[0] >= 1 and counts[-1] <= len(train_y)
    hyper = gp.hyper()
    nmax = counts[-1]
    train_x = np.asarray(to_numpy(train_x), dtype=np.float64)[:nmax]
    if train_x.ndim == 1:
        train_x = train_x[:, None]
    test_x = np.asarray(to_numpy(test_x), dtype=np.float64)
    if test_x.ndim == 1:
        test_x = test_x[:, None]
    y = np.asarray(to_numpy(train_y), dtype=np.float64).reshape(-1)[:nmax]
    tvar = None if train_var is None else np.asarray(train_var, dtype=np.float64)[:nmax]
    f = _factor_for(gp, hyper, train_x, tvar)
    dev = f.L.device
    xs = engine.to_dev(test_x, device=dev)
    M = xs.shape[0]
    Ks, _ = f.cross(xs)
    V, _ = f.whiten(Ks, want_V=True, want_norm=False)
    yp = torch.zeros(f.Npad, dtype=torch.float64, device=dev)
    yp[:nmax] = engine.to_dev(y, device=dev)
    ones = torch.zeros(f.Npad, dtype=torch.float64, device=dev)
    ones[:nmax] = 1.0
    beta, gamma = torch.empty_like(yp), torch.empty_like(yp)
    call("algp_gemv_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(yp), ptr(beta), stream())
    call("algp_gemv_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(ones), ptr(gamma), stream())
    pre = engine.to_dev(np.asarray(counts, dtype=np.int32), dtype=torch.int32, device=dev)
    out = torch.empty((len(counts), M, 3), dtype=torch.float64, device=dev)
    call("algp_prefix_reduce", ptr(V), V.stride(0), M, ptr(beta), ptr(gamma), ptr(pre), len(counts), ptr(out), stream())
    sums = out.cpu().numpy()
    f.check()
    tv = np.zeros(M) if test_var is None else np.asarray(test_var, dtype=np.float64)
    want_full = return_cov or return_mi
    Kxx = None
    if want_full:
        Mpad = Ks.shape[0]
        tvd = None if test_var is None else engine.to_dev(tv, device=dev)
        Kxx, _ = engine.kbuild(hyper, xs, None, Mpad, Mpad, diag_add=tvd, diag_scalar=0.0, pad_identity=True)
        ld_xx = engine.chol_logdet(Kxx.clone(), M) if return_mi else None
    results = []
    for i, c in enumerate(counts):
        ybar = float(np.mean(y[:c]))                                    # utils.py:294 on the prefix
        mu = sums[i, :, 0] - ybar * sums[i, :, 1] + ybar
        if not (return_var or return_cov or return_mi):
            results.append(mu)
            continue
        res = None
        if return_var:
            res = (mu, hyper.outputscale + tv - sums[i, :, 2])
        if want_full:
            Vp = V.clone()
            Vp[:, c:] = 0.0
            cov = Kxx.clone()
            engine.gemm_nt(Vp, Vp, cov, -1.0, 1.0)                      # utils.py:305 on the prefix
            cov_h = cov[:M, :M].cpu().numpy() if return_cov else None
            if return_cov:
                res = (mu, cov_h)
            if return_mi:
                mi = 0.5 * (ld_xx - engine.chol_logdet(cov, M))         # utils.py:314
                res = (mu, mi)
            if return_cov and return_mi:
                res = (mu, cov_h, mi)
        results.append(res)
    return results
