"""CPU restatement (NumPy) of the reference's GP / information-gain hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This file is the
checker the CUDA path is compared with, and the timed CPU baseline of
``bench.py``.  It is never imported by ``algp_b200``.

Every function cites the reference code it follows (paths are relative to
``/root/reference``).  Two dtype flavours exist:

``fp64``   everything in float64, Cholesky based.  Truth for the fp64 tier
           (rel 1e-9 mean / variance, 1e-8 log-det).
``ref32``  the reference's own dtype flow: ``to_torch`` forces float32
           (utils.py:13-20), so the kernel matrix is evaluated and stored in
           float32 (models.py:170-173); ``predictive_distribution`` inverts it
           with an explicit float32 ``np.linalg.inv`` (utils.py:300) and the
           entropies are float64 ``slogdet`` over the float32-rounded entries
           (utils.py:188-194, agent.py:308).  Truth for the 1e-4 tier and the
           timed "reference CPU path".

PARITY PIN.  The reference has no tests, golden vectors or seeds.  Its
kernel evaluation and marginal likelihood live in gpytorch (2018 "Beta",
un-pinned, README.md:9), which is absent here, so those two pieces are
restated from the published closed forms: **parity unpinned** for the kernel
closed forms and the MLL.  Everything the reference computes in-tree
(cov_mat's noise handling, predictive_distribution, entropy_from_cov,
get_sampled_dataset, greedy, best_path) IS pinned: tests/golden/make_golden.py
imports the reference's own utils.py / agent.py / models.py in this container
(with a small stand-in for the missing gpytorch package) and freezes their
outputs under tests/golden/, and tests/test_oracle_golden.py checks this file
against them.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "CONST", "Theta", "OracleGP", "kernel_matrix", "entropy_from_cov",
    "predictive_distribution", "predictive_distribution_chol",
    "get_sampled_dataset", "precisions_from_flags", "greedy_literal",
    "best_path_literal", "set_entropy_literal", "posterior_state",
    "greedy_restructured", "score_sets_restructured", "LeanEpisode", "mll_loss", "mll_loss_grad",
    "gaussian_mixture_field",
]

# utils.py:10
CONST = 0.5 * np.log(2 * np.pi * np.exp(1))


class Theta(object):
    """Hyper-parameters of ``ScaleKernel(RBF | Matern(1.5), ARD d)`` plus the
    Gaussian likelihood noise, stored as logs like the reference-era gpytorch
    (models.py:216-220, models.py:180, run.py:36-37)."""

    def __init__(self, log_lengthscale, log_outputscale=0.0, log_noise=0.0, kind="rbf"):
        self.log_lengthscale = np.atleast_1d(np.asarray(log_lengthscale, dtype=np.float64))
        self.log_outputscale = float(log_outputscale)
        self.log_noise = float(log_noise)
        if kind is None:
            kind = "rbf"          # models.py:216-218
        if kind not in ("rbf", "matern"):
            raise NotImplementedError(kind)  # models.py:226-227
        self.kind = kind

    @classmethod
    def from_values(cls, lengthscale, outputscale=1.0, noise=1.0, kind="rbf"):
        return cls(np.log(np.asarray(lengthscale, dtype=np.float64)), np.log(outputscale), np.log(noise), kind)

    @property
    def d(self):
        return self.log_lengthscale.shape[0]


def kernel_matrix(theta, x1, x2=None, flavour="fp64"):
    """``ScaleKernel(RBFKernel | MaternKernel(nu=1.5))(x1, x2).evaluate()``.

    Closed forms (gpytorch is absent, version un-pinned):
    r^2 = sum_j ((x_j - x'_j)/l_j)^2, RBF k = s^2 exp(-r^2/2),
    Matern-1.5 k = s^2 (1 + sqrt(3) r) exp(-sqrt(3) r).
    Call sites: models.py:170 (x2 None) and models.py:173.
    ``ref32`` evaluates in float32 because ``to_torch`` builds FloatTensors
    (utils.py:19) and the parameters are float32 ``nn.Parameter``s.
    """
    dt = np.float64 if flavour == "fp64" else np.float32
    x1 = np.asarray(x1, dtype=np.float64).astype(dt)
    x2 = x1 if x2 is None else np.asarray(x2, dtype=np.float64).astype(dt)
    if x1.ndim == 1:
        x1 = x1[:, None]
    if x2.ndim == 1:
        x2 = x2[:, None]
    ls = np.exp(theta.log_lengthscale.astype(dt)).astype(dt)
    os_ = dt(np.exp(dt(theta.log_outputscale)))
    a = x1 / ls
    b = x2 / ls
    # pairwise squared distance by explicit differences (no |a|^2+|b|^2-2ab
    # cancellation); chunked over rows to bound memory.
    n1, n2 = a.shape[0], b.shape[0]
    out = np.empty((n1, n2), dtype=dt)
    step = max(1, int(4e6 // max(1, n2)))
    s3 = dt(np.sqrt(3.0))
    for lo in range(0, n1, step):
        hi = min(n1, lo + step)
        r2 = np.zeros((hi - lo, n2), dtype=dt)
        for j in range(a.shape[1]):
            diff = a[lo:hi, j:j + 1] - b[None, :, j]
            r2 += diff * diff
        if theta.kind == "rbf":
            out[lo:hi] = os_ * np.exp(dt(-0.5) * r2)
        else:
            r = np.sqrt(r2)
            out[lo:hi] = os_ * (dt(1.0) + s3 * r) * np.exp(-s3 * r)
    return out


class OracleGP(object):
    """The slice of ``models.GPR`` the hot path uses: ``cov_mat`` with explicit
    hyper-parameters (the Adam trajectory of ``GPR.fit`` is not a parity
    target; the MLL value and gradient at a given theta are, see mll_loss)."""

    def __init__(self, theta, flavour="fp64"):
        assert flavour in ("fp64", "ref32")
        self.theta = theta
        self.flavour = flavour

    @property
    def noise(self):
        # likelihood.log_noise.exp().item()  (models.py:180)
        if self.flavour == "ref32":
            return float(np.exp(np.float32(self.theta.log_noise)))
        return float(np.exp(self.theta.log_noise))

    def cov_mat(self, x1, x2=None, white_noise_var=None, add_likelihood_var=False):
        """models.py:161-181.  Returns s^2 k(x1,x2) [+ diag(white_noise_var)]
        [+ sigma_n^2 I]; the WhiteNoiseKernel term is never included on its
        own (kernel_covar_module only, models.py:170,173)."""
        if x2 is not None and np.array_equal(np.asarray(x1), np.asarray(x2)):
            x2 = None                                   # torch.equal branch, models.py:169
        cov = kernel_matrix(self.theta, x1, x2, self.flavour)
        if white_noise_var is not None:
            cov += np.diag(white_noise_var).astype(cov.dtype)     # models.py:175-176 (in-place, keeps dtype)
        if add_likelihood_var:
            cov += (self.noise * np.eye(len(cov))).astype(cov.dtype)  # models.py:179-180
        return cov


def entropy_from_cov(cov, constant=CONST):
    """utils.py:188-194: k*CONST + 0.5*logabsdet(cov) via LU slogdet (float64
    LAPACK for float64 input, float32 input is promoted by the caller's
    ``+ np.diag(var)``)."""
    if constant is None:
        constant = CONST
    return cov.shape[0] * constant + 0.5 * np.linalg.slogdet(cov)[1].item()


def predictive_distribution(gp, train_x, train_y, test_x, train_var=None, test_var=None,
                            return_var=False, return_cov=False, return_mi=False):
    """Literal utils.py:293-319 (explicit inverse, flag-dependent returns)."""
    train_y = np.asarray(train_y)
    train_y_mean = np.mean(train_y)                                          # utils.py:294
    cov_aa = gp.cov_mat(x1=train_x, white_noise_var=train_var, add_likelihood_var=True)   # :296
    cov_xx = gp.cov_mat(x1=test_x, white_noise_var=test_var)                 # :297
    cov_xa = gp.cov_mat(x1=test_x, x2=train_x)                               # :298
    mat1 = np.dot(cov_xa, np.linalg.inv(cov_aa))                             # :300
    mu = np.dot(mat1, (train_y - train_y_mean)) + train_y_mean               # :301
    if not (return_var or return_cov or return_mi):
        return mu
    cov = cov_xx - np.dot(mat1, cov_xa.T)                                    # :305
    res = None
    if return_var:
        res = (mu, np.diag(cov))
    if return_cov:
        res = (mu, cov)
    if return_mi:
        mi = entropy_from_cov(cov_xx) - entropy_from_cov(cov)                # :314
        res = (mu, mi)
    if return_cov and return_mi:
        res = (mu, cov, mi)
    return res


def predictive_distribution_chol(gp, train_x, train_y, test_x, train_var=None, test_var=None,
                                 return_var=False, return_cov=False, return_mi=False):
    """Same quantities as utils.py:293-319 through a float64 Cholesky instead
    of the explicit inverse: the fp64 truth (mathematically identical)."""
    train_y = np.asarray(train_y, dtype=np.float64)
    m = np.mean(train_y)
    A = np.asarray(gp.cov_mat(x1=train_x, white_noise_var=train_var, add_likelihood_var=True), dtype=np.float64)
    Kxa = np.asarray(gp.cov_mat(x1=test_x, x2=train_x), dtype=np.float64)
    import scipy.linalg as sla
    L = np.linalg.cholesky(A)
    beta = sla.solve_triangular(L, train_y - m, lower=True)
    alpha = sla.solve_triangular(L, beta, lower=True, trans="T")
    mu = Kxa @ alpha + m
    if not (return_var or return_cov or return_mi):
        return mu
    V = sla.solve_triangular(L, Kxa.T, lower=True)          # [N, M]
    res = None
    if return_var and not (return_cov or return_mi):
        # diag only: never forms the M x M matrix (feasible at M = 65536)
        from numpy import einsum
        kss = np.full(Kxa.shape[0], np.exp(gp.theta.log_outputscale))
        if test_var is not None:
            kss = kss + np.asarray(test_var, dtype=np.float64)
        return mu, kss - einsum("ij,ij->j", V, V)
    Kxx = np.asarray(gp.cov_mat(x1=test_x, white_noise_var=test_var), dtype=np.float64)
    cov = Kxx - V.T @ V
    if return_var:
        res = (mu, np.diag(cov))
    if return_cov:
        res = (mu, cov)
    if return_mi:
        mi = entropy_from_cov(Kxx) - entropy_from_cov(cov)
        res = (mu, mi)
    if return_cov and return_mi:
        res = (mu, cov, mi)
    return res


# ----------------------------------------------------------------------------
# Agent-side host logic
# ----------------------------------------------------------------------------

def get_sampled_dataset(static_data, mobile_data, static_std, mobile_std):
    """agent.py:92-117: fuse static / mobile readings per location."""
    all_y, all_var, indices = [], [], []
    for i in range(len(static_data)):
        if len(mobile_data[i]) > 0 and len(static_data[i]) > 0:
            yc = np.mean(mobile_data[i])
            ys = np.mean(static_data[i])
            yeq = (mobile_std ** 2 * ys + static_std ** 2 * yc) / (mobile_std ** 2 + static_std ** 2)
            var = 1 / (1 / (static_std ** 2) + 1 / (mobile_std ** 2))
        elif len(static_data[i]) > 0:
            yeq = np.mean(static_data[i])
            var = static_std ** 2
        elif len(mobile_data[i]) > 0:
            yeq = np.mean(mobile_data[i])
            var = mobile_std ** 2
        else:
            continue
        all_y.append(yeq)
        all_var.append(var)
        indices.append(i)
    return indices, np.array(all_y), np.array(all_var)


def precisions_from_flags(static_sampled, mobile_sampled, static_std, mobile_std):
    """pi_j = st_j/sigma_s^2 + mo_j/sigma_m^2  (agent.py:298-307: flags are
    booleans, repeated readings do not add precision)."""
    return np.asarray(static_sampled, dtype=np.float64) / static_std ** 2 + \
        np.asarray(mobile_sampled, dtype=np.float64) / mobile_std ** 2


def set_entropy_literal(cov_matrix, static_sampled, mobile_sampled, static_std, mobile_std):
    """H(S) exactly as agent.py:299-309 / 380-387 builds it: fancy-indexed
    copy + np.diag(var), then slogdet."""
    n = len(static_sampled)
    mobile_var = np.full(n, np.inf)
    mobile_var[mobile_sampled] = mobile_std ** 2
    static_var = np.full(n, np.inf)
    static_var[static_sampled] = static_std ** 2
    sampled = static_sampled | mobile_sampled
    var = 1.0 / (1.0 / static_var[sampled] + 1.0 / mobile_var[sampled])
    cov_a = cov_matrix[sampled].T[sampled].T + np.diag(var)
    return entropy_from_cov(cov_a)


def _mi_terms(cov_matrix, static_var, mobile_var, sampled):
    # agent.py:330-339 / 388-397
    cov_abar = cov_matrix[~sampled].T[~sampled].T
    ent_abar = entropy_from_cov(cov_abar)
    precision = 1.0 / static_var + 1.0 / mobile_var
    precision[precision == 0] = np.inf
    var = 1.0 / precision
    cov_all = cov_matrix + np.diag(var)
    ent_all = entropy_from_cov(cov_all)
    return ent_abar, ent_all


def greedy_literal(cov_matrix, static_sampled, mobile_sampled, static_std, mobile_std,
                   num_samples, criterion="entropy", return_utilities=False, candidates=None):
    """agent.py:295-356, loop for loop.  ``candidates`` (not in the reference)
    restricts the inner loop to a subset so the CPU baseline can time a
    bounded sample; None = all n locations as in the reference."""
    n = len(static_sampled)
    mobile_sampled = np.array(mobile_sampled, dtype=bool)
    static_sampled = np.array(static_sampled, dtype=bool)
    mobile_var = np.full(n, np.inf)
    mobile_var[mobile_sampled] = mobile_std ** 2
    static_var = np.full(n, np.inf)
    static_var[static_sampled] = static_std ** 2

    sampled = static_sampled | mobile_sampled
    var = 1.0 / (1.0 / static_var[sampled] + 1.0 / mobile_var[sampled])
    cov_v = cov_matrix[sampled].T[sampled].T + np.diag(var)
    ent_v = entropy_from_cov(cov_v)

    cumm_utilities, new_samples, all_utilities = [], [], []
    for _ in range(num_samples):
        utilities = np.full(n, -np.inf)
        cond = ent_v + sum(cumm_utilities)
        it = range(n) if candidates is None else candidates
        for i in it:
            if static_sampled[i]:
                continue
            static_sampled[i] = True
            static_var[i] = static_std ** 2
            sampled = static_sampled | mobile_sampled
            var = 1.0 / (1.0 / static_var[sampled] + 1.0 / mobile_var[sampled])
            cov_a = cov_matrix[sampled].T[sampled].T + np.diag(var)
            ent_a = entropy_from_cov(cov_a)
            if criterion == "mutual_information":
                ent_abar, ent_all = _mi_terms(cov_matrix, static_var, mobile_var, sampled)
                ut = ent_a + ent_abar - ent_all
            else:
                ut = ent_a - cond
            utilities[i] = ut
            static_sampled[i] = False
            static_var[i] = np.inf
        best_sample = np.argmax(utilities)
        cumm_utilities.append(utilities[best_sample])
        new_samples.append(best_sample)
        all_utilities.append(utilities)
        static_sampled[best_sample] = True
        static_var[best_sample] = static_std ** 2
    if return_utilities:
        return new_samples, np.array(all_utilities)
    return new_samples


def best_path_literal(cov_matrix, static_sampled, org_mobile_sampled, static_std, mobile_std,
                      paths_mobile_indices, static_indices, criterion="entropy", return_utilities=False):
    """agent.py:358-403, loop for loop (single path -> 0 without scoring)."""
    if len(paths_mobile_indices) == 1 and not return_utilities:
        return 0
    n = len(static_sampled)
    org_mobile_sampled = np.array(org_mobile_sampled, dtype=bool)
    static_sampled = np.array(static_sampled, dtype=bool)
    static_sampled[static_indices] = True
    static_var = np.full(n, np.inf)
    static_var[static_sampled] = static_std ** 2
    all_ut = []
    for i in range(len(paths_mobile_indices)):
        mobile_sampled = np.copy(org_mobile_sampled)
        mobile_indices = paths_mobile_indices[i]
        mobile_sampled[mobile_indices] = True
        mobile_var = np.full(n, np.inf)
        mobile_var[mobile_sampled] = mobile_std ** 2
        sampled = static_sampled | mobile_sampled
        var = 1.0 / (1.0 / static_var[sampled] + 1.0 / mobile_var[sampled])
        cov_a = cov_matrix[sampled].T[sampled].T + np.diag(var)
        ent_a = entropy_from_cov(cov_a)
        if criterion == "mutual_information":
            ent_abar, ent_all = _mi_terms(cov_matrix, static_var, mobile_var, sampled)
            ut = ent_a + ent_abar - ent_all
        else:
            ut = ent_a
        all_ut.append(ut)
    idx = int(np.argmax(all_ut))
    if return_utilities:
        return idx, np.array(all_ut)
    return idx


# ----------------------------------------------------------------------------
# Restructured forms -- what the CUDA path computes; asserted
# equal to the literal forms in tests/test_oracle.py.
# ----------------------------------------------------------------------------

def posterior_state(cov_matrix, pi0):
    """Base set B = {pi0 > 0}: A_B = Sigma_BB + diag(1/pi0_B) = L L^T,
    W = L^-1 Sigma_{B,:}, P = Sigma - W^T W, H(B)."""
    import scipy.linalg as sla
    cov = np.asarray(cov_matrix, dtype=np.float64)
    pi0 = np.asarray(pi0, dtype=np.float64)
    base = np.nonzero(pi0 > 0)[0]
    if len(base) == 0:
        return dict(base=base, L=np.zeros((0, 0)), W=np.zeros((0, cov.shape[0])), P=cov.copy(), H=0.0)
    A = cov[np.ix_(base, base)] + np.diag(1.0 / pi0[base])
    L = np.linalg.cholesky(A)
    W = sla.solve_triangular(L, cov[base, :], lower=True)
    P = cov - W.T @ W
    H = len(base) * CONST + np.sum(np.log(np.diag(L)))
    return dict(base=base, L=L, W=W, P=P, H=H)


def score_sets_restructured(P, pi0, idx, delta, H_base=0.0):
    """H(S1) for candidate sets: rows of ``idx`` (int, -1 = empty slot) with
    per-slot precision increments ``delta``:
    H(S1) = H(B) + n_new*CONST + 0.5*[logdet(I + D P_CC D) - sum_slots(log(pi0+delta) - [pi0>0] log pi0)].
    Duplicate indices inside a row are idempotent (agent.py:377)."""
    idx = np.asarray(idx)
    delta = np.asarray(delta, dtype=np.float64)
    out = np.empty(idx.shape[0])
    for c in range(idx.shape[0]):
        seen = set()
        ii, dd = [], []
        for j, dj in zip(idx[c], delta[c]):
            if j < 0 or dj <= 0 or int(j) in seen:
                continue
            seen.add(int(j))
            ii.append(int(j))
            dd.append(dj)
        if not ii:
            out[c] = H_base
            continue
        ii = np.array(ii)
        dd = np.array(dd)
        sq = np.sqrt(dd)
        M = np.eye(len(ii)) + sq[:, None] * P[np.ix_(ii, ii)] * sq[None, :]
        ld = 2.0 * np.sum(np.log(np.diag(np.linalg.cholesky(M))))
        p0 = pi0[ii]
        n_new = np.sum(p0 == 0)
        corr = np.sum(np.log(p0 + dd)) - np.sum(np.log(p0[p0 > 0]))
        out[c] = H_base + n_new * CONST + 0.5 * (ld - corr)
    return out


def greedy_restructured(cov_matrix, static_sampled, mobile_sampled, static_std, mobile_std, num_samples,
                        return_utilities=False):
    """Entropy-criterion greedy (agent.py:313-354) as rank-1 downdates of P:
    ut_i = [pi0_i == 0]*CONST + 0.5*[log1p(d P_ii) - log(pi0_i + d) + [pi0_i>0] log pi0_i], d = 1/sigma_s^2;
    after committing j: P <- P - P_:j P_j: / (P_jj + 1/d)."""
    static_sampled = np.array(static_sampled, dtype=bool)
    mobile_sampled = np.array(mobile_sampled, dtype=bool)
    pi = precisions_from_flags(static_sampled, mobile_sampled, static_std, mobile_std)
    P = posterior_state(cov_matrix, pi)["P"]
    d = 1.0 / static_std ** 2
    picks, uts = [], []
    for _ in range(num_samples):
        pd = np.diag(P)
        with np.errstate(divide="ignore"):
            ut = np.where(pi == 0, CONST, 0.0) + 0.5 * (np.log1p(d * pd) - np.log(pi + d)
                                                         + np.where(pi > 0, np.log(np.where(pi > 0, pi, 1.0)), 0.0))
        ut[static_sampled] = -np.inf
        j = int(np.argmax(ut))
        picks.append(j)
        uts.append(ut)
        col = P[:, j].copy()
        P = P - np.outer(col, col) / (P[j, j] + 1.0 / d)
        pi[j] += d
        static_sampled[j] = True
    if return_utilities:
        return picks, np.array(uts)
    return picks


class LeanEpisode(object):
    """The episode of Agent.run_ipp (agent.py:133-201: greedy picks, path scoring, commits) in the restructured form of
    SURVEY.md 9.3 WITHOUT ever forming an n x n matrix, so that BASELINE configs[4] (a 200 x 200 field, n = 40 000) can
    be checked on the host: only W = L^-1 Sigma_{B,:} (one row per sampled reading, n columns) and diag(P) are kept;
    a commit at location j with precision increment delta appends the row  P_{j,:} / sqrt(P_jj + 1/delta)  to W.
    Pinned against greedy_literal / best_path_literal on small fields in tests/test_oracle.py."""

    def __init__(self, theta, X, static_sampled, mobile_sampled, static_std, mobile_std):
        import scipy.linalg as sla
        self.theta, self.X = theta, np.asarray(X, dtype=np.float64)
        self.ss, self.ms = static_std, mobile_std
        self.static = np.array(static_sampled, dtype=bool)
        self.mobile = np.array(mobile_sampled, dtype=bool)
        self.pi = precisions_from_flags(self.static, self.mobile, static_std, mobile_std)
        self.prior = np.exp(theta.log_outputscale) + np.exp(theta.log_noise)          # diag of cov_matrix (agent.py:90)
        base = np.nonzero(self.pi > 0)[0]
        n = len(self.X)
        if len(base):
            A = self._sigma(base, base) + np.diag(1.0 / self.pi[base])
            L = np.linalg.cholesky(A)
            self.W = sla.solve_triangular(L, self._sigma(base, None), lower=True)
            self.H = len(base) * CONST + np.sum(np.log(np.diag(L)))
        else:
            self.W = np.zeros((0, n))
            self.H = 0.0
        self.diagP = self.prior - np.sum(self.W * self.W, axis=0)

    def _sigma(self, rows, cols):
        """Sigma[rows, cols] = s^2 k + sigma_n^2 [row location == column location] (cov_matrix of agent.py:90)."""
        rows = np.asarray(rows)
        xc = self.X if cols is None else self.X[np.asarray(cols)]
        K = kernel_matrix(self.theta, self.X[rows], xc, "fp64")
        cidx = np.arange(len(self.X)) if cols is None else np.asarray(cols)
        K[rows[:, None] == cidx[None, :]] += np.exp(self.theta.log_noise)
        return K

    def _P(self, rows, cols):
        rows, cidx = np.asarray(rows), (np.arange(len(self.X)) if cols is None else np.asarray(cols))
        return self._sigma(rows, cols) - self.W[:, rows].T @ self.W[:, cidx]

    def commit(self, j, delta, make_static):
        row = self._P([j], None)[0]
        w = row / np.sqrt(row[j] + 1.0 / delta)
        self.W = np.vstack([self.W, w[None, :]])
        self.diagP = self.diagP - w * w
        self.pi[j] += delta
        if make_static:
            self.static[j] = True
        else:
            self.mobile[j] = True

    def greedy(self, num_samples):
        """agent.py:313-354 (entropy criterion): picks, and H advanced by the chosen utilities."""
        d = 1.0 / self.ss ** 2
        picks = []
        for _ in range(num_samples):
            pi = self.pi
            with np.errstate(divide="ignore"):
                ut = np.where(pi == 0, CONST, 0.0) + 0.5 * (np.log1p(d * self.diagP) - np.log(pi + d)
                                                             + np.where(pi > 0, np.log(np.where(pi > 0, pi, 1.0)), 0.0))
            ut[self.static] = -np.inf
            j = int(np.argmax(ut))
            picks.append(j)
            self.H += ut[j]
            self.commit(j, d, True)
        return picks

    def score_paths(self, paths):
        """agent.py:373-400 (entropy criterion): H(S + mobile readings of the path) for every path (rows of -1-padded
        indices; already-mobile locations and repeats add nothing)."""
        dm = 1.0 / self.ms ** 2
        out = np.empty(len(paths))
        for c, path in enumerate(paths):
            seen, ii = set(), []
            for j in path:
                j = int(j)
                if j < 0 or self.mobile[j] or j in seen:
                    continue
                seen.add(j)
                ii.append(j)
            if not ii:
                out[c] = self.H
                continue
            ii = np.array(ii)
            M = np.eye(len(ii)) + dm * self._P(ii, ii)
            ld = 2.0 * np.sum(np.log(np.diag(np.linalg.cholesky(M))))
            p0 = self.pi[ii]
            corr = np.sum(np.log(p0 + dm)) - np.sum(np.log(p0[p0 > 0]))
            out[c] = self.H + np.sum(p0 == 0) * CONST + 0.5 * (ld - corr)
        return out

    def commit_path(self, path, score):
        dm = 1.0 / self.ms ** 2
        for j in path:
            j = int(j)
            if j >= 0 and not self.mobile[j]:
                self.commit(j, dm, False)
        self.H = float(score)


# ----------------------------------------------------------------------------
# Marginal likelihood (models.py:145-149)
# ----------------------------------------------------------------------------

def _train_cov(theta, x, var):
    K = kernel_matrix(theta, x, None, "fp64")
    return K, K + np.diag(np.asarray(var, dtype=np.float64)) + np.exp(theta.log_noise) * np.eye(len(K))


def mll_loss(theta, x, y, var):
    """loss = -(1/N) log N(y0; 0, K + diag(var) + sigma_n^2 I), y0 = y - mean(y)
    (models.py:129-130, 148; gpytorch's ExactMarginalLogLikelihood divides by N)."""
    y = np.asarray(y, dtype=np.float64)
    y0 = y - y.mean()
    _, A = _train_cov(theta, x, var)
    import scipy.linalg as sla
    L = np.linalg.cholesky(A)
    beta = sla.solve_triangular(L, y0, lower=True)
    N = len(y0)
    ll = -0.5 * beta @ beta - np.sum(np.log(np.diag(L))) - 0.5 * N * np.log(2 * np.pi)
    return -ll / N


def mll_loss_grad(theta, x, y, var):
    """Gradient of mll_loss w.r.t. (log_lengthscale[d], log_outputscale, log_noise):
    d(-ll)/dtheta = -0.5 tr((alpha alpha^T - A^-1) dA/dtheta), divided by N."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, dtype=np.float64)
    y0 = y - y.mean()
    K, A = _train_cov(theta, x, var)
    N = len(y0)
    Ainv = np.linalg.inv(A)
    alpha = Ainv @ y0
    G = np.outer(alpha, alpha) - Ainv
    ls = np.exp(theta.log_lengthscale)
    g_ls = np.zeros(theta.d)
    for p in range(theta.d):
        dp = ((x[:, p:p + 1] - x[None, :, p]) / ls[p]) ** 2
        if theta.kind == "rbf":
            dK = K * dp
        else:
            r2 = np.zeros_like(K)
            for q in range(theta.d):
                r2 += ((x[:, q:q + 1] - x[None, :, q]) / ls[q]) ** 2
            r = np.sqrt(r2)
            dK = 3.0 * np.exp(theta.log_outputscale) * np.exp(-np.sqrt(3.0) * r) * dp
        g_ls[p] = 0.5 * np.sum(G * dK)
    g_os = 0.5 * np.sum(G * K)
    g_noise = 0.5 * np.trace(G) * np.exp(theta.log_noise)
    return -np.concatenate([g_ls, [g_os, g_noise]]) / N


# ----------------------------------------------------------------------------
# Synthetic field (utils.py:90-108), seeded (the reference never seeds)
# ----------------------------------------------------------------------------

def gaussian_mixture_field(num_rows, num_cols, k=5, min_var=10, max_var=100, seed=1):
    """generate_gaussian_data(algo='sum') with numpy.random.default_rng(seed)."""
    rng = np.random.default_rng(seed)
    xx, yy = np.meshgrid(np.arange(num_cols), np.arange(num_rows))
    grid = np.vstack([yy.flatten(), xx.flatten()]).transpose().astype(np.float64)
    means = np.vstack([rng.uniform(0, num_rows, size=k), rng.uniform(0, num_cols, size=k)]).transpose()
    variances = rng.uniform(min_var, max_var, size=k)
    y = np.zeros(num_rows * num_cols)
    for i in range(k):
        dist_sq = np.sum(np.square(grid - means[i].reshape(1, -1)), axis=1)
        y += np.exp(-dist_sq / variances[i])
    return grid, y
