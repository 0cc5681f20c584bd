"""Oracle self-consistency: the restructured forms the CUDA path computes
 equal the literal reference loops; MLL gradient vs finite
differences.  CPU only."""
import numpy as np
import pytest

import oracle as O


def small_problem(kind="rbf", n=60, seed=3, d=2):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 12, size=(n, d))
    th = O.Theta.from_values([2.0 + 0.5 * i for i in range(d)], 1.2, 0.03, kind)
    gp = O.OracleGP(th, "fp64")
    cov = gp.cov_mat(X, add_likelihood_var=True)
    static = np.zeros(n, bool)
    mobile = np.zeros(n, bool)
    static[rng.choice(n, 12, replace=False)] = True
    mobile[rng.choice(n, 15, replace=False)] = True
    return X, th, cov, static, mobile, rng


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_greedy_restructured_equals_literal(kind):
    X, th, cov, static, mobile, rng = small_problem(kind)
    ss, ms = 0.1, 1.0
    p_lit, u_lit = O.greedy_literal(cov, static, mobile, ss, ms, 4, return_utilities=True)
    p_res, u_res = O.greedy_restructured(cov, static, mobile, ss, ms, 4, return_utilities=True)
    assert [int(p) for p in p_lit] == p_res
    fin = np.isfinite(u_lit)
    assert (fin == np.isfinite(u_res)).all()
    np.testing.assert_allclose(u_res[fin], u_lit[fin], rtol=0, atol=5e-12)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_score_sets_restructured_equals_literal_best_path(kind):
    X, th, cov, static, mobile, rng = small_problem(kind)
    n = len(X)
    ss, ms = 0.1, 1.0
    static_idx = [int(i) for i in np.nonzero(~static)[0][:2]]
    paths = []
    for p in range(20):
        L = int(rng.integers(1, 9))
        path = [int(v) for v in rng.choice(n, L, replace=False)]
        if p % 3 == 0:
            path += path[:1]
        if p % 4 == 1:
            path[0] = static_idx[0]
        paths.append(path)
    best, ut = O.best_path_literal(cov, static, mobile, ss, ms, paths, static_idx, return_utilities=True)

    st2 = static.copy()
    st2[static_idx] = True
    pi0 = O.precisions_from_flags(st2, mobile, ss, ms)
    stt = O.posterior_state(cov, pi0)
    k = max(len(p) for p in paths)
    idx = np.full((len(paths), k), -1, dtype=np.int64)
    delta = np.zeros((len(paths), k))
    for c, p in enumerate(paths):
        idx[c, :len(p)] = p
        delta[c, :len(p)] = np.where(mobile[p], 0.0, 1.0 / ms ** 2)     # mobile flag is idempotent
    sc = O.score_sets_restructured(stt["P"], pi0, idx, delta, stt["H"])
    np.testing.assert_allclose(sc, ut, rtol=1e-12, atol=1e-10)
    assert int(np.argmax(sc)) == best
    # H(B) itself
    assert stt["H"] == pytest.approx(O.set_entropy_literal(cov, st2, mobile, ss, ms), rel=1e-13)


def test_posterior_variance_identity():
    """P_ii of posterior_state equals predictive_distribution's latent variance
    minus nothing: P = Sigma - W^T W with Sigma including sigma_n^2 I."""
    X, th, cov, static, mobile, rng = small_problem("rbf")
    pi0 = O.precisions_from_flags(static, mobile, 0.1, 1.0)
    stt = O.posterior_state(cov, pi0)
    B = stt["base"]
    A = cov[np.ix_(B, B)] + np.diag(1 / pi0[B])
    P = cov - cov[:, B] @ np.linalg.solve(A, cov[B, :])
    np.testing.assert_allclose(stt["P"], P, rtol=0, atol=1e-12)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_mll_grad_finite_differences(kind):
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 10, size=(40, 3))
    y = np.sin(x[:, 0]) + 0.1 * rng.normal(size=40)
    var = np.full(40, 0.01)
    th = O.Theta(np.log([1.5, 2.0, 3.0]), np.log(0.8), np.log(0.05), kind)
    g = O.mll_loss_grad(th, x, y, var)
    eps = 1e-6
    num = []
    for p in range(5):
        def shifted(s):
            ls = th.log_lengthscale.copy()
            os_, nz = th.log_outputscale, th.log_noise
            if p < 3:
                ls[p] += s
            elif p == 3:
                os_ += s
            else:
                nz += s
            return O.mll_loss(O.Theta(ls, os_, nz, kind), x, y, var)
        num.append((shifted(eps) - shifted(-eps)) / (2 * eps))
    np.testing.assert_allclose(g, num, rtol=1e-5, atol=1e-8)


def test_field_generator_seeded():
    g1, y1 = O.gaussian_mixture_field(16, 12, seed=1)
    g2, y2 = O.gaussian_mixture_field(16, 12, seed=1)
    assert g1.shape == (192, 2) and (y1 == y2).all() and y1.max() > 0
    assert (g1[1] == [0, 1]).all()          # row-major (row, col) grid as utils.py:91-92


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_lean_episode_equals_literal_loops(kind):
    """oracle.LeanEpisode (the memory-lean restructured episode used to check BASELINE configs[4] at its own 200 x 200
    scale) against the literal loops of agent.py:295-403 over three batches of greedy picks, path scoring and commits."""
    X, th, cov, static, mobile, rng = small_problem(kind, n=70, seed=5)
    ss, ms = 0.1, 1.0
    ep = O.LeanEpisode(th, X, static, mobile, ss, ms)
    st, mo = static.copy(), mobile.copy()
    assert ep.H == pytest.approx(O.set_entropy_literal(cov, st, mo, ss, ms), rel=1e-12)
    for b in range(3):
        picks = ep.greedy(2)
        assert picks == [int(p) for p in O.greedy_literal(cov, st, mo, ss, ms, 2)]
        paths = np.stack([rng.choice(len(X), 6, replace=False) for _ in range(9)]).astype(np.int64)
        paths[0, 4] = paths[0, 1]                           # a repeat inside a path
        paths[2, 5] = -1                                    # a ragged path
        scores = ep.score_paths(paths)
        lists = [[int(v) for v in row if v >= 0] for row in paths]
        best, ut = O.best_path_literal(cov, st, mo, ss, ms, lists, picks, return_utilities=True)
        np.testing.assert_allclose(scores, ut, rtol=1e-11, atol=1e-10)
        assert int(np.argmax(scores)) == best
        ep.commit_path(paths[best], scores[best])
        st[picks] = True
        mo[lists[best]] = True
        assert ep.H == pytest.approx(O.set_entropy_literal(cov, st, mo, ss, ms), rel=1e-11)
