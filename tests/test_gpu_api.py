"""Parity of the reference-facing API (GPR / predictive_distribution / entropy_from_cov /
Agent.greedy / Agent.best_path) with the oracle and with golden vectors frozen from the
reference's own code.  B200 only.

Tolerances (north star): fp64 tier rel 1e-9 on mean / variance (variance relative to the prior
scale s^2, SURVEY.md 7 "variance cancellation") and 1e-8 on log-dets / entropies; against the
reference's float32 flow only the 1e-4 tier is meaningful (its own float32 inverse limits it
to ~1e-3 of the prior scale on these conditionings)."""
import os
import types

import numpy as np
import pytest
import torch

import oracle as O
import algp_b200
from algp_b200 import engine
from gpu_helpers import dev, field_problem, hyper_pair

pytestmark = pytest.mark.gpu


def make_gpr(kind, log_ls, log_os, log_noise, x, y, var):
    gp = algp_b200.GPR(kernel_params={'type': kind})
    gp.reset(x, y, var)
    with torch.no_grad():
        gp.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(log_ls).view(1, 1, -1))
        gp.model.kernel_covar_module.log_outputscale.fill_(float(log_os))
        gp.likelihood.log_noise.fill_(float(log_noise))
    return gp


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_predictive_distribution_fp64_tier(kind):
    X, y, tr, ytr, rng = field_problem(24, 20, 300, seed=2)
    te = rng.choice(len(X), 150, replace=False)
    var = np.where(rng.random(300) < 0.5, 0.01, 1.0 / (1 / 0.01 + 1 / 1.0))
    tvar = np.full(150, 0.02)
    th, hy = hyper_pair([2.5, 3.5], 1.3, 0.02, kind)
    gp = make_gpr(kind, th.log_lengthscale, th.log_outputscale, th.log_noise, X[tr], ytr, var)
    ogp = O.OracleGP(th, "fp64")
    a = (X[tr], ytr, X[te], var)
    s2 = np.exp(th.log_outputscale)

    mu = algp_b200.predictive_distribution(gp, *a)
    mu_o = O.predictive_distribution_chol(ogp, *a)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-9, atol=1e-9 * np.abs(mu_o).max())

    mu2, var2 = algp_b200.predictive_distribution(gp, *a, return_var=True)
    _, var_o = O.predictive_distribution_chol(ogp, *a, return_var=True)
    np.testing.assert_allclose(mu2, mu_o, rtol=1e-9, atol=1e-9 * np.abs(mu_o).max())
    np.testing.assert_allclose(var2, var_o, rtol=0, atol=1e-9 * s2)

    mu3, cov3 = algp_b200.predictive_distribution(gp, *a, return_cov=True)
    _, cov_o = O.predictive_distribution_chol(ogp, *a, return_cov=True)
    np.testing.assert_allclose(cov3, cov_o, rtol=0, atol=1e-9 * s2)

    mu4, mi4 = algp_b200.predictive_distribution(gp, *a, test_var=tvar, return_mi=True)
    _, mi_o = O.predictive_distribution_chol(ogp, *a, test_var=tvar, return_mi=True)
    assert mi4 == pytest.approx(mi_o, rel=1e-8)

    mu5, cov5, mi5 = algp_b200.predictive_distribution(gp, *a, test_var=tvar, return_cov=True, return_mi=True)
    assert mi5 == pytest.approx(mi_o, rel=1e-8)
    assert cov5.shape == (150, 150)
    # literal explicit-inverse formula in float64 agrees too
    _, var_l = O.predictive_distribution(ogp, *a, return_var=True)
    np.testing.assert_allclose(var2, var_l, rtol=0, atol=1e-8 * s2)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_against_reference_golden_gp(golden_dir, kind):
    """Outputs of the reference's own utils.predictive_distribution / GPR.cov_mat (float32 flow)."""
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = make_gpr(kind, g["log_ls"], g["log_os"], g["log_noise"], g["train_x"], g["train_y"], g["train_var"])
    s2 = float(np.exp(g["log_os"]))
    K = gp.cov_mat(g["train_x"], white_noise_var=g["train_var"], add_likelihood_var=True)
    np.testing.assert_allclose(K, g["K_train_noise"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp.cov_mat(g["test_x"], g["train_x"]), g["K_test_train"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp.cov_mat(g["train_x"][:20], g["train_x"][:20].copy(), add_likelihood_var=True),
                               g["K_same_x2"], rtol=1e-5, atol=1e-6)
    gp.dtype = np.float32                         # the reference's own dtype
    K32 = gp.cov_mat(g["train_x"], white_noise_var=g["train_var"], add_likelihood_var=True)
    assert K32.dtype == np.float32
    np.testing.assert_allclose(K32, g["K_train_noise"], rtol=2e-6, atol=2e-7)
    gp.dtype = np.float64
    a = (g["train_x"], g["train_y"], g["test_x"], g["train_var"])
    mu, var = algp_b200.predictive_distribution(gp, *a, return_var=True)
    np.testing.assert_allclose(mu, g["pd_mu"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(var, g["pd_var"], rtol=0, atol=2e-3 * s2)
    _, cov = algp_b200.predictive_distribution(gp, *a, return_cov=True)
    np.testing.assert_allclose(cov, g["pd_cov"], rtol=0, atol=2e-3 * s2)
    _, mi = algp_b200.predictive_distribution(gp, *a, test_var=g["test_var"], return_mi=True)
    assert mi == pytest.approx(float(g["pd_mi"]), rel=2e-3)
    # entropy_from_cov on the reference's own matrices.  The golden numbers come from the reference's
    # float32 LU (np.linalg.slogdet of a float32 array, utils.py:193 via utils.py:314): 1e-4 tier.
    assert algp_b200.entropy_from_cov(g["K_train_noise"]) == pytest.approx(float(g["ent_K_train_noise"]), rel=1e-5)
    assert algp_b200.entropy_from_cov(g["pd_cov_tv"]) == pytest.approx(float(g["ent_pd_cov_tv"]), rel=1e-5)
    # the same matrix in float64, against a float64 slogdet: log-det tier (1e-8)
    K64 = g["K_train_noise"].astype(np.float64)
    assert algp_b200.entropy_from_cov(K64) == pytest.approx(O.entropy_from_cov(K64), rel=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        algp_b200.entropy_from_cov(-np.eye(5))
    # GPR.predict (noise-inclusive; un-pinned stand-in semantics) at the reference-fitted theta
    gp2 = make_gpr(kind, g["fit_log_ls"], g["fit_log_os"], g["fit_log_noise"], g["train_x"], g["train_y"], g["train_var"])
    pm, pv = gp2.predict(g["test_x"], return_std=True)
    np.testing.assert_allclose(pm, g["predict_mean"], rtol=0, atol=5e-3)
    np.testing.assert_allclose(pv, g["predict_var"], rtol=0, atol=5e-3)


def unflatten(counts, values):
    out, p = [], 0
    for c in counts:
        out.append(list(values[p:p + c]))
        p += c
    return out


class GoldenEnv(object):
    def __init__(self, g):
        self.X, self.test_X = g["X"], g["test_X"]
        self.num_samples = len(self.X)


def agent_from_golden(g, kind):
    """The reference's Agent state (sample lists, theta) re-created on the accelerated Agent."""
    ag = algp_b200.Agent.__new__(algp_b200.Agent)
    ag.env = GoldenEnv(g)
    ag.static_std, ag.mobile_std = float(g["static_std"]), float(g["mobile_std"])
    ag.static_data = unflatten(g["static_counts"], g["static_values"])
    ag.mobile_data = unflatten(g["mobile_counts"], g["mobile_values"])
    ag.criterion = 'entropy'
    ind, y, var = ag.get_sampled_dataset()
    ag.gp = make_gpr(kind, g["log_ls"], g["log_os"], g["log_noise"], ag.env.X[ind], y, var)
    ag._post_update()
    return ag


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_agent_matches_reference_golden(golden_dir, kind):
    g = load(golden_dir, "ref_agent_%s_entropy.npz" % kind)
    ag = agent_from_golden(g, kind)
    ind, y, var = ag.get_sampled_dataset()
    assert list(ind) == list(g["ds_indices"])
    np.testing.assert_array_equal(y, g["ds_y"])
    np.testing.assert_array_equal(var, g["ds_var"])
    np.testing.assert_allclose(ag.cov_matrix, g["cov_matrix"], rtol=1e-5, atol=1e-6)
    assert ag.greedy(3) == [int(v) for v in g["greedy"]]
    paths = unflatten(g["path_lens"], g["path_flat"])
    static_idx = [int(v) for v in g["static_indices"]]
    assert ag.best_path(paths, static_idx) == int(g["best_path"])
    assert ag.best_path(paths[:1], static_idx) == 0
    mu, v = ag.predict(return_var=True)
    np.testing.assert_allclose(mu, g["pred_mu"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(v, g["pred_var"], rtol=0, atol=2e-3)
    # every path utility equals the literal reference entropy (oracle run on the float64 kernel matrix)
    th = O.Theta(g["log_ls"], float(g["log_os"]), float(g["log_noise"]), kind)
    cov = O.OracleGP(th, "fp64").cov_mat(g["X"], add_likelihood_var=True)
    best, ut = O.best_path_literal(cov, g["static_sampled"], g["mobile_sampled"], ag.static_std, ag.mobile_std,
                                   paths, static_idx, return_utilities=True)
    assert best == int(g["best_path"])


@pytest.mark.parametrize("cov_mode", ["always", "auto"])
@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_agent_best_path_resident_cov_matches_reference_golden(golden_dir, kind, cov_mode):
    """Agent.best_path with the resident posterior covariance (agent.cov_mode): the reference's own choice, whether the
    covariance is built up front ("always") or by the rent-or-buy rule after repeated calls on an unchanged state
    ("auto"); a greedy step afterwards (which commits samples) still matches the reference."""
    g = load(golden_dir, "ref_agent_%s_entropy.npz" % kind)
    ag = agent_from_golden(g, kind)
    ag.cov_mode = cov_mode
    paths = unflatten(g["path_lens"], g["path_flat"])
    static_idx = [int(v) for v in g["static_indices"]]
    assert ag.best_path(paths, static_idx) == int(g["best_path"])
    state = ag._hot_state["state"]
    if cov_mode == "always":
        assert state.P is not None
    else:
        assert state.P is None                      # a few hundred candidates do not pay for the build ...
        state._stream_s = state.cov_build_seconds()
        assert ag.best_path(paths, static_idx) == int(g["best_path"])
        assert state.P is not None                  # ... but enough of them do
    assert ag.best_path(paths, static_idx) == int(g["best_path"])
    ag2 = agent_from_golden(g, kind)
    assert ag.greedy(3) == ag2.greedy(3)


def test_agent_mutual_information_matches_reference(golden_dir):
    """criterion='mutual_information' (agent.py:330-339, 388-397) against the reference's own picks and,
    utility by utility, against the literal loops."""
    g = load(golden_dir, "ref_agent_rbf_mutual_information.npz")
    ag = agent_from_golden(g, "rbf")
    ag.criterion = 'mutual_information'
    th = O.Theta(g["log_ls"], float(g["log_os"]), float(g["log_noise"]), "rbf")
    cov = O.OracleGP(th, "fp64").cov_mat(g["X"], add_likelihood_var=True)
    ss, ms = ag.static_std, ag.mobile_std
    assert ag.greedy(3) == [int(v) for v in g["greedy"]]
    paths = unflatten(g["path_lens"], g["path_flat"])
    static_idx = [int(v) for v in g["static_indices"]]
    assert ag.best_path(paths, static_idx) == int(g["best_path"])
    best, ut = O.best_path_literal(cov, g["static_sampled"], g["mobile_sampled"], ss, ms, paths, static_idx,
                                   criterion="mutual_information", return_utilities=True)
    np.testing.assert_allclose(ag._last_path_scores.cpu().numpy(), ut, rtol=1e-8, atol=1e-8)
    # first-pick utilities against the literal loop
    ag2 = agent_from_golden(g, "rbf")
    ag2.criterion = 'mutual_information'
    static, mobile = ag2._sample_flags()
    state, pi = ag2._state_for(static, mobile, capacity=8)
    ctx = engine.MIContext(state.hyper, state.X, pi)
    ent_a = state.H_base_dev + state.greedy_utilities(1 / ss ** 2)
    got = ctx.greedy_utilities(ent_a, ss, ms).cpu().numpy()
    _, u_lit = O.greedy_literal(cov, static, mobile, ss, ms, 1, criterion="mutual_information", return_utilities=True)
    fin = np.isfinite(u_lit[0])
    assert (np.isfinite(got) == fin).all()
    np.testing.assert_allclose(got[fin], u_lit[0][fin], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_mutual_information_rank1_maintenance_matches_refactorisation_and_literal_loop(kind):
    """MIContext.commit (rank-1 updates of diag(A2^-1), diag(A3^-1) and the two log-dets after a pick)
    against a context factored from scratch on the updated flags, and Agent.greedy with the MI criterion for 4 picks
    -- unsampled AND mobile-only locations get picked -- against the literal loop of agent.py:313-354, utility by utility."""
    X, yf, tr, ytr, rng = field_problem(12, 11, 40, seed=8)
    n = len(X)
    th, hy = hyper_pair([2.0, 2.5], 1.2, 0.05, kind)
    ss, ms = 0.1, 1.0
    static = np.zeros(n, bool); static[tr[:15]] = True
    mobile = np.zeros(n, bool); mobile[tr[10:]] = True
    pi = O.precisions_from_flags(static, mobile, ss, ms)
    Xd = dev(X)
    ctx = engine.MIContext(hy, Xd, pi)
    mobile_only = np.nonzero(mobile & ~static)[0]
    unsampled = np.nonzero(~(mobile | static))[0]
    picks = [int(unsampled[3]), int(mobile_only[2]), int(unsampled[40]), int(unsampled[7]), int(mobile_only[0])]
    for j in picks:
        assert ctx.commit(j, ss, ms)
        static[j] = True
        pi = O.precisions_from_flags(static, mobile, ss, ms)
        fresh = engine.MIContext(hy, Xd, pi)
        assert ctx.n_abar == fresh.n_abar
        np.testing.assert_allclose(ctx.ld2.cpu().numpy(), fresh.ld2.cpu().numpy(), rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(ctx.ld3.cpu().numpy(), fresh.ld3.cpu().numpy(), rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(ctx.diag3.cpu().numpy(), fresh.diag3.cpu().numpy(), rtol=1e-10, atol=1e-12)
        p_old, p_new = ctx.pos2.cpu().numpy(), fresh.pos2.cpu().numpy()
        assert ((p_old >= 0) == (p_new >= 0)).all()
        keep = p_new >= 0
        np.testing.assert_allclose(ctx.diag2.cpu().numpy()[p_old[keep]], fresh.diag2.cpu().numpy()[p_new[keep]], rtol=1e-10, atol=1e-12)
    # through the agent: 4 MI picks against the literal loop
    class Env(object):
        pass
    env = Env()
    env.X, env.test_X, env.num_samples = X, X[:3], n
    ag = algp_b200.Agent.__new__(algp_b200.Agent)
    ag.env, ag.static_std, ag.mobile_std, ag.criterion = env, ss, ms, 'mutual_information'
    st0 = np.zeros(n, bool); st0[tr[:15]] = True
    ag.static_data = [[1.0] if st0[i] else [] for i in range(n)]
    ag.mobile_data = [[1.0] if mobile[i] else [] for i in range(n)]
    ag.gp = make_gpr(kind, th.log_lengthscale, th.log_outputscale, th.log_noise, X[tr], ytr, np.full(len(tr), ss ** 2))
    ag._post_update()
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    want, ut = O.greedy_literal(cov, st0, mobile, ss, ms, 4, criterion="mutual_information", return_utilities=True)
    assert ag.greedy(4) == [int(p) for p in want]


def test_state_dict_roundtrip_and_unknown_kernel():
    x = np.random.default_rng(0).uniform(0, 5, (20, 2))
    gp = make_gpr("matern", np.log([1.5, 2.0]), 0.3, -2.0, x, x[:, 0], np.full(20, 0.01))
    sd = gp.model.state_dict()
    assert 'kernel_covar_module.log_outputscale' in sd and 'likelihood.log_noise' in sd      # run.py:36-37
    gp2 = algp_b200.GPR(kernel_params={'type': 'matern'})
    gp2.reset(gp.train_x, gp.train_y, gp.train_var)
    gp2.model.load_state_dict(sd)                                                            # agent.py:39-41
    assert gp2.hyper().key() == gp.hyper().key()
    with pytest.raises(NotImplementedError):
        algp_b200.GPR(kernel_params={'type': 'periodic'}).reset(x, x[:, 0], np.full(20, 0.01))


# ------------------------------------------------------------------ full-size properties
def test_fit_predict_n4096_properties():
    """BASELINE config A: N=4096 training points, 64x64 grid, fp64.  The oracle takes seconds here,
    so compare in full, then check size-independent properties."""
    rng = np.random.default_rng(1)
    grid, yf = O.gaussian_mixture_field(64, 64, seed=1)
    x = rng.uniform(0, 64, size=(4096, 2))
    y = np.zeros(4096)
    for _ in range(3):
        c = rng.uniform(0, 64, 2)
        y += np.exp(-((x - c) ** 2).sum(1) / 60.0)
    y = np.maximum(0, y + rng.normal(0, 0.1, 4096))
    var = np.full(4096, 0.01)
    th, hy = hyper_pair([4.0, 4.0], 1.0, 0.01, "rbf")
    gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, var)
    mu, v = algp_b200.predictive_distribution(gp, x, y, grid, var, return_var=True)
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, grid, var, return_var=True)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-9, atol=1e-9 * np.abs(mu_o).max())
    np.testing.assert_allclose(v, v_o, rtol=0, atol=1e-9)
    assert (v > 0).all() and (v < 1.0 + 1e-12).all()          # 0 < posterior var <= prior var s^2
    # linearity of the mean in y (same factor, cached): mean(2y) - ybar-shift == 2*mean(y) - shift
    mu2 = algp_b200.predictive_distribution(gp, x, 2 * y, grid, var)
    np.testing.assert_allclose(mu2, 2 * mu, rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ hyper-parameter learning
@pytest.mark.parametrize("kind", ["rbf", "matern"])
@pytest.mark.parametrize("n,d", [(40, 3), (300, 2), (200, 6)])
def test_mll_loss_and_grad_match_oracle(kind, n, d):
    from algp_b200.mll import mll_loss_and_grad
    rng = np.random.default_rng(n + d)
    x = rng.uniform(0, 10, size=(n, d))
    y = np.sin(x[:, 0]) + 0.1 * rng.normal(size=n)
    var = rng.uniform(0.005, 0.02, n)
    th, hy = hyper_pair(np.linspace(1.5, 3.0, d), 0.8, 0.05, kind)
    loss, g = mll_loss_and_grad(hy, x, y - y.mean(), var)
    assert loss == pytest.approx(O.mll_loss(th, x, y, var), rel=1e-10)
    g_o = O.mll_loss_grad(th, x, y, var)
    np.testing.assert_allclose(g, g_o, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_fit_follows_reference_adam_trajectory(golden_dir, kind, capsys):
    """GPR.fit = 25 Adam steps on -mll/N from theta = 0 (models.py:137-159).  The reference ran the
    same loop in float32 through the stand-in's autograd; float64 + analytic gradients stay close."""
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = algp_b200.GPR(lr=0.1, max_iterations=25, kernel_params={'type': kind})
    gp.fit(g["train_x"], g["train_y"], g["train_var"])
    out = capsys.readouterr().out
    assert "Initial LogLikelihood" in out and "Final LogLikelihood" in out
    hy = gp.hyper()
    np.testing.assert_allclose(hy.log_ls, g["fit_log_ls"], rtol=0, atol=0.05)
    assert hy.log_os == pytest.approx(float(g["fit_log_os"]), abs=0.05)
    assert hy.log_noise == pytest.approx(float(g["fit_log_noise"]), abs=0.05)
    loss, _ = gp.loss_and_grad()
    assert loss == pytest.approx(float(g["fit_loss_at_theta"]), abs=0.02)
    # fit with var=None uses the reference's default 1e-5 (models.py:138-139) and must still factor
    gp2 = algp_b200.GPR(lr=0.1, max_iterations=3, kernel_params={'type': kind})
    gp2.fit(g["train_x"], g["train_y"])
    assert np.allclose(gp2.train_var, 1e-5)


# ------------------------------------------------------------------ episode (BASELINE configs[4])
@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_episode_matches_literal_reference_loops(kind):
    """Greedy picks + best path + commits over several batches, against the oracle's literal
    agent.py loops run on the same flags (small field so the CPU side takes seconds)."""
    from algp_b200.episode import run_episode
    X, y, tr, ytr, rng = field_problem(13, 12, 30, seed=4)
    n = len(X)
    th, hy = hyper_pair([2.0, 2.5], 1.0, 0.03, kind)
    ss, ms = 0.1, 1.0
    static = np.zeros(n, bool)
    mobile = np.zeros(n, bool)
    static[tr[:18]] = True
    mobile[tr[12:]] = True                      # some locations carry both readings
    pi0 = O.precisions_from_flags(static, mobile, ss, ms)
    batches, per_batch = 3, 2
    prng = np.random.default_rng(9)
    all_paths = [np.stack([prng.choice(n, 7, replace=False) for _ in range(12)]).astype(np.int32) for _ in range(batches)]
    for b in range(batches):
        all_paths[b][0, 5] = all_paths[b][0, 1]          # a repeat inside a path
        all_paths[b][3, 6] = -1                          # ragged path
    res = run_episode(hy, dev(X), static, mobile, ss, ms, batches, per_batch,
                      lambda b, picks: all_paths[b], return_scores=True)

    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    st, mo = static.copy(), mobile.copy()
    for b in range(batches):
        picks = O.greedy_literal(cov, st, mo, ss, ms, per_batch)
        assert res["picks"][b] == [int(p) for p in picks]
        paths = [[int(v) for v in row if v >= 0] for row in all_paths[b]]
        best, ut = O.best_path_literal(cov, st, mo, ss, ms, paths, picks, return_utilities=True)
        assert res["best_paths"][b] == best
        assert res["scores"][b] == pytest.approx(float(ut[best]), rel=1e-8)
        st[picks] = True
        mo[paths[best]] = True
    # the device state now equals a fresh factorisation of the final flags
    pi_fin = O.precisions_from_flags(st, mo, ss, ms)
    ost = O.posterior_state(cov, pi_fin)
    np.testing.assert_allclose(res["state"].diagP.cpu().numpy(), np.diag(ost["P"]), rtol=0, atol=1e-9)
    np.testing.assert_allclose(res["state"].pi.cpu().numpy(), pi_fin, rtol=1e-12)


# ------------------------------------------------------------------ growing prefixes (agent.py:497-518)
@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_prefix_posteriors_match_per_prefix_solves(kind):
    from algp_b200.utils import predictive_distribution_prefixes
    X, yf, tr, ytr, rng = field_problem(20, 18, 200, seed=6)
    order = rng.integers(0, len(X), 230)                 # with repeats: a location read twice (static, then mobile)
    x = X[order]
    y = np.maximum(0, yf[order] + rng.normal(0, 0.1, len(order)))
    var = np.where(rng.random(len(order)) < 0.6, 0.01, 1.0)
    te = rng.choice(len(X), 37, replace=False)
    tvar = np.full(37, 0.02)
    th, hy = hyper_pair([2.5, 3.0], 1.1, 0.03, kind)
    gp = make_gpr(kind, th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, var)
    ogp = O.OracleGP(th, "fp64")
    counts = [10, 25, 26, 130, 230]
    res = predictive_distribution_prefixes(gp, x, y, X[te], var, counts, test_var=tvar, return_cov=True, return_mi=True)
    resv = predictive_distribution_prefixes(gp, x, y, X[te], var, counts, return_var=True)
    resm = predictive_distribution_prefixes(gp, x, y, X[te], var, counts)
    for i, c in enumerate(counts):
        mu_o, cov_o, mi_o = O.predictive_distribution_chol(ogp, x[:c], y[:c], X[te], var[:c], test_var=tvar,
                                                           return_cov=True, return_mi=True)
        mu, cov, mi = res[i]
        np.testing.assert_allclose(mu, mu_o, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(cov, cov_o, rtol=0, atol=1e-9 * 1.1)
        assert mi == pytest.approx(mi_o, rel=1e-8)
        _, var_o = O.predictive_distribution_chol(ogp, x[:c], y[:c], X[te], var[:c], return_var=True)
        np.testing.assert_allclose(resv[i][1], var_o, rtol=0, atol=1e-9 * 1.1)
        np.testing.assert_allclose(resv[i][0], mu_o, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(resm[i], mu_o, rtol=1e-9, atol=1e-9)


def test_prediction_vs_distance_matches_literal_loop():
    X, yf, tr, ytr, rng = field_problem(15, 14, 50, seed=8)
    te = rng.choice(len(X), 20, replace=False)
    ag = algp_b200.Agent.__new__(algp_b200.Agent)
    ag.env = GoldenEnv(dict(X=X, test_X=X[te]))
    ag.env.test_Y = yf[te]
    ag.static_std, ag.mobile_std, ag.criterion = 0.1, 1.0, 'entropy'
    n_read = 60
    inds = list(rng.integers(0, len(X), n_read))
    inds[7] = -1                                          # an invalid reading is skipped (agent.py:504-505)
    stds = [0.1 if rng.random() < 0.5 else 1.0 for _ in range(n_read)]
    ys = [None if i == -1 else float(max(0, yf[i] + rng.normal(0, s))) for i, s in zip(inds, stds)]
    ag.collected = {'ind': inds, 'std': stds, 'y': ys}
    th, hy = hyper_pair([2.0, 2.5], 1.0, 0.05, "rbf")
    ag.gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, X[:2], yf[:2], np.full(2, 0.01))
    got = ag.prediction_vs_distance(test_every=12, num_runs=5)
    ogp = O.OracleGP(th, "fp64")
    count = 0
    err, mis, mv = [], [], []
    while count < 60:                                     # the reference loop, agent.py:503-516
        count += 12
        ii = np.array(inds[:count])
        valid = ii != -1
        x = X[ii[valid]]
        var = np.array(stds)[:count][valid] ** 2
        y = np.array([v for v in ys[:count] if v is not None])
        mu, cov, mi = O.predictive_distribution_chol(ogp, x, y, X[te], var, return_mi=True, return_cov=True)
        err.append(np.mean(np.abs(yf[te] - mu)))
        mis.append(mi)
        mv.append(np.diag(cov).mean())
    np.testing.assert_allclose(got['error'], err, rtol=1e-8)
    np.testing.assert_allclose(got['mi'], mis, rtol=1e-8)
    np.testing.assert_allclose(got['mean_var'], mv, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(got['mean'], mu, rtol=1e-9, atol=1e-9)


def test_prediction_vs_distance_degenerate_prefixes():
    """Prefixes that hold no valid reading yet (only -1 gaps), and a request that runs past the end of the collected
    list (the reference then repeats the last prefix): no assertion, no IndexError, reference-shaped results."""
    X, yf, tr, ytr, rng = field_problem(12, 11, 30, seed=3)
    te = rng.choice(len(X), 9, replace=False)
    ag = algp_b200.Agent.__new__(algp_b200.Agent)
    ag.env = GoldenEnv(dict(X=X, test_X=X[te]))
    ag.env.test_Y = yf[te]
    ag.static_std, ag.mobile_std, ag.criterion = 0.1, 1.0, 'entropy'
    inds = [-1, -1, -1, 4, 9, -1, 17, 30]
    stds = [0.1] * len(inds)
    ys = [None if i == -1 else float(yf[i]) for i in inds]
    ag.collected = {'ind': inds, 'std': stds, 'y': ys}
    th, hy = hyper_pair([2.0, 2.5], 1.0, 0.05, "rbf")
    ag.gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, X[:2], yf[:2], np.full(2, 0.01))
    got = ag.prediction_vs_distance(test_every=3, num_runs=4)            # prefixes of 3, 6, 9 (-> 8), 12 (-> 8) entries
    ogp = O.OracleGP(th, "fp64")
    assert np.isnan(got['error'][0]) and got['mi'][0] == 0.0             # nothing read yet: NaN mean, prior covariance
    assert got['mean_var'][0] == pytest.approx(np.exp(th.log_outputscale), rel=1e-12)
    for slot, count in ((1, 6), (2, 8), (3, 8)):
        ii = np.array(inds[:count])
        valid = ii != -1
        y = np.array([v for v in ys[:count] if v is not None])
        mu, cov, mi = O.predictive_distribution_chol(ogp, X[ii[valid]], y, X[te], np.array(stds)[:count][valid] ** 2,
                                                     return_mi=True, return_cov=True)
        assert got['error'][slot] == pytest.approx(np.mean(np.abs(yf[te] - mu)), rel=1e-8)
        assert got['mi'][slot] == pytest.approx(mi, rel=1e-8)
        assert got['mean_var'][slot] == pytest.approx(np.diag(cov).mean(), rel=1e-8)
    ag.collected = {'ind': [], 'std': [], 'y': []}
    empty = ag.prediction_vs_distance(test_every=2, num_runs=2)
    assert len(empty['error']) == 2 and all(np.isnan(e) for e in empty['error'])


def test_i8fast_precision_meets_the_1e4_tier(monkeypatch):
    """gp.precision = "i8fast": the digit path with 5 planes in the factorisation and 4 in the variance product -- the
    fp32 / TF32 tier of the north star (rel 1e-4 on mean and variance), against the fp64 oracle; the digit factorisation
    and the Z-order staging are forced on at this size."""
    rng = np.random.default_rng(32)
    monkeypatch.setattr(engine, "I8_REORDER_MIN", 2048)
    monkeypatch.setattr(engine, "I8_FACTOR_MIN", 2048)
    N, M = 2304, 1500
    x = rng.uniform(0, 120, (N, 2))
    xs = rng.uniform(-5, 125, (M, 2))
    y = np.sin(x[:, 0] / 7.0) + np.cos(x[:, 1] / 5.0) + rng.normal(0, 0.1, N)
    var = rng.uniform(0.01, 0.02, N)
    th, hy = hyper_pair([6.0, 5.0], 1.0, 0.01, "rbf")
    gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, var)
    gp.precision = "i8fast"
    mu, v = algp_b200.predictive_distribution(gp, x, y, xs, var, return_var=True)
    assert gp._cache["factor"].perm is not None
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, return_var=True)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-4, atol=1e-4 * np.abs(mu_o).max())
    np.testing.assert_allclose(v, v_o, rtol=1e-4, atol=1e-4)          # s^2 = 1: absolute 1e-4 of the prior variance
    with pytest.raises(ValueError):
        gp.precision = "i4"
        algp_b200.predictive_distribution(gp, x, y, xs, var, return_var=True)


def test_i8_precision_reorders_points_and_matches_fp64(monkeypatch):
    """gp.precision = "i8" from N = engine.I8_REORDER_MIN (lowered here): training and test points are sorted along a
    Z curve internally (so that the
    digit GEMM can skip far-apart tiles); the caller sees the same mean / variance, in the caller's order."""
    rng = np.random.default_rng(31)
    monkeypatch.setattr(engine, "I8_REORDER_MIN", 2048)
    N, M = 2304, 1500
    x = rng.uniform(0, 120, (N, 2))
    xs = rng.uniform(-5, 125, (M, 2))                      # some test points outside the training box
    y = np.sin(x[:, 0] / 7.0) + np.cos(x[:, 1] / 5.0) + rng.normal(0, 0.1, N)
    var = rng.uniform(0.01, 0.02, N)
    tvar = rng.uniform(0.0, 0.01, M)
    th, hy = hyper_pair([6.0, 5.0], 1.0, 0.01, "rbf")
    gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, var)
    mu64, v64 = algp_b200.predictive_distribution(gp, x, y, xs, var, tvar, return_var=True)
    gp.precision = "i8"
    mu8, v8 = algp_b200.predictive_distribution(gp, x, y, xs, var, tvar, return_var=True)
    f = gp._cache["factor"]
    assert f.perm is not None and sorted(f.perm.cpu().tolist()) == list(range(N))
    np.testing.assert_allclose(mu8, mu64, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v8, v64, rtol=0, atol=1e-9)
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, tvar, return_var=True)
    np.testing.assert_allclose(mu8, mu_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v8, v_o, rtol=0, atol=1e-9)
    # a covariance request must not see the reordered factor
    mu_c, cov = algp_b200.predictive_distribution(gp, x, y, xs[:64], var, tvar[:64], return_cov=True)
    np.testing.assert_allclose(np.diag(cov), v64[:64], rtol=0, atol=1e-9)


@pytest.mark.parametrize("precision", ["fp64", "i8"])
def test_sharded_mean_var_single_rank_equals_posterior(precision, monkeypatch):
    """dist.sharded_mean_var with one rank is the plain posterior path (same factor, same row order)."""
    from algp_b200 import dist as adist
    monkeypatch.setattr(engine, "I8_REORDER_MIN", 1024)
    rng = np.random.default_rng(41)
    N, M = 1300, 777
    x = rng.uniform(0, 90, (N, 2))
    xs = rng.uniform(0, 90, (M, 2))
    y = np.sin(x[:, 0] / 6.0) + rng.normal(0, 0.1, N)
    var = rng.uniform(0.01, 0.02, N)
    th, hy = hyper_pair([5.0, 5.0], 1.0, 0.01, "rbf")
    dev_ = lambda a: engine.to_dev(np.asarray(a, dtype=np.float64))
    mu, v = adist.sharded_mean_var(hy, dev_(x), dev_(var), dev_(y - y.mean()), float(y.mean()), dev_(xs), precision=precision)
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, None, return_var=True)
    np.testing.assert_allclose(mu.cpu().numpy(), mu_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), v_o, rtol=0, atol=1e-9)


def test_tracing_reports_device_time_per_entry_point_and_changes_nothing():
    """algp_b200.tracing (the per-phase timers of run_ipp, agent.py:138-219, per kernel family): the traced call gives the
    same numbers as the plain one, every C-ABI entry point it went through is listed with its call count and a
    positive device time, and nothing is recorded once tracing is off."""
    from algp_b200 import _lib, tracing
    X, y, tr, ytr, rng = field_problem(24, 20, 300, seed=5)
    te = rng.choice(len(X), 150, replace=False)
    var = np.full(300, 0.01)
    th, hy = hyper_pair([2.5, 3.5], 1.3, 0.02, "rbf")
    gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, X[tr], ytr, var)
    plain = algp_b200.predictive_distribution(gp, X[tr], ytr, X[te], var, return_var=True)
    gp._cache.clear()
    with tracing.trace() as t:
        traced = algp_b200.predictive_distribution(gp, X[tr], ytr, X[te], var, return_var=True)
        launches = _lib.launch_count
    np.testing.assert_array_equal(plain[0], traced[0])
    np.testing.assert_array_equal(plain[1], traced[1])
    s = t.summary()
    for name in ("algp_kbuild", "algp_potrf", "algp_trtri"):
        assert s[name]["calls"] >= 1 and s[name]["ms"] > 0.0, (name, s)
    assert list(s) == sorted(s, key=lambda k: -s[k]["ms"])          # slowest first
    assert _lib._trace is None
    gp._cache.clear()
    algp_b200.predictive_distribution(gp, X[tr], ytr, X[te], var, return_var=True)
    assert _lib.launch_count > launches and t.summary() == s          # untraced calls leave the tracer alone
