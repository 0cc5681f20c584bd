"""Pin the oracle (oracle/oracle.py) against outputs of the REFERENCE's own
code frozen by tests/golden/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest

import oracle as O

KERNELS = ["rbf", "matern"]


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def theta_of(g, kind):
    return O.Theta(g["log_ls"], float(g["log_os"]), float(g["log_noise"]), kind)


@pytest.mark.parametrize("kind", KERNELS)
def test_cov_mat_ref32_matches_reference(golden_dir, kind):
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = O.OracleGP(theta_of(g, kind), "ref32")
    # float32 kernel entries: same closed form, different libm -> a few ulp of float32
    tol = dict(rtol=2e-6, atol=2e-7)
    K = gp.cov_mat(g["train_x"])
    assert K.dtype == np.float32 and g["K_train"].dtype == np.float32     # utils.py:19 dtype flow
    np.testing.assert_allclose(K, g["K_train"], **tol)
    np.testing.assert_allclose(gp.cov_mat(g["train_x"], white_noise_var=g["train_var"], add_likelihood_var=True),
                               g["K_train_noise"], **tol)
    np.testing.assert_allclose(gp.cov_mat(g["test_x"], g["train_x"]), g["K_test_train"], **tol)
    np.testing.assert_allclose(gp.cov_mat(g["test_x"], white_noise_var=g["test_var"]), g["K_test_wn"], **tol)
    np.testing.assert_allclose(gp.cov_mat(g["train_x"][:20], g["train_x"][:20].copy(), add_likelihood_var=True),
                               g["K_same_x2"], **tol)
    assert float(g["CONST"]) == O.CONST


@pytest.mark.parametrize("kind", KERNELS)
def test_cov_mat_fp64_close_to_reference(golden_dir, kind):
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = O.OracleGP(theta_of(g, kind), "fp64")
    np.testing.assert_allclose(gp.cov_mat(g["train_x"], white_noise_var=g["train_var"], add_likelihood_var=True),
                               g["K_train_noise"].astype(np.float64), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", KERNELS)
def test_entropy_from_cov_exact_on_reference_matrices(golden_dir, kind):
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    # same LAPACK slogdet on the same stored matrices -> tight
    assert O.entropy_from_cov(g["K_train_noise"]) == pytest.approx(float(g["ent_K_train_noise"]), rel=1e-12)
    assert O.entropy_from_cov(g["pd_cov_tv"]) == pytest.approx(float(g["ent_pd_cov_tv"]), rel=1e-12)


@pytest.mark.parametrize("kind", KERNELS)
def test_predictive_distribution_ref32(golden_dir, kind):
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = O.OracleGP(theta_of(g, kind), "ref32")
    a = (gp, g["train_x"], g["train_y"], g["test_x"], g["train_var"])
    # float32 inverse of a cond ~1e2..1e3 matrix: agree to ~1e-3 of the prior scale
    mu = O.predictive_distribution(*a)
    np.testing.assert_allclose(mu, g["pd_mu"], rtol=0, atol=2e-3)
    mu2, var = O.predictive_distribution(*a, return_var=True)
    np.testing.assert_allclose(var, g["pd_var"], rtol=0, atol=2e-3)
    _, cov = O.predictive_distribution(*a, return_cov=True)
    np.testing.assert_allclose(cov, g["pd_cov"], rtol=0, atol=2e-3)
    _, mi = O.predictive_distribution(*a, test_var=g["test_var"], return_mi=True)
    assert mi == pytest.approx(float(g["pd_mi"]), rel=2e-3)
    mu3, cov3, mi3 = O.predictive_distribution(*a, test_var=g["test_var"], return_cov=True, return_mi=True)
    assert mi3 == pytest.approx(float(g["pd_mi2"]), rel=2e-3)
    assert cov3.shape == g["pd_cov_tv"].shape


@pytest.mark.parametrize("kind", KERNELS)
def test_predictive_distribution_fp64_chol_vs_reference(golden_dir, kind):
    """The fp64 truth agrees with the float32 reference at the 1e-4 tier
    (relative to the prior scale s^2)."""
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    gp = O.OracleGP(theta_of(g, kind), "fp64")
    a = (gp, g["train_x"], g["train_y"], g["test_x"], g["train_var"])
    s2 = float(np.exp(g["log_os"]))
    mu, var = O.predictive_distribution_chol(*a, return_var=True)
    np.testing.assert_allclose(mu, g["pd_mu"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(var, g["pd_var"], rtol=0, atol=2e-3 * s2)
    # and with the literal explicit-inverse formula in float64
    mu_l, var_l = O.predictive_distribution(*a, return_var=True)
    np.testing.assert_allclose(mu, mu_l, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(var, var_l, rtol=0, atol=1e-9 * s2)
    _, mi = O.predictive_distribution_chol(*a, test_var=g["test_var"], return_mi=True)
    assert mi == pytest.approx(float(g["pd_mi"]), rel=2e-3)


@pytest.mark.parametrize("kind", KERNELS)
def test_mll_loss_matches_reference_fit_objective(golden_dir, kind):
    g = load(golden_dir, "ref_gp_%s.npz" % kind)
    th = O.Theta(g["fit_log_ls"], float(g["fit_log_os"]), float(g["fit_log_noise"]), kind)
    loss = O.mll_loss(th, g["train_x"], g["train_y"], g["train_var"])
    assert loss == pytest.approx(float(g["fit_loss_at_theta"]), rel=2e-3, abs=2e-3)


AGENT_CASES = [("rbf", "entropy"), ("matern", "entropy"), ("rbf", "mutual_information")]


def unflatten(counts, values):
    out, p = [], 0
    for c in counts:
        out.append(list(values[p:p + c]))
        p += c
    return out


@pytest.mark.parametrize("kind,crit", AGENT_CASES)
def test_agent_hot_path_matches_reference(golden_dir, kind, crit):
    g = load(golden_dir, "ref_agent_%s_%s.npz" % (kind, crit))
    st = unflatten(g["static_counts"], g["static_values"])
    mo = unflatten(g["mobile_counts"], g["mobile_values"])
    ss, ms = float(g["static_std"]), float(g["mobile_std"])
    ind, y, var = O.get_sampled_dataset(st, mo, ss, ms)
    assert list(ind) == list(g["ds_indices"])
    np.testing.assert_array_equal(y, g["ds_y"])
    np.testing.assert_array_equal(var, g["ds_var"])

    gp = O.OracleGP(theta_of(g, kind), "ref32")
    cov = gp.cov_mat(g["X"], add_likelihood_var=True)               # agent.py:89-90
    np.testing.assert_allclose(cov, g["cov_matrix"], rtol=2e-6, atol=2e-7)

    # on the reference's own cov_matrix the literal loops must reproduce the picks exactly
    picks = O.greedy_literal(g["cov_matrix"], g["static_sampled"], g["mobile_sampled"], ss, ms, 3, criterion=crit)
    assert [int(p) for p in picks] == [int(p) for p in g["greedy"]]
    paths = unflatten(g["path_lens"], g["path_flat"])
    bp = O.best_path_literal(g["cov_matrix"], g["static_sampled"], g["mobile_sampled"], ss, ms,
                             paths, list(g["static_indices"]), criterion=crit)
    assert bp == int(g["best_path"])
    assert O.best_path_literal(g["cov_matrix"], g["static_sampled"], g["mobile_sampled"], ss, ms,
                               paths[:1], list(g["static_indices"]), criterion=crit) == int(g["best_path_single"]) == 0
    # and from the oracle's own kernel matrix (float32 libm differences only)
    picks2 = O.greedy_literal(cov, g["static_sampled"], g["mobile_sampled"], ss, ms, 3, criterion=crit)
    assert [int(p) for p in picks2] == [int(p) for p in g["greedy"]]

    mu, v = O.predictive_distribution(gp, g["X"][ind], y, g["test_X"], var, return_var=True)   # agent.py:289-293
    np.testing.assert_allclose(mu, g["pred_mu"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(v, g["pred_var"], rtol=0, atol=2e-3)
