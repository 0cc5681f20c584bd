"""Freeze golden vectors by running the REFERENCE's own code in this container.

    python tests/golden/make_golden.py        # writes tests/golden/ref_*.npz

Imports /root/reference/{utils,models,agent}.py unmodified (read-only), with
tests/golden/gpytorch_standin.py standing in for the missing 2018 gpytorch and
inert stubs for ipdb / seaborn / matplotlib.  Everything the reference computes
in-tree -- GPR.cov_mat's dtype flow and noise handling, predictive_distribution,
entropy_from_cov, Agent.get_sampled_dataset / greedy / best_path -- therefore
executes literally; only the kernel closed forms and the MLL come from the
stand-in.  /root/reference does not exist on the GPU box, so the outputs are
committed as small fixtures and this script is the record of how they were made.
The reference never seeds its RNGs (arguments.py:33 is parsed, never applied);
seeds are fixed here.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpytorch_standin  # noqa: E402

gpytorch_standin.install()
sys.path.insert(0, "/root/reference")
import utils as ref_utils      # noqa: E402
import models as ref_models    # noqa: E402
import agent as ref_agent      # noqa: E402


def mixture_field(rows, cols, seed):
    np.random.seed(seed)
    grid, y = ref_utils.generate_gaussian_data(rows, cols)     # utils.py:90-108
    return grid.astype(np.float64), y


class FakeEnv(object):
    """The attributes Agent's hot path reads from FieldEnv (env.py:58-113)."""

    def __init__(self, X, Y, test_X):
        self.X = X
        self.Y = Y
        self.test_X = test_X
        self.num_samples = len(X)
        self.rng = np.random.RandomState(7)

    def collect_samples(self, idx, std):
        # env.py:108-113: noisy reading clipped at 0
        return max(0.0, float(self.Y[idx] + self.rng.normal(0, std)))


def make_args(kernel, max_iterations=25):
    return types.SimpleNamespace(kernel=kernel, latent=None, lr=0.1, max_iterations=max_iterations,
                                 static_std=0.1, num_samples_per_batch=4, update_every=1,
                                 fraction_pretrain=0.3)


def set_theta(gp, log_ls, log_os, log_noise):
    m = gp.model
    with torch.no_grad():
        m.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(log_ls, dtype=torch.float32).view(1, 1, -1))
        m.kernel_covar_module.log_outputscale.fill_(log_os)
        gp.likelihood.log_noise.fill_(log_noise)


def get_theta(gp):
    m = gp.model
    return (m.kernel_covar_module.base_kernel.log_lengthscale.detach().numpy().reshape(-1).astype(np.float64),
            float(m.kernel_covar_module.log_outputscale.item()), float(gp.likelihood.log_noise.item()))


def case_gp(kernel, seed):
    """cov_mat / predictive_distribution / entropy / fit on a 2-D field."""
    rows, cols = 14, 12
    grid, y = mixture_field(rows, cols, seed)
    rng = np.random.RandomState(seed + 100)
    perm = rng.permutation(len(grid))
    tr, te = perm[:90], perm[90:130]
    train_x, test_x = grid[tr], grid[te]
    train_y = np.maximum(0, y[tr] + rng.normal(0, 0.1, len(tr)))
    train_var = np.where(rng.rand(len(tr)) < 0.5, 0.01, 1.0 / (1 / 0.01 + 1 / 1.0))
    test_var = np.full(len(te), 0.02)

    out = dict(train_x=train_x, train_y=train_y, train_var=train_var, test_x=test_x, test_var=test_var)
    torch.manual_seed(seed)
    gp = ref_models.GPR(lr=0.1, max_iterations=25, kernel_params={'type': kernel})
    gp.fit(train_x, train_y, train_var)                    # models.py:137-159 (stand-in MLL)
    ls, os_, nz = get_theta(gp)
    out.update(fit_log_ls=ls, fit_log_os=os_, fit_log_noise=nz)
    # loss at the fitted theta, exactly as fit() evaluates it (models.py:147-148)
    gp.model.train(); gp.likelihood.train()
    with torch.no_grad():
        out["fit_loss_at_theta"] = float(-gp.mll(gp.model(gp._train_x), gp._zero_mean_train_y).item())
    out["predict_mean"], out["predict_var"] = gp.predict(test_x, return_std=True)      # models.py:183-197 (un-pinned)

    # fixed, well-conditioned theta for the parity tiers
    log_ls = np.log(np.array([2.5, 3.5])); log_os = float(np.log(1.3)); log_noise = float(np.log(0.02))
    set_theta(gp, log_ls, log_os, log_noise)
    out.update(log_ls=log_ls, log_os=log_os, log_noise=log_noise)
    out["K_train"] = gp.cov_mat(train_x)
    out["K_train_noise"] = gp.cov_mat(x1=train_x, white_noise_var=train_var, add_likelihood_var=True)
    out["K_test_train"] = gp.cov_mat(x1=test_x, x2=train_x)
    out["K_test_wn"] = gp.cov_mat(x1=test_x, white_noise_var=test_var)
    out["K_same_x2"] = gp.cov_mat(x1=train_x[:20], x2=train_x[:20].copy(), add_likelihood_var=True)
    pd = ref_utils.predictive_distribution
    out["pd_mu"] = pd(gp, train_x, train_y, test_x, train_var)
    out["pd_mu_v"], out["pd_var"] = pd(gp, train_x, train_y, test_x, train_var, return_var=True)
    _, out["pd_cov"] = pd(gp, train_x, train_y, test_x, train_var, return_cov=True)
    _, out["pd_mi"] = pd(gp, train_x, train_y, test_x, train_var, test_var=test_var, return_mi=True)
    _, out["pd_cov_tv"], out["pd_mi2"] = pd(gp, train_x, train_y, test_x, train_var, test_var=test_var,
                                            return_cov=True, return_mi=True)
    out["ent_K_train_noise"] = ref_utils.entropy_from_cov(out["K_train_noise"])
    out["ent_pd_cov_tv"] = ref_utils.entropy_from_cov(out["pd_cov_tv"])
    out["CONST"] = ref_utils.CONST
    return out


def case_agent(kernel, criterion, seed):
    """Agent.get_sampled_dataset / _post_update / greedy / best_path / predict."""
    rows, cols = 10, 9
    grid, y = mixture_field(rows, cols, seed)
    rng = np.random.RandomState(seed + 200)
    perm = rng.permutation(len(grid))
    te = perm[:12]
    tr = np.sort(perm[12:])
    X, Y, test_X = grid[tr], y[tr], grid[te]
    env = FakeEnv(X, Y, test_X)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ag = ref_agent.Agent(env, make_args(kernel, max_iterations=3))          # agent.py:13-32 (pilot survey + fit)
    # add mobile readings: some on fresh locations, some on top of static ones
    mob = rng.permutation(env.num_samples)[:15]
    ag._add_samples(list(mob), stds=[ag.mobile_std] * len(mob))             # agent.py:64-82
    log_ls = np.log(np.array([2.0, 3.0])); log_os = float(np.log(0.9)); log_noise = float(np.log(0.05))
    set_theta(ag.gp, log_ls, log_os, log_noise)
    ag.criterion = criterion
    ag._post_update()                                                        # agent.py:89-90

    out = dict(X=X, Y=Y, test_X=test_X, log_ls=log_ls, log_os=log_os, log_noise=log_noise,
               static_std=ag.static_std, mobile_std=ag.mobile_std,
               static_sampled=np.array([len(v) > 0 for v in ag.static_data]),
               mobile_sampled=np.array([len(v) > 0 for v in ag.mobile_data]),
               cov_matrix=ag.cov_matrix)
    # ragged reading lists, flattened for npz
    out["static_counts"] = np.array([len(v) for v in ag.static_data])
    out["static_values"] = np.array([v for lst in ag.static_data for v in lst])
    out["mobile_counts"] = np.array([len(v) for v in ag.mobile_data])
    out["mobile_values"] = np.array([v for lst in ag.mobile_data for v in lst])
    ind, yy, vv = ag.get_sampled_dataset()                                   # agent.py:92-117
    out.update(ds_indices=np.array(ind), ds_y=yy, ds_var=vv)
    out["greedy"] = np.array(ag.greedy(3))                                   # agent.py:295-356
    static_idx = list(out["greedy"][:2])
    n = env.num_samples
    paths = []
    for p in range(9):
        L = rng.randint(3, 9)
        path = list(rng.permutation(n)[:L])
        if p % 3 == 0:
            path = path + path[:2]                   # duplicates are idempotent (agent.py:377)
        if p % 4 == 1:
            path[0] = int(static_idx[0])             # mobile reading on a new static waypoint
        paths.append([int(v) for v in path])
    out["path_lens"] = np.array([len(p) for p in paths])
    out["path_flat"] = np.array([v for p in paths for v in p])
    out["static_indices"] = np.array(static_idx)
    out["best_path"] = int(ag.best_path(paths, static_idx))                  # agent.py:358-403
    out["best_path_single"] = int(ag.best_path(paths[:1], static_idx))       # agent.py:362-363
    mu, var = ag.predict(return_var=True)                                    # agent.py:289-293
    out.update(pred_mu=mu, pred_var=var)
    return out


def main():
    np.set_printoptions(precision=6, suppress=True)
    for kernel in ("rbf", "matern"):
        d = case_gp(kernel, seed=11)
        np.savez_compressed(os.path.join(HERE, "ref_gp_%s.npz" % kernel), **d)
        print("gp", kernel, "pd_mi", d["pd_mi"], "fit theta", d["fit_log_ls"], d["fit_log_os"], d["fit_log_noise"])
    for kernel, crit in (("rbf", "entropy"), ("matern", "entropy"), ("rbf", "mutual_information")):
        d = case_agent(kernel, crit, seed=5)
        np.savez_compressed(os.path.join(HERE, "ref_agent_%s_%s.npz" % (kernel, crit)), **d)
        print("agent", kernel, crit, "greedy", d["greedy"], "best_path", d["best_path"])


if __name__ == "__main__":
    main()
