"""Edge cases through the reference-facing API: tiny / ragged sizes, 1-D inputs, empty and degenerate
candidate sets (the reference tests none of these; its NumPy code defines the expected behaviour)."""
import numpy as np
import pytest
import torch

import oracle as O
import algp_b200
from algp_b200 import engine
from gpu_helpers import dev, hyper_pair
from test_gpu_api import make_gpr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,M", [(1, 1), (2, 5), (127, 1), (129, 3)])
def test_tiny_and_ragged_sizes(N, M):
    rng = np.random.default_rng(N * 10 + M)
    x, xs = rng.uniform(0, 5, (N, 2)), rng.uniform(0, 5, (M, 2))
    y = rng.normal(size=N)
    var = np.full(N, 0.05)
    th, hy = hyper_pair([1.5, 2.0], 0.7, 0.1, "matern")
    gp = make_gpr("matern", th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, var)
    mu, v = algp_b200.predictive_distribution(gp, x, y, xs, var, return_var=True)
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, return_var=True)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(v, v_o, rtol=0, atol=1e-10)
    _, cov, mi = algp_b200.predictive_distribution(gp, x, y, xs, var, test_var=np.full(M, 0.01), return_cov=True, return_mi=True)
    _, cov_o, mi_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, test_var=np.full(M, 0.01),
                                                     return_cov=True, return_mi=True)
    np.testing.assert_allclose(cov, cov_o, rtol=0, atol=1e-10)
    assert mi == pytest.approx(mi_o, rel=1e-8, abs=1e-10)


def test_one_dimensional_inputs_and_no_train_var():
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, 10, 50))                 # 1-D array, like a 1-D field
    y = np.sin(x)
    xs = np.linspace(0, 10, 17)
    th, hy = hyper_pair([1.2], 1.0, 0.01, "rbf")
    gp = make_gpr("rbf", th.log_lengthscale, th.log_outputscale, th.log_noise, x, y, np.full(50, 1e-5))
    mu, v = algp_b200.predictive_distribution(gp, x, y, xs, None, return_var=True)          # train_var=None (utils.py:293)
    mu_o, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, None, return_var=True)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v, v_o, rtol=0, atol=1e-9)
    K = gp.cov_mat(x, xs)
    assert K.shape == (50, 17)
    np.testing.assert_allclose(K, O.kernel_matrix(th, x, xs), rtol=1e-12, atol=1e-14)
    assert gp.get_embeddings(x).shape == (50, 1)


def test_degenerate_candidate_sets():
    rng = np.random.default_rng(2)
    X = rng.uniform(0, 10, (150, 2))
    th, hy = hyper_pair([2.0, 2.0], 1.0, 0.05, "rbf")
    pi0 = np.zeros(150)
    pi0[:40] = 100.0
    pi0[30:60] += 1.0
    state = engine.PosteriorState(hy, dev(X), np.nonzero(pi0 > 0)[0], pi0)
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    ost = O.posterior_state(cov, pi0)
    # all slots empty -> H(B); zero increments -> H(B); one candidate only
    idx = np.array([[-1, -1, -1], [5, 70, 71], [70, 70, 70]], dtype=np.int32)
    delta = np.array([[1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    got = state.score_sets(dev(idx, torch.int32), dev(delta)).cpu().numpy()
    want = O.score_sets_restructured(ost["P"], pi0, idx, delta, ost["H"])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-10)
    assert got[0] == pytest.approx(ost["H"]) and got[1] == pytest.approx(ost["H"])
    one = state.score_sets(dev(idx[2:3], torch.int32), dev(delta[2:3])).cpu().numpy()
    assert one[0] == pytest.approx(want[2], rel=1e-10)
    # 20-slot candidates (generic kernel) that are mostly empty
    idx20 = -np.ones((5, 20), dtype=np.int32)
    idx20[:, 3] = [61, 62, 63, 64, 65]
    idx20[:, 17] = [100, 101, 102, 103, 104]
    d20 = np.ones((5, 20))
    got20 = state.score_sets(dev(idx20, torch.int32), dev(d20)).cpu().numpy()
    np.testing.assert_allclose(got20, O.score_sets_restructured(ost["P"], pi0, idx20, d20, ost["H"]), rtol=1e-9, atol=1e-10)


class _Env(object):
    def __init__(self, X):
        self.X, self.test_X, self.num_samples = X, X[:3], len(X)


def _agent(X, static_idx, mobile_idx, kind="rbf"):
    ag = algp_b200.Agent.__new__(algp_b200.Agent)
    ag.env = _Env(X)
    ag.static_std, ag.mobile_std, ag.criterion = 0.1, 1.0, 'entropy'
    ag.static_data = [[1.0] if i in static_idx else [] for i in range(len(X))]
    ag.mobile_data = [[1.0, 2.0] if i in mobile_idx else [] for i in range(len(X))]
    ind, y, var = ag.get_sampled_dataset()
    th, hy = hyper_pair([2.0, 2.0], 1.0, 0.05, kind)
    ag.gp = make_gpr(kind, th.log_lengthscale, th.log_outputscale, th.log_noise,
                     X[ind] if len(ind) else X[:1], y if len(ind) else np.zeros(1), var if len(ind) else np.ones(1))
    ag._post_update()
    return ag, th


def test_agent_edge_semantics():
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 8, (40, 2))
    # nothing sampled yet: greedy from an empty base set (agent.py:308 would slogdet a 0x0 matrix: entropy 0)
    ag, th = _agent(X, set(), set())
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    picks = ag.greedy(3)
    assert picks == O.greedy_restructured(cov, np.zeros(40, bool), np.zeros(40, bool), 0.1, 1.0, 3)
    # everything already static: np.argmax of all -inf is 0 (agent.py:349)
    ag2, _ = _agent(X, set(range(40)), set())
    assert ag2.greedy(2) == [0, 0]
    # paths that add nothing new tie at H(B): first path wins (agent.py:402)
    ag3, _ = _agent(X, {1, 2, 3}, {4, 5, 6})
    assert ag3.best_path([[4, 5], [5, 6], [4]], [7]) == 0
    assert ag3.best_path([[4, 5]], [7]) == 0                                   # single path: agent.py:362-363
    # lists of different lengths, duplicates, overlap with the new static waypoint
    paths = [[10, 11, 12, 10], [7, 20], [21, 22, 23, 24, 25, 26, 27, 28, 29]]
    st = np.zeros(40, bool); st[[1, 2, 3]] = True
    mo = np.zeros(40, bool); mo[[4, 5, 6]] = True
    assert ag3.best_path(paths, [7]) == O.best_path_literal(cov, st, mo, 0.1, 1.0, paths, [7])


@pytest.mark.parametrize("criterion", ["entropy", "mutual_information"])
def test_long_paths_match_literal_loop(criterion):
    """Paths of hundreds of mobile locations (agent.py:373-400 scores whatever env.get_all_paths returns): more than
    the 128 slots of the shared-memory kernel, ragged, with repeats and already-sampled locations."""
    rng = np.random.default_rng(9)
    n = 700
    X = rng.uniform(0, 30, (n, 2))
    static = set(rng.choice(n, 60, replace=False).tolist())
    mobile = set(rng.choice(n, 40, replace=False).tolist())
    ag, th = _agent(X, static, mobile)
    ag.criterion = criterion
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    lens = [130, 257, 301, 12, 200, 129]
    paths = [rng.choice(n, L, replace=True).tolist() for L in lens]
    paths.append(paths[1][:200])                      # a prefix of another path
    st = np.zeros(n, bool); st[list(static)] = True
    mo = np.zeros(n, bool); mo[list(mobile)] = True
    new_static = [int(i) for i in rng.choice(n, 3, replace=False)]
    want, ut = O.best_path_literal(cov, st, mo, 0.1, 1.0, paths, new_static, criterion=criterion, return_utilities=True)
    assert ag.best_path(paths, new_static) == want
    got = ag._last_path_scores.cpu().numpy()
    np.testing.assert_allclose(got, ut, rtol=1e-8, atol=1e-8)


def test_enumerated_paths_score_the_same_as_lists():
    """Path enumeration -> slot matrix -> best_path on the device: the array form (no host lists) picks the path the
    list form and the literal reference loop pick (golden graph of the reference's default field: 860 locations)."""
    import os
    from algp_b200 import paths as P
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_paths.npz"))
    k = 2
    c = {name: g["c%d_%s" % (k, name)] for name in ("rc", "adj_ptr", "adj", "eptr", "eidx", "start", "heading", "waypoints",
                                                     "least_cost", "slack")}
    nodes = [tuple(r) for r in c["rc"].tolist()]
    ps = P.enumerate_paths_arrays(nodes, c["rc"], c["adj_ptr"], c["adj"], c["eptr"], c["eidx"], int(c["start"]),
                                  tuple(c["heading"].tolist()), c["waypoints"], float(c["least_cost"]), float(c["slack"]))
    assert len(ps) > 100
    n = 860
    rng = np.random.default_rng(2)
    X = rng.uniform(0, 30, (n, 2))
    static = set(rng.choice(n, 50, replace=False).tolist())
    mobile = set(rng.choice(n, 30, replace=False).tolist())
    ag, th = _agent(X, static, mobile)
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    st = np.zeros(n, bool); st[list(static)] = True
    mo = np.zeros(n, bool); mo[list(mobile)] = True
    new_static = [int(i) for i in rng.choice(n, 4, replace=False)]
    lists = ps.indices()
    want, ut = O.best_path_literal(cov, st, mo, 0.1, 1.0, lists, new_static, return_utilities=True)
    assert ag.best_path(lists, new_static) == want
    s_list = ag._last_path_scores.cpu().numpy()
    assert ag.best_path(ps.slots(), new_static) == want
    s_arr = ag._last_path_scores.cpu().numpy()
    np.testing.assert_allclose(s_list, ut, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(s_arr, ut, rtol=1e-8, atol=1e-8)


def test_state_is_extended_not_rebuilt_between_planning_steps():
    """Between two planning steps the sample flags only gain entries: the cached posterior state is extended by block
    appends (same object, more columns) and gives the picks / path / scores of a state factored from scratch."""
    rng = np.random.default_rng(12)
    n = 400
    X = rng.uniform(0, 20, (n, 2))
    static0 = set(rng.choice(n, 40, replace=False).tolist())
    mobile0 = set(rng.choice(n, 25, replace=False).tolist())
    ag, th = _agent(X, static0, mobile0)
    picks = ag.greedy(3)
    paths = [rng.choice(n, L, replace=False).tolist() for L in (12, 30, 7, 18)]
    best = ag.best_path(paths, picks)
    state0 = ag._hot_state["state"]
    ncols0 = state0.ncols
    # the robot collects: the picks become static samples, the chosen path's locations mobile samples
    for j in picks:
        ag.static_data[j] = ag.static_data[j] + [1.0]
    for j in paths[best]:
        ag.mobile_data[j] = ag.mobile_data[j] + [0.7]
    ag.collected = {'ind': list(range(len(picks) + len(paths[best]))), 'std': [], 'y': []}     # changes the flag stamp
    picks2 = ag.greedy(2)
    paths2 = [rng.choice(n, L, replace=False).tolist() for L in (9, 21, 140, 33)]
    best2 = ag.best_path(paths2, picks2)
    scores2 = ag._last_path_scores.cpu().numpy()
    assert ag._hot_state["state"] is state0 and state0.ncols > ncols0            # extended in place
    # the same situation on an agent that never saw the first step
    static1 = static0 | set(picks)
    mobile1 = mobile0 | set(paths[best])
    fresh, _ = _agent(X, static1, mobile1)
    assert fresh.greedy(2) == picks2
    assert fresh.best_path(paths2, picks2) == best2
    np.testing.assert_allclose(scores2, fresh._last_path_scores.cpu().numpy(), rtol=1e-10, atol=1e-9)
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    st = np.zeros(n, bool); st[list(static1)] = True
    mo = np.zeros(n, bool); mo[list(mobile1)] = True
    assert O.greedy_restructured(cov, st, mo, 0.1, 1.0, 2) == picks2
    assert O.best_path_literal(cov, st, mo, 0.1, 1.0, paths2, picks2) == best2


def test_staging_buffers_are_safe_to_reuse():
    """engine.to_dev copies through two reused pinned buffers per device; a buffer is only overwritten once the copy
    that last read it has finished, so a burst of uploads of different sizes and dtypes lands intact."""
    rng = np.random.default_rng(0)
    hosts, devs = [], []
    for i in range(12):
        n = int(rng.integers(1, 400000))
        if i % 3 == 0:
            a = rng.integers(-5, 5, size=(n, 3)).astype(np.int32)
            t = engine.to_dev(a, dtype=torch.int32)
        elif i % 3 == 1:
            a = rng.normal(size=n)
            t = engine.to_dev(a)
        else:
            a = (rng.random(n) < 0.5).astype(np.uint8)
            t = engine.to_dev(a, dtype=torch.uint8)
        hosts.append(a)
        devs.append(t)
    big = rng.normal(size=3_000_000)                         # larger than the initial buffers: they grow
    tb = engine.to_dev(big)
    torch.cuda.synchronize()
    for a, t in zip(hosts, devs):
        np.testing.assert_array_equal(t.cpu().numpy(), a)
    np.testing.assert_array_equal(tb.cpu().numpy(), big)
    assert engine.to_dev(np.zeros((0, 4))).shape == (0, 4)
    # float32 host data is converted, a device tensor passes through
    f32 = rng.normal(size=100).astype(np.float32)
    np.testing.assert_array_equal(engine.to_dev(f32).cpu().numpy(), f32.astype(np.float64))
    assert engine.to_dev(tb) is not None and engine.to_dev(tb).data_ptr() == tb.data_ptr()


def test_out_of_range_path_indices_raise_index_error():
    """agent.py:377 indexes a NumPy flag array with every path, so a location >= n raises IndexError in the reference.
    Lists go through NumPy indexing on the host; a caller's [P, k] slot array is range-checked on the device (the
    count comes back with the winner) -- and the agent stays usable afterwards."""
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 8, (40, 2))
    ag, th = _agent(X, {1, 2, 3}, {4, 5, 6})
    with pytest.raises(IndexError):
        ag.best_path([[10, 11], [12, 40]], [7])
    good = np.array([[10, 11, -1], [12, 13, 14], [20, 21, 22]], dtype=np.int32)
    want = ag.best_path(good, [7])
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    st = np.zeros(40, bool); st[[1, 2, 3]] = True
    mo = np.zeros(40, bool); mo[[4, 5, 6]] = True
    assert want == O.best_path_literal(cov, st, mo, 0.1, 1.0, [[10, 11], [12, 13, 14], [20, 21, 22]], [7])
    for bad_value in (40, 2 ** 31 - 1, -2, -(2 ** 31)):
        bad = good.copy()
        bad[1, 2] = bad_value
        with pytest.raises(IndexError):
            ag.best_path(bad, [7])
        assert ag.best_path(good, [7]) == want                  # no sticky CUDA error, same answer as before
    # the C entry point itself: count + blanking
    from algp_b200._lib import call, ptr, stream
    idx = engine.to_dev(np.array([0, -1, 39, 40, -2, 7], dtype=np.int32), dtype=torch.int32)
    cnt = torch.empty(1, dtype=torch.int64, device=idx.device)
    call("algp_check_indices", ptr(idx), 6, 40, ptr(cnt), stream())
    assert int(cnt.item()) == 2 and idx.cpu().tolist() == [0, -1, 39, -1, -1, 7]
    call("algp_check_indices", ptr(idx), 0, 40, ptr(cnt), stream())
    assert int(cnt.item()) == 0
