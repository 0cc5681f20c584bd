"""Parity of each CUDA kernel (through the C ABI) with the CPU oracle.  B200 only."""
import numpy as np
import pytest
import torch

import oracle as O
from algp_b200 import _lib, engine
from algp_b200._lib import call, ptr, stream
from gpu_helpers import dev, field_problem, hyper_pair

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- K1 kbuild
@pytest.mark.parametrize("kind", ["rbf", "matern"])
@pytest.mark.parametrize("d", [2, 6, 3])
def test_kbuild_fp64_matches_oracle(kind, d):
    rng = np.random.default_rng(d)
    x1 = rng.uniform(0, 20, size=(333, d))
    x2 = rng.uniform(0, 20, size=(517, d))
    th, hy = hyper_pair(np.linspace(2.0, 4.0, d), 1.3, 0.05, kind)
    K, _ = engine.kbuild(hy, dev(x1), dev(x2))
    ref = O.kernel_matrix(th, x1, x2, "fp64")
    out = K.cpu().numpy()
    assert out.shape == (333, 518)
    # bit-level agreement is not expected (different libm exp); 1e-13 relative to s^2 is
    np.testing.assert_allclose(out[:, :517], ref, rtol=1e-12, atol=1e-14)
    assert (out[:, 517:] == 0).all()


def test_kbuild_diag_padding_and_dot():
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 15, size=(200, 2))
    th, hy = hyper_pair([2.0, 3.0], 0.9, 0.02, "rbf")
    var = rng.uniform(0.01, 0.5, 200)
    vec = rng.normal(size=200)
    vpad = np.zeros(256)
    vpad[:200] = vec
    K, part = engine.kbuild(hy, dev(x), None, 256, 256, diag_add=dev(var), diag_scalar=hy.noise, pad_identity=True,
                            dot_vec=dev(vpad))
    out = K.cpu().numpy()
    ref = O.OracleGP(th, "fp64").cov_mat(x, white_noise_var=var, add_likelihood_var=True)
    np.testing.assert_allclose(out[:200, :200], ref, rtol=1e-12, atol=1e-14)
    assert (out[200:, :200] == 0).all() and (out[:200, 200:] == 0).all()
    np.testing.assert_array_equal(out[200:, 200:], np.eye(56))
    s = engine.rowsum(part, 1.0, 0.5).cpu().numpy()
    np.testing.assert_allclose(s[:200], ref @ vec + 0.5, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_kbuild_fp32_matches_ref32(kind):
    rng = np.random.default_rng(1)
    x1 = rng.uniform(0, 20, size=(300, 2))
    th, hy = hyper_pair([2.5, 3.5], 1.3, 0.02, kind)
    K, _ = engine.kbuild(hy, dev(x1), None, dtype=torch.float32)
    ref = O.kernel_matrix(th, x1, None, "ref32")
    np.testing.assert_allclose(K.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- K2 potrf / trtri / solves
def spd_problem(N, seed, kind="rbf"):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 40, size=(N, 2))
    th, hy = hyper_pair([3.0, 4.0], 1.0, 0.01, kind)
    var = rng.uniform(0.01, 0.02, N)
    A = O.OracleGP(th, "fp64").cov_mat(x, white_noise_var=var, add_likelihood_var=True)
    return x, var, th, hy, A


@pytest.mark.parametrize("N", [100, 128, 384, 645, 1000])
def test_potrf_trtri_match_numpy(N):
    x, var, th, hy, A = spd_problem(N, N)
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    f.check()
    L = np.tril(f.L.cpu().numpy())[:N, :N]
    Lref = np.linalg.cholesky(A)
    scale = np.abs(Lref).max()
    np.testing.assert_allclose(L, Lref, rtol=0, atol=1e-11 * scale)
    Linv = f.Linv.cpu().numpy()
    assert np.abs(np.triu(Linv, 1)).max() == 0.0
    np.testing.assert_allclose(Linv[:N, :N] @ Lref, np.eye(N), rtol=0, atol=1e-9)
    # padded part of both factors is the identity
    np.testing.assert_array_equal(np.tril(f.L.cpu().numpy())[N:, N:], np.eye(f.Npad - N))
    np.testing.assert_array_equal(Linv[N:, N:], np.eye(f.Npad - N))
    ldq = f.logdet_quad().cpu().numpy()
    assert ldq[0] == pytest.approx(np.linalg.slogdet(A)[1], rel=1e-11)


def test_potrf_reports_not_positive_definite():
    x, var, th, hy, A = spd_problem(200, 3)
    bad = -np.ones(200)            # pushes the diagonal negative
    f = engine.GPFactor(hy, dev(x), diag_add=dev(bad), diag_scalar=0.0)
    with pytest.raises(np.linalg.LinAlgError):
        f.check()
    assert int(f.info.item()) >= 1


@pytest.mark.parametrize("N", [200, 640])
def test_solve_matches_numpy(N):
    x, var, th, hy, A = spd_problem(N, N + 1)
    rng = np.random.default_rng(5)
    y = rng.normal(size=N)
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    alpha, beta = f.solve(dev(y))
    ref = np.linalg.solve(A, y)
    np.testing.assert_allclose(alpha.cpu().numpy()[:N], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    quad = f.logdet_quad(beta).cpu().numpy()[1]
    assert quad == pytest.approx(y @ ref, rel=1e-10)


def test_gemm_nt_matches_numpy():
    rng = np.random.default_rng(2)
    A = rng.normal(size=(256, 384))
    B = rng.normal(size=(384, 384))
    C = rng.normal(size=(256, 384))
    Cd = dev(C)
    engine.gemm_nt(dev(A), dev(B), Cd, -1.0, 1.0)
    np.testing.assert_allclose(Cd.cpu().numpy(), C - A @ B.T, rtol=1e-12, atol=1e-11)
    # lower-only SYRK form
    S = rng.normal(size=(384, 128))
    C2 = rng.normal(size=(384, 384))
    C2d = dev(C2)
    engine.gemm_nt(dev(S), dev(S), C2d, 1.0, 1.0, lower_only=True)
    full = C2 + S @ S.T
    got = C2d.cpu().numpy()
    # contract: the lower triangle (diagonal included) is updated; strictly-upper 128-blocks are
    # untouched; elements above the diagonal inside diagonal blocks are unspecified
    np.testing.assert_allclose(np.tril(got), np.tril(full), rtol=1e-12, atol=1e-11)
    for bi in range(3):
        for bj in range(bi + 1, 3):
            blk = (slice(bi * 128, bi * 128 + 128), slice(bj * 128, bj * 128 + 128))
            np.testing.assert_array_equal(got[blk], C2[blk])
    # a grid large enough for the 128x128 throughput shape (>= 120 tiles)
    A3 = rng.normal(size=(1536, 256))
    B3 = rng.normal(size=(1536, 256))
    C3 = rng.normal(size=(1536, 1536))
    C3d = dev(C3)
    engine.gemm_nt(dev(A3), dev(B3), C3d, 0.5, -1.0)
    np.testing.assert_allclose(C3d.cpu().numpy(), 0.5 * A3 @ B3.T - C3, rtol=1e-12, atol=1e-11)


def test_whiten_rownorm_matches_numpy():
    x, var, th, hy, A = spd_problem(300, 9)
    rng = np.random.default_rng(4)
    xs = rng.uniform(0, 40, size=(150, 2))
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    Ks, _ = f.cross(dev(xs))
    V, rn = f.whiten(Ks)
    Kref = O.kernel_matrix(th, xs, x, "fp64")
    import scipy.linalg as sla
    Vref = sla.solve_triangular(np.linalg.cholesky(A), Kref.T, lower=True).T
    np.testing.assert_allclose(V.cpu().numpy()[:150, :300], Vref, rtol=0, atol=1e-10)
    sq = engine.rowsum(rn).cpu().numpy()[:150]
    np.testing.assert_allclose(sq, (Vref ** 2).sum(1), rtol=1e-10)


# ---------------------------------------------------------------- K3 scoring
def scoring_problem(kind, n_side=20, n_base=70, seed=3, d_extra=0):
    X, y, tr, ytr, rng = field_problem(n_side, n_side + 3, n_base, seed, d_extra)
    n = len(X)
    d = X.shape[1]
    th, hy = hyper_pair(np.linspace(2.0, 3.0, d), 1.1, 0.04, kind)
    ss, ms = 0.1, 1.0
    static = np.zeros(n, bool)
    mobile = np.zeros(n, bool)
    static[tr[: n_base // 2]] = True
    mobile[tr[n_base // 3:]] = True
    pi0 = O.precisions_from_flags(static, mobile, ss, ms)
    cov = O.OracleGP(th, "fp64").cov_mat(X, add_likelihood_var=True)
    return X, th, hy, static, mobile, pi0, cov, ss, ms, rng


@pytest.mark.parametrize("kind", ["rbf", "matern"])
@pytest.mark.parametrize("k", [1, 5, 8, 11, 40, 128])
def test_score_sets_matches_oracle(kind, k):
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem(kind, d_extra=(1 if k == 11 else 0))
    n = len(X)
    Bc = 300 if k <= 8 else 40
    idx = np.full((Bc, k), -1, dtype=np.int32)
    delta = np.zeros((Bc, k))
    for c in range(Bc):
        m = int(rng.integers(1, k + 1))
        sel = rng.choice(n, m, replace=False)
        idx[c, :m] = sel
        delta[c, :m] = np.where(rng.random(m) < 0.3, 1 / ss ** 2, 1 / ms ** 2)
        if m >= 3 and c % 4 == 0:
            idx[c, m - 1] = idx[c, 0]           # duplicate -> idempotent
        if m >= 2 and c % 5 == 0:
            delta[c, 1] = 0.0                   # zero increment -> inactive slot
    base = np.nonzero(pi0 > 0)[0]
    state = engine.PosteriorState(hy, dev(X), base, pi0)
    got = state.score_sets(dev(idx, torch.int32), dev(delta)).cpu().numpy()
    ost = O.posterior_state(cov, pi0)
    assert state.H_base == pytest.approx(ost["H"], rel=1e-11)
    np.testing.assert_allclose(state.diagP.cpu().numpy(), np.diag(ost["P"]), rtol=0, atol=1e-10)
    want = O.score_sets_restructured(ost["P"], pi0, idx, delta, ost["H"])
    # log-det tier of the north star: rel 1e-8
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9)
    assert int(np.argmax(got)) == int(np.argmax(want))


@pytest.mark.parametrize("kind", ["rbf", "matern"])
@pytest.mark.parametrize("k,tile", [(1, 64), (5, 64), (8, 64), (8, 128), (8, 0), (3, 192), (8, -2), (6, -4),
                                    (8, "r64"), (5, "r128"), (1, "r192"), (7, "r1024")])
def test_score_sets_tiled_matches_oracle_and_streaming(kind, k, tile, monkeypatch):
    """algp_score_sets_tiled -- one launch of the k <= 8 kernel per column chunk with the accumulator fragments parked in
    a work buffer, or (tile 0 / -2 / -4) one launch with 4 / 2 / 4 independent warps per candidate and a last-arriver
    epilogue, or ("r<chunk>") the persistent launch that sweeps the chunks with the partial Grams in shared memory --
    against the oracle and the plain single launch, several calls on the same workspace: several chunks per candidate, a column count that is not a multiple of the
    64-column step (appended columns), empty / duplicate / zero-increment slots and skip flags."""
    from algp_b200 import _lib
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem(kind, n_side=24, n_base=300, d_extra=(2 if k == 5 else 0))
    n = len(X)
    Bc = 500
    idx = np.full((Bc, k), -1, dtype=np.int32)
    delta = np.zeros((Bc, k))
    for c in range(Bc):
        m = int(rng.integers(1, k + 1))
        idx[c, :m] = rng.choice(n, m, replace=False)
        delta[c, :m] = np.where(rng.random(m) < 0.3, 1 / ss ** 2, 1 / ms ** 2)
        if m >= 3 and c % 4 == 0:
            idx[c, m - 1] = idx[c, 0]
        if m >= 2 and c % 5 == 0:
            delta[c, 1] = 0.0
    base = np.nonzero(pi0 > 0)[0]
    state = engine.PosteriorState(hy, dev(X), base, pi0, capacity=8, cov_mode="never")
    picks = state.greedy(3, 1 / ss ** 2)                 # three appended columns: ncols = Npad + 3
    pi1 = pi0.copy()
    pi1[picks] += 1 / ss ** 2
    ost = O.posterior_state(cov, pi1)
    skip = np.zeros(n, dtype=np.uint8)
    skip[rng.choice(n, n // 10, replace=False)] = 1
    if isinstance(tile, str):                      # "r<chunk>": the persistent form
        assert _lib.lib.algp_set_score_resident(1) == 0
        tile = int(tile[1:])
    elif tile < 0:                                 # -2 / -4: the default policy with 2 / 4 warps per candidate forced
        monkeypatch.setenv("ALGP_SCORE_PARTS", str(-tile))
        tile = 0
    assert _lib.lib.algp_set_score_tile_cols(tile) == 0
    try:
        for sk in (None, skip):
            skd = None if sk is None else dev(sk, torch.uint8)
            state.score_mode = "tiled"
            got = state.score_sets(dev(idx, torch.int32), dev(delta), skip=skd).cpu().numpy()
            state.score_mode = "stream"
            ref = state.score_sets(dev(idx, torch.int32), dev(delta), skip=skd).cpu().numpy()
            np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-10)
            idx_o = idx if sk is None else np.where((idx >= 0) & (skip[np.maximum(idx, 0)] == 1), -1, idx)
            want = O.score_sets_restructured(ost["P"], pi1, idx_o, delta, ost["H"])
            np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9)
            assert int(np.argmax(got)) == int(np.argmax(want))
        # scalar delta, one candidate, and B = 0
        state.score_mode = "tiled"
        one = state.score_sets(dev(idx[7:8], torch.int32), None, delta_scalar=1 / ms ** 2).cpu().numpy()
        want1 = O.score_sets_restructured(ost["P"], pi1, idx[7:8], np.full((1, k), 1 / ms ** 2), ost["H"])
        np.testing.assert_allclose(one, want1, rtol=1e-8, atol=1e-9)
        assert state.score_sets(dev(idx[:0], torch.int32), dev(delta[:0])).numel() == 0
    finally:
        _lib.lib.algp_set_score_tile_cols(0)
        _lib.lib.algp_set_score_resident(0)
    assert _lib.lib.algp_set_score_tile_cols(100) == 1       # not a multiple of 64: rejected
    assert _lib.lib.algp_set_score_resident(2) == 1


def test_score_sets_resident_sweep_large_batch():
    """The persistent sweep on more candidates than one launch owns (80 % of 148 x 24 x 24 shared-memory slots): the
    batch is cut into several launches; every candidate is scored exactly once and equals the plain single launch.
    Also a batch smaller than the number of resident warps, and the launch count the library reports."""
    from algp_b200 import _lib
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem("rbf", n_side=24, n_base=300)
    n = len(X)
    B = 150000
    idx = rng.integers(0, n, size=(B, 8)).astype(np.int32)          # repeats inside a set are allowed (first one counts)
    idx[rng.random((B, 8)) < 0.1] = -1
    state = engine.PosteriorState(hy, dev(X), np.nonzero(pi0 > 0)[0], pi0, cov_mode="never")
    idx_d = dev(idx, torch.int32)
    state.score_mode = "stream"
    ref = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2)
    state.score_mode = "tiled"
    try:
        assert _lib.lib.algp_set_score_resident(1) == 0
        assert _lib.lib.algp_set_score_tile_cols(128) == 0
        nl = _lib.lib.algp_score_sets_tiled_launches(8, B, state.ncols, state.n_pad)
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        assert nl == -(-B // (sms * 24 * 24 * 4 // 5)) and nl >= 2
        for Bc in (B, 1000, 3):
            out = torch.full((Bc,), float("nan"), dtype=torch.float64, device="cuda")
            got = state.score_sets(idx_d[:Bc], None, delta_scalar=1 / ms ** 2, out=out)
            assert not bool(torch.isnan(got).any())
            assert float((got - ref[:Bc]).abs().max()) < 1e-10
    finally:
        _lib.lib.algp_set_score_tile_cols(0)
        _lib.lib.algp_set_score_resident(0)


def test_score_sets_split_candidates_share_one_workspace(monkeypatch):
    """Calls with 2 and with 4 warps per candidate alternate on the same state (same workspace): the arrival counters
    of the two variants are separate, so each call still finds its counters at a multiple of its PARTS."""
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem("rbf", n_side=24, n_base=300)
    n = len(X)
    idx = np.stack([rng.choice(n, 8, replace=False) for _ in range(900)]).astype(np.int32)
    delta = np.full(idx.shape, 1 / ms ** 2)
    state = engine.PosteriorState(hy, dev(X), np.nonzero(pi0 > 0)[0], pi0, cov_mode="never")
    state.score_mode = "stream"
    ref = state.score_sets(dev(idx, torch.int32), dev(delta)).cpu().numpy()
    state.score_mode = "tiled"
    for parts in (2, 4, 4, 2, 4, 2, 2):
        monkeypatch.setenv("ALGP_SCORE_PARTS", str(parts))
        for B in (900, 37):
            got = state.score_sets(dev(idx[:B], torch.int32), dev(delta[:B])).cpu().numpy()
            np.testing.assert_allclose(got, ref[:B], rtol=1e-12, atol=1e-10)


def test_score_sets_tiled_empty_base():
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem("rbf")
    n = len(X)
    state = engine.PosteriorState(hy, dev(X), np.zeros(0, dtype=np.int64), np.zeros(n), capacity=4, cov_mode="never")
    idx = np.stack([rng.choice(n, 6, replace=False) for _ in range(50)]).astype(np.int32)
    delta = np.full(idx.shape, 1 / ms ** 2)
    state.score_mode = "tiled"
    got = state.score_sets(dev(idx, torch.int32), dev(delta)).cpu().numpy()
    ost = O.posterior_state(cov, np.zeros(n))
    np.testing.assert_allclose(got, O.score_sets_restructured(ost["P"], np.zeros(n), idx, delta, ost["H"]), rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
@pytest.mark.parametrize("k", [1, 5, 8, 11, 40, 128])
def test_score_sets_resident_cov_matches_oracle(kind, k):
    """cov_mode="always": the posterior covariance P of the base set is built once (kernel matrix + SYRK) and every
    candidate is scored from its k(k+1)/2 gathered entries -- same oracle, same slot semantics (empty, duplicate and
    zero-increment slots, skip flags) as the streaming kernels."""
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem(kind, d_extra=(1 if k == 11 else 0))
    n = len(X)
    Bc = 300 if k <= 8 else 40
    idx = np.full((Bc, k), -1, dtype=np.int32)
    delta = np.zeros((Bc, k))
    for c in range(Bc):
        m = int(rng.integers(1, k + 1))
        sel = rng.choice(n, m, replace=False)
        idx[c, :m] = sel
        delta[c, :m] = np.where(rng.random(m) < 0.3, 1 / ss ** 2, 1 / ms ** 2)
        if m >= 3 and c % 4 == 0:
            idx[c, m - 1] = idx[c, 0]
        if m >= 2 and c % 5 == 0:
            delta[c, 1] = 0.0
    base = np.nonzero(pi0 > 0)[0]
    state = engine.PosteriorState(hy, dev(X), base, pi0, cov_mode="always")
    stream = engine.PosteriorState(hy, dev(X), base, pi0, cov_mode="never")
    skip = np.zeros(n, dtype=np.uint8)
    skip[rng.choice(n, n // 10, replace=False)] = 1
    for sk in (None, dev(skip, torch.uint8)):
        got = state.score_sets(dev(idx, torch.int32), dev(delta), skip=sk).cpu().numpy()
        assert state.P is not None and stream.P is None
        ref = stream.score_sets(dev(idx, torch.int32), dev(delta), skip=sk).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-10)
    ost = O.posterior_state(cov, pi0)
    Pd = np.tril(state.P.cpu().numpy()[:n, :n])
    np.testing.assert_allclose(Pd, np.tril(ost["P"]), rtol=0, atol=1e-10)
    want = O.score_sets_restructured(ost["P"], pi0, idx, delta, ost["H"])
    got = state.score_sets(dev(idx, torch.int32), dev(delta)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9)
    assert int(np.argmax(got)) == int(np.argmax(want))


def test_resident_cov_policy_and_commits():
    """cov_mode="auto" streams until the streamed work would have paid for the build, then switches; a commit (append
    / append_block) KEEPS the covariance: the appended columns are folded in lazily by rank-k downdates
    (algp_cov_downdate) before the next scoring call, and the scores are those of a freshly built state."""
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem("rbf")
    n = len(X)
    base = np.nonzero(pi0 > 0)[0]
    state = engine.PosteriorState(hy, dev(X), base, pi0, is_static=static, capacity=40)
    free = np.nonzero(pi0 == 0)[0]
    idx = rng.choice(free, (4000, 8)).astype(np.int32)
    idx_d = dev(idx, torch.int32)
    first = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2).cpu().numpy()
    assert state.P is None                                         # one small batch does not pay for the build
    state._stream_s = state.cov_build_seconds()                    # ... as if enough had been streamed
    second = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2).cpu().numpy()
    assert state.P is not None
    np.testing.assert_allclose(second, first, rtol=1e-11, atol=1e-10)
    P_before = state.P
    pi1 = pi0.copy()
    # one rank-1 commit, then blocks of 19 and 16 (36 pending columns: two downdate passes, 32 + 4), then another single one
    j = torch.tensor([int(free[0])], dtype=torch.int64, device="cuda")
    state.append(j, 1 / ss ** 2)
    pi1[free[0]] += 1 / ss ** 2
    for blk in (free[5:24], free[40:56]):
        state.append_block(blk, 1 / ms ** 2, mark_static=False)
        pi1[blk] += 1 / ms ** 2
    assert state.ncols - state._P_ncols == 36 > _lib.lib.algp_cov_downdate_max_cols()
    assert state.P is P_before and state._P_ncols < state.ncols      # kept, not yet synchronised
    after = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2).cpu().numpy()
    assert state.P is P_before and state._P_ncols == state.ncols
    fresh = engine.PosteriorState(hy, dev(X), np.nonzero(pi1 > 0)[0], pi1, cov_mode="always")
    want = fresh.score_sets(idx_d, None, delta_scalar=1 / ms ** 2, H_base=state.H_base).cpu().numpy()
    np.testing.assert_allclose(after, want, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.tril(state.P.cpu().numpy()[:n, :n]), np.tril(fresh.P.cpu().numpy()[:n, :n]), rtol=0, atol=1e-10)
    state.score_mode = "stream"
    mode, state.cov_mode = state.cov_mode, "never"
    P_keep, state.P = state.P, None
    streamed = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2).cpu().numpy()
    state.P, state.cov_mode = P_keep, mode
    np.testing.assert_allclose(after, streamed, rtol=1e-10, atol=1e-9)
    state.append(torch.tensor([int(free[30])], dtype=torch.int64, device="cuda"), 1 / ss ** 2)
    pi1[free[30]] += 1 / ss ** 2
    again = state.score_sets(idx_d, None, delta_scalar=1 / ms ** 2).cpu().numpy()
    fresh2 = engine.PosteriorState(hy, dev(X), np.nonzero(pi1 > 0)[0], pi1, cov_mode="always")
    np.testing.assert_allclose(again, fresh2.score_sets(idx_d, None, delta_scalar=1 / ms ** 2, H_base=state.H_base).cpu().numpy(),
                               rtol=1e-9, atol=1e-9)
    state.drop_cov()
    assert state.P is None and state._stream_s == 0.0
    never = engine.PosteriorState(hy, dev(X), base, pi0, cov_mode="never")
    never._stream_s = 1e9
    never.score_sets(idx_d, None, delta_scalar=1.0)
    assert never.P is None
    with pytest.raises(ValueError):
        engine.PosteriorState(hy, dev(X), base, pi0, cov_mode="sometimes")


def test_resident_cov_i8_build_matches_dmma():
    """precision="i8": the SYRK of the covariance build runs as an exact INT8 digit GEMM (lower tiles only)."""
    X, y, tr, ytr, rng = field_problem(48, 48, 1100, seed=4)
    th, hy = hyper_pair([3.0, 3.0], 1.0, 0.01, "rbf")
    n = len(X)
    pi0 = np.zeros(n); pi0[tr] = 100.0
    s64 = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0, cov_mode="always")
    s8 = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0, precision="i8", cov_mode="always")
    free = np.setdiff1d(np.arange(n), tr)
    idx = dev(rng.choice(free, (500, 8)).astype(np.int32), torch.int32)
    sc64 = s64.score_sets(idx, None, delta_scalar=1.0).cpu().numpy()
    sc8 = s8.score_sets(idx, None, delta_scalar=1.0).cpu().numpy()
    assert s8.P is not None and s64.P is not None
    np.testing.assert_allclose(np.tril(s8.P.cpu().numpy()[:n, :n]), np.tril(s64.P.cpu().numpy()[:n, :n]), rtol=0, atol=1e-11)
    np.testing.assert_allclose(sc8, sc64, rtol=1e-10, atol=1e-9)


def test_score_sets_empty_base_and_scalar_delta():
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem("rbf")
    n = len(X)
    pi_none = np.zeros(n)
    state = engine.PosteriorState(hy, dev(X), [], pi_none)
    idx = np.stack([rng.choice(n, 6, replace=False) for _ in range(64)]).astype(np.int32)
    got = state.score_sets(dev(idx, torch.int32), None, delta_scalar=1.0).cpu().numpy()
    want = O.score_sets_restructured(cov, pi_none, idx, np.ones(idx.shape), 0.0)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_greedy_matches_literal_reference_loop(kind):
    X, th, hy, static, mobile, pi0, cov, ss, ms, rng = scoring_problem(kind, n_side=12, n_base=40)
    base = np.nonzero(pi0 > 0)[0]
    state = engine.PosteriorState(hy, dev(X), base, pi0, is_static=static, capacity=8)
    picks, uts = state.greedy(4, 1 / ss ** 2, return_utilities=True)
    p_lit, u_lit = O.greedy_literal(cov, static, mobile, ss, ms, 4, return_utilities=True)
    assert picks == [int(p) for p in p_lit]
    fin = np.isfinite(u_lit)
    assert (np.isfinite(uts) == fin).all()
    np.testing.assert_allclose(uts[fin], u_lit[fin], rtol=0, atol=1e-9)
    # scoring against the appended state == literal entropy of the grown set
    st2 = static.copy()
    st2[picks] = True
    assert state.H_base == pytest.approx(O.set_entropy_literal(cov, st2, mobile, ss, ms), rel=1e-10)


def test_argmax_first_max_semantics():
    x = np.array([1.0, 5.0, -np.inf, 5.0, 2.0] * 1000)
    state_work = torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device="cuda")
    out = torch.empty(2, dtype=torch.int64, device="cuda")
    xd = dev(x)
    _lib.call("algp_argmax", _lib.ptr(xd), len(x), 7, _lib.ptr(out), _lib.ptr(state_work), _lib.stream())
    assert int(out[1].item()) == 1 + 7
    assert out[0:1].view(torch.float64).item() == 5.0
    # one-CTA path (n <= 2^18) and two-kernel path (beyond), ties, all -inf, a single value
    rng = np.random.default_rng(3)
    for n in (1, 31, 1024, 5000, 16384, 16385, 262145, 700001):
        v = np.round(rng.normal(size=n), 2)
        vd = dev(v)
        _lib.call("algp_argmax", _lib.ptr(vd), n, 0, _lib.ptr(out), _lib.ptr(state_work), _lib.stream())
        assert int(out[1].item()) == int(np.argmax(v)), n
        assert out[0:1].view(torch.float64).item() == v.max()
    inf = dev(np.full(777, -np.inf))
    _lib.call("algp_argmax", _lib.ptr(inf), 777, 0, _lib.ptr(out), _lib.ptr(state_work), _lib.stream())
    assert int(out[1].item()) == 0                                        # np.argmax of all -inf is 0 (agent.py:349)


# ---------------------------------------------------------------- TF32 mode (tcgen05)
def test_split_tf32_planes():
    rng = np.random.default_rng(0)
    M = rng.normal(size=(256, 384)) * np.exp(rng.normal(size=(256, 384)) * 3)
    f = engine.GPFactor.__new__(engine.GPFactor)
    hi, lo = engine.GPFactor.split_tf32(f, dev(M))
    hi, lo = hi.cpu().numpy(), lo.cpu().numpy()
    assert hi.shape == (256, 384)
    assert (hi.view(np.uint32) & 0x1FFF == 0).all()                 # exactly representable in TF32
    np.testing.assert_allclose(hi.astype(np.float64) + lo.astype(np.float64), M, rtol=2e-7 * 2 ** -10 + 1e-10, atol=0)
    assert np.abs(lo).max() <= np.abs(M).max() * 2.0 ** -10


@pytest.mark.parametrize("N,M", [(128, 128), (300, 150), (640, 1000), (1536, 700)])
def test_variance_tf32_tier(N, M):
    """precision='tf32': split-TF32 tcgen05 path against the fp64 oracle at the north star's 1e-4 tier
    (absolute, relative to the prior scale s^2 = 1)."""
    x, var, th, hy, A = spd_problem(N, N + 7)
    rng = np.random.default_rng(N)
    xs = rng.uniform(0, 40, size=(M, 2))
    y = rng.normal(size=N)
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    mu64, v64 = f.mean_var(dev(xs), dev(y - y.mean()), float(y.mean()))
    mu32, v32 = f.mean_var(dev(xs), dev(y - y.mean()), float(y.mean()), precision="tf32")
    f.check()
    _, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, return_var=True)
    np.testing.assert_allclose(v64.cpu().numpy(), v_o, rtol=0, atol=1e-9)
    np.testing.assert_allclose(v32.cpu().numpy(), v_o, rtol=0, atol=1e-4)
    np.testing.assert_array_equal(mu32.cpu().numpy(), mu64.cpu().numpy())   # the mean stays fp64


# ---------------------------------------------------------------- INT8 digit mode (tcgen05 kind::i8)
def untile_i8(tiles, rows, cols, S, tile_rows):
    """[row tile][32-col k chunk][plane][row group][k half][8 rows][16 B] (include/algp_b200.h) -> [S, rows, cols]"""
    t = tiles.reshape(rows // tile_rows, cols // 32, S, tile_rows // 8, 2, 8, 16)
    return t.transpose(2, 0, 3, 5, 1, 4, 6).reshape(S, rows, cols)


@pytest.mark.parametrize("S,tile_rows", [(2, 128), (5, 64), (7, 128), (8, 64)])
def test_split_i8_planes(S, tile_rows):
    """Digit expansion: |d| <= 64 and the planes reconstruct every row to 2^(-7S) of its scale."""
    rng = np.random.default_rng(S)
    M = rng.normal(size=(256, 384)) * np.exp(rng.normal(size=(256, 384)) * 3)
    M[3] = 0.0                                           # an all-zero row
    M[5, 7] = 1.0; M[5, 8:] *= 1e-30                     # power-of-two row maximum
    f = engine.GPFactor.__new__(engine.GPFactor)
    tiles, scale = engine.GPFactor.split_i8(f, dev(M), S, tile_rows)
    planes = untile_i8(tiles.cpu().numpy(), 256, 384, S, tile_rows).astype(np.float64)
    scale = scale.cpu().numpy()
    assert planes.shape == (S, 256, 384) and np.abs(planes).max() <= 64
    mx = np.abs(M).max(axis=1)
    nz = mx > 0
    assert (scale[nz] > mx[nz]).all() and (scale[nz] <= 2 * mx[nz]).all()      # 2^e with |row| / 2^e in [0.5, 1)
    assert np.array_equal(np.log2(scale), np.round(np.log2(scale)))
    rec = sum(planes[p] * 2.0 ** (-6 - 7 * p) for p in range(S)) * scale[:, None]
    assert (np.abs(rec - M) <= 2.0 ** (-7 * S) * scale[:, None]).all()


@pytest.mark.parametrize("N,M", [(128, 128), (300, 150), (640, 1000), (1536, 700)])
def test_variance_i8_fp64_tier(N, M):
    """precision='i8': exact INT8 digit GEMMs on tcgen05 against the fp64 oracle at the fp64 tier
    (abs 1e-9 relative to the prior scale s^2 = 1), and bit-identical means."""
    x, var, th, hy, A = spd_problem(N, N + 7)
    rng = np.random.default_rng(N)
    xs = rng.uniform(0, 40, size=(M, 2))
    y = rng.normal(size=N)
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    mu64, v64 = f.mean_var(dev(xs), dev(y - y.mean()), float(y.mean()))
    mu8, v8 = f.mean_var(dev(xs), dev(y - y.mean()), float(y.mean()), precision="i8")
    f.check()
    _, v_o = O.predictive_distribution_chol(O.OracleGP(th, "fp64"), x, y, xs, var, return_var=True)
    np.testing.assert_allclose(v8.cpu().numpy(), v_o, rtol=0, atol=1e-9)
    np.testing.assert_allclose(v8.cpu().numpy(), v64.cpu().numpy(), rtol=0, atol=1e-10)
    np.testing.assert_array_equal(mu8.cpu().numpy(), mu64.cpu().numpy())


def test_trmm_i8_slices_converge():
    """Each extra digit plane buys ~7 bits: the row norms converge geometrically to the DMMA result."""
    N, M = 512, 256
    x, var, th, hy, A = spd_problem(N, N + 3)
    rng = np.random.default_rng(5)
    xs = rng.uniform(0, 40, size=(M, 2))
    f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    Ks, _ = f.cross(dev(xs))
    _, rn = f.whiten(Ks, want_V=False)
    ref = rn.sum(1).cpu().numpy()
    errs = []
    for S in (3, 4, 5, 6, 7, 8):
        got = f.whiten_norm_i8(Ks, nslices=S).sum(1).cpu().numpy()
        errs.append(np.abs(got - ref).max())
    assert errs[-1] < 1e-12 and errs[-2] < 1e-10
    assert all(errs[i + 1] < errs[i] / 16 or errs[i + 1] < 1e-13 for i in range(len(errs) - 1)), errs


def test_gemm_nt_i8_matches_numpy():
    """C = alpha A B^T + beta C from digit tiles (8 planes): fp64-grade, normal and transposed stores."""
    rng = np.random.default_rng(11)
    M, N, K = 256, 192, 320
    A = rng.normal(size=(M, K)) * np.exp(rng.normal(size=(M, 1)) * 2)
    B = rng.normal(size=(N, K)) * np.exp(rng.normal(size=(N, 1)) * 2)
    C0 = rng.normal(size=(M, N))
    f = engine.GPFactor.__new__(engine.GPFactor)
    at, asc = engine.GPFactor.split_i8(f, dev(A), 8, 128)
    bt, bsc = engine.GPFactor.split_i8(f, dev(B), 8, 64)
    ref = A @ B.T
    tol = 1e-13 * np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * K
    C = dev(C0.copy())
    call("algp_gemm_nt_i8", ptr(at), ptr(asc), M, ptr(bt), ptr(bsc), N, K, 8, -0.5, 2.0, ptr(C), N, 0, 0, stream())
    assert (np.abs(C.cpu().numpy() - (-0.5 * ref + 2.0 * C0)) <= tol + 1e-14).all()
    Ct = dev(np.zeros((N, M)))
    call("algp_gemm_nt_i8", ptr(at), ptr(asc), M, ptr(bt), ptr(bsc), N, K, 8, 1.0, 0.0, ptr(Ct), M, 1, 0, stream())
    assert (np.abs(Ct.cpu().numpy().T - ref) <= tol).all()


@pytest.mark.parametrize("N,base", [(256, 128), (640, 128), (1000, 256), (1536, 512), (2304, 256)])
def test_potrf_inv_i8_matches_dmma(N, base):
    """Recursive INT8 digit factorisation: L and L^-1 agree with the DMMA potrf + trtri and with the oracle."""
    x, var, th, hy, A = spd_problem(N, N + 1)
    f64 = engine.GPFactor(hy, dev(x), diag_add=dev(var))
    f8 = engine.GPFactor.__new__(engine.GPFactor)
    Npad = f64.Npad
    L, _ = engine.kbuild(hy, dev(x), None, Npad, Npad, dev(var), hy.noise, True)
    Linv = torch.full((Npad, Npad), float("nan"), dtype=torch.float64, device=L.device)
    info = torch.zeros(1, dtype=torch.int32, device=L.device)
    engine.potrf_inv_i8(L, Linv, info, nslices=8, base=base)
    assert int(info.item()) == 0
    Lh, Lr = torch.tril(L).cpu().numpy()[:N, :N], torch.tril(f64.L).cpu().numpy()[:N, :N]
    Lo = np.linalg.cholesky(A)
    np.testing.assert_allclose(Lh, Lo, rtol=0, atol=1e-11)
    np.testing.assert_allclose(Lr, Lo, rtol=0, atol=1e-11)
    Li = Linv.cpu().numpy()
    assert np.isfinite(Li).all() and np.abs(np.triu(Li, 1)).max() == 0.0          # clean lower-triangular inverse
    Li_ref = f64.Linv.cpu().numpy()
    scale = np.abs(Li_ref).max()
    np.testing.assert_allclose(Li, Li_ref, rtol=0, atol=1e-9 * scale)
    np.testing.assert_allclose((Li[:N, :N] @ Lo), np.eye(N), rtol=0, atol=1e-9)


def test_potrf_inv_i8_reports_not_pd():
    x, var, th, hy, A = spd_problem(700, 3)
    Npad = 768
    L, _ = engine.kbuild(hy, dev(x), None, Npad, Npad, dev(var), hy.noise, True)
    L[600, 600] = -5.0
    Linv = torch.empty((Npad, Npad), dtype=torch.float64, device=L.device)
    info = torch.zeros(1, dtype=torch.int32, device=L.device)
    engine.potrf_inv_i8(L, Linv, info, nslices=8, base=256)
    assert int(info.item()) == 601


@pytest.mark.parametrize("N", [128, 200, 645, 1000])
def test_potf2_rank_variants_agree(N):
    """The diagonal-block kernel eliminates 1, 2 or 4 columns per barrier: same L and L^-1, and the same
    not-positive-definite column."""
    x, var, th, hy, A = spd_problem(N, N + 2)
    Lo = np.linalg.cholesky(A)
    try:
        for r in (1, 2, 4):
            call("algp_set_potf2_rank", r)
            f = engine.GPFactor(hy, dev(x), diag_add=dev(var))
            f.check()
            np.testing.assert_allclose(np.tril(f.L.cpu().numpy())[:N, :N], Lo, rtol=0, atol=1e-11)
            Li = f.Linv.cpu().numpy()[:N, :N]
            np.testing.assert_allclose(Li @ Lo, np.eye(N), rtol=0, atol=1e-9)
            assert np.abs(np.triu(f.Linv.cpu().numpy(), 1)).max() == 0.0
            # a negative pivot in the middle of a 4-column step
            Npad = f.Npad
            Abad, _ = engine.kbuild(hy, dev(x), None, Npad, Npad, dev(var), hy.noise, True)
            bad = min(N - 1, 70)
            Abad[bad, bad] = -1.0
            Linv = torch.empty_like(Abad)
            info = torch.zeros(1, dtype=torch.int32, device=Abad.device)
            call("algp_potrf", ptr(Abad), Npad, Npad, ptr(Linv), Npad, ptr(info), stream())
            assert int(info.item()) == bad + 1
    finally:
        call("algp_set_potf2_rank", 2)


def test_posterior_state_i8_build_matches_dmma():
    """precision='i8': W^T = Sigma_{:,B} L^-T and diag(P) through the digit GEMM (store + row-norm epilogue) agree with
    the DMMA build, and so do the scores computed from them."""
    X, y, tr, ytr, rng = field_problem(48, 48, 1100, seed=4)
    th, hy = hyper_pair([3.0, 3.0], 1.0, 0.01, "rbf")
    n = len(X)
    pi0 = np.zeros(n); pi0[tr] = 100.0
    s64 = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0)
    s8 = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0, precision="i8")
    assert s8.Npad >= 1024                                       # the digit path is the one that ran
    np.testing.assert_allclose(s8.Wt.cpu().numpy(), s64.Wt.cpu().numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(s8.diagP.cpu().numpy(), s64.diagP.cpu().numpy(), rtol=0, atol=1e-11)
    free = np.setdiff1d(np.arange(n), tr)
    idx = rng.choice(free, (500, 8)).astype(np.int32)
    sc64 = s64.score_sets(dev(idx, torch.int32), None, delta_scalar=1.0).cpu().numpy()
    sc8 = s8.score_sets(dev(idx, torch.int32), None, delta_scalar=1.0).cpu().numpy()
    np.testing.assert_allclose(sc8, sc64, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("scalar,pre", [(0, 0), (0, 3), (1, 0), (1, 5)])
@pytest.mark.parametrize("k", [1, 5, 16, 23])
def test_append_block_equals_successive_appends(k, scalar, pre):
    """k locations committed at once: the same Wt columns, diag(P), precisions and scores as k rank-1 appends -- through
    the DMMA pass and the scalar one, on a column count that is a multiple of 16 and (after `pre` greedy picks) not."""
    from algp_b200 import _lib
    assert _lib.lib.algp_set_append_block_scalar(scalar) == 0
    try:
        _append_block_case(k, pre)
    finally:
        _lib.lib.algp_set_append_block_scalar(0)


def _append_block_case(k, pre):
    X, y, tr, ytr, rng = field_problem(30, 34, 300, seed=8)
    th, hy = hyper_pair([3.0, 2.5], 1.3, 0.02, "matern")
    n = len(X)
    pi0 = np.zeros(n); pi0[tr] = 100.0
    free = np.setdiff1d(np.arange(n), tr)
    chosen = rng.choice(free, k, replace=False)
    chosen[0] = tr[3]                                   # an already-sampled location gets more precision
    seq = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0, capacity=40)
    blk = engine.PosteriorState(hy, dev(X), tr, pi0, is_static=pi0 > 0, capacity=40)
    jb = torch.empty(1, dtype=torch.int64, device=seq.X.device)
    if pre:
        assert seq.greedy(pre, 100.0) == blk.greedy(pre, 100.0)
    for j in chosen:
        jb.fill_(int(j))
        seq.append(jb, 1.0, mark_static=False)
    blk.append_block([int(j) for j in chosen], 1.0, mark_static=False)
    assert blk.ncols == seq.ncols == seq.Npad + pre + k
    np.testing.assert_allclose(blk.Wt.cpu().numpy(), seq.Wt.cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(blk.diagP.cpu().numpy(), seq.diagP.cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_array_equal(blk.pi.cpu().numpy(), seq.pi.cpu().numpy())
    np.testing.assert_array_equal(blk.is_static.cpu().numpy(), seq.is_static.cpu().numpy())
    idx = rng.choice(free, (64, 8)).astype(np.int32)
    np.testing.assert_allclose(blk.score_sets(dev(idx, torch.int32), None, delta_scalar=1.0).cpu().numpy(),
                               seq.score_sets(dev(idx, torch.int32), None, delta_scalar=1.0).cpu().numpy(), rtol=1e-11, atol=1e-10)


def test_split_i8_occupancy_masks():
    """mask[tile][chunk] bit p <=> plane p of that (row tile, 32-column chunk) holds a non-zero digit."""
    rng = np.random.default_rng(21)
    S, tile_rows, rows, cols = 7, 64, 192, 320
    M = rng.normal(size=(rows, cols))
    M[:, 0] = 3.0                                            # every row's scale is 4: digits of the other columns shift down
    M[64:128, 32:96] = 0.0                                   # two all-zero chunks of the second tile
    M[128:, 160:192] = 2.0 ** -30 * rng.normal(size=(64, 32))   # a chunk that only reaches the low planes
    f = engine.GPFactor.__new__(engine.GPFactor)
    tiles, scale, mask = engine.GPFactor.split_i8(f, dev(M), S, tile_rows, want_mask=True)
    planes = untile_i8(tiles.cpu().numpy(), rows, cols, S, tile_rows)
    want = np.zeros((rows // tile_rows, cols // 32), dtype=np.uint8)
    for p in range(S):
        nz = (planes[p].reshape(rows // tile_rows, tile_rows, cols // 32, 32) != 0).any(axis=(1, 3))
        want |= (nz.astype(np.uint8) << p)
    ld = (cols // 32 + 7) // 8 * 8
    got = mask.cpu().numpy()[: (rows // tile_rows) * ld].reshape(rows // tile_rows, ld)
    np.testing.assert_array_equal(got[:, : cols // 32], want)
    assert (got[:, cols // 32:] == 0).all()
    assert want[1, 1] == 0 and want[1, 2] == 0 and want[2, 5] != 0 and (want[2, 5] & 0b111) == 0


@pytest.mark.parametrize("S", [5, 7, 8])
def test_trmm_i8_masked_equals_unmasked_on_sparse_operands(S):
    """Occupancy masks only skip digit tiles that are exactly zero: same result bit for bit, on operands whose
    entries decay away from the diagonal (so that many tiles / planes are empty), and the fp64 tier vs DMMA."""
    rng = np.random.default_rng(S)
    N, M = 768, 512
    x = np.sort(rng.uniform(0, 400, N))[:, None]             # 1-D, sorted: K(X,X) and Linv decay away from the diagonal
    xs = np.sort(rng.uniform(0, 400, M))[:, None]
    th, hy = hyper_pair([2.0], 1.0, 0.05, "rbf")
    f = engine.GPFactor(hy, dev(x), diag_add=None)
    f.check()
    Ks, _ = f.cross(dev(xs))
    _, rn = f.whiten(Ks, want_V=False)
    ref = rn.sum(1).cpu().numpy()
    masked = f.whiten_norm_i8(Ks, nslices=S, use_masks=True).cpu().numpy()
    plain = f.whiten_norm_i8(Ks, nslices=S, use_masks=False).cpu().numpy()
    np.testing.assert_array_equal(masked, plain)
    tol = {5: 1e-7, 7: 1e-10, 8: 1e-11}[S]
    np.testing.assert_allclose(masked.sum(1), ref, rtol=0, atol=tol)
    _, _, km = f.split_i8(Ks, S, 128, want_mask=True)
    assert float((km == 0).float().mean()) > 0.3             # the case really exercises skipped chunks
