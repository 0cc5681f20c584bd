#!/usr/bin/env python
"""Benchmark of the GP / information-gain hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline metric: info-gain candidates/sec at N=4096 training points -- BASELINE.json
configs[2]: 65 536 candidate sample sets of size 8 scored against the N=4096 factor of a
128x128 mixture-of-Gaussians field.  One "step" = score every candidate set
of the batch + argmax (+ the 16-byte all-gather of per-rank winners when N > 1).  Each rank
holds the replicated factor and its own 65 536 sets (weak scaling).  The one-off factor / W
build is reported separately.  The second BASELINE metric, GP fit+predict ms (N=4096 and
N=16384, fp64), rides along under "fit_predict" with its own per-kernel rooflines.

`--impl reference` times the reference's own CPU algorithm for the same metric (the oracle
port of agent.py:373-400: fancy-index + np.linalg.slogdet per candidate, all host cores) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "info-gain candidates/sec at N=4096 train pts"
UNIT = "candidates/s"
FIELD = 128            # 128 x 128 locations
N_BASE = 4096
N_CAND = 65536
K_SET = 8
STATIC_STD = 0.1
MOBILE_STD = 1.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def field_grid():
    """The synthetic field of SURVEY.md 8d: the product's own generate_gaussian_data (reference utils.py:90-108) with
    numpy.random.default_rng(1)."""
    from algp_b200.utils import generate_gaussian_data
    grid, y = generate_gaussian_data(FIELD, FIELD, seed=1)
    return grid.astype(np.float64), y


def candidate_sets(rest, seed, n_cand=N_CAND):
    """n_cand sets of K_SET distinct non-base locations: the first K_SET distinct values of 32 uniform draws per set
    (vectorised; the same sets as the round-1 per-row np.unique loop)."""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(rest), size=(n_cand, 32))
    idx = np.empty((n_cand, K_SET), dtype=np.int32)
    for lo in range(0, n_cand, 8192):
        p = pick[lo:lo + 8192]
        same = p[:, :, None] == p[:, None, :]                                  # [c, j, i]: draw j equals draw i
        earlier = np.tril(np.ones((32, 32), dtype=bool), -1)[None]              # i < j
        first = ~(same & earlier).any(axis=2)                                  # draw j is the first of its value
        rank = np.cumsum(first, axis=1)
        keep = first & (rank <= K_SET)
        assert (keep.sum(axis=1) == K_SET).all()
        idx[lo:lo + 8192] = rest[p[keep].reshape(-1, K_SET)]
    return idx


def workload(seed_sets=2, n_cand=N_CAND):
    """Synthetic config B: field, base set (seed 1), candidate sets (seed_sets)."""
    grid, y = field_grid()
    n = len(grid)
    rng = np.random.default_rng(1)
    base = np.sort(rng.choice(n, N_BASE, replace=False))
    rest = np.setdiff1d(np.arange(n), base)
    idx = candidate_sets(rest, seed_sets, n_cand)
    delta = np.full((n_cand, K_SET), 1.0 / MOBILE_STD ** 2)
    delta[:, 0] = 1.0 / STATIC_STD ** 2        # one static + seven mobile readings per set
    hyper = dict(ls=[FIELD / 16.0, FIELD / 16.0], os=1.0, noise=1e-2, kind="rbf")
    return grid, y, base, idx, delta, hyper


class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), polled through NVML
    every few ms (the timed region is tens of ms; spawning nvidia-smi would see one sample)."""

    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index=0):
        self.rows = []
        self.stop = False
        self.index = index
        self.max_mhz = None
        self.nv = self.h = None
        try:                                        # NVML start-up takes ~100 ms: do it before the timed region
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv = nv
        except Exception as e:
            self.err = repr(e)
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample(self):
        mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        r = int(self.get_reasons(self.h))
        self.rows.append((mhz, [k for k, b in self.BITS.items() if r & b]))

    def run(self):
        if self.nv is not None:
            while not self.stop:
                try:
                    self.sample()
                except Exception:
                    break
                time.sleep(0.002)
            return
        try:                                        # NVML missing: one nvidia-smi sample
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            self.rows.append((a, []))
            self.max_mhz = b
        except Exception:
            self.rows.append((float("nan"), ["unavailable"]))

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [r[0] for r in self.rows if r[0] == r[0]]
        reasons = sorted({x for r in self.rows for x in r[1]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle's literal per-candidate slogdet loop
# --------------------------------------------------------------------------------------
def cpu_reference_setup():
    import oracle as O
    grid, y, base, idx, delta, hy = workload()
    th = O.Theta.from_values(hy["ls"], hy["os"], hy["noise"], hy["kind"])
    cov = O.OracleGP(th, "ref32").cov_mat(grid, add_likelihood_var=True)      # agent.py:90 (float32)
    static = np.zeros(len(grid), bool)
    static[base] = True
    mobile = np.zeros(len(grid), bool)
    return O, cov, static, mobile, idx


def cpu_score_sample(ctx, start, count):
    """`count` candidates scored exactly as Agent.best_path scores a path (agent.py:373-387)."""
    O, cov, static, mobile, idx = ctx
    t0 = time.perf_counter()
    for c in range(start, start + count):
        st = static.copy()
        st[idx[c, 0]] = True                      # the static reading of the set
        mo = mobile.copy()
        mo[idx[c, 1:]] = True
        O.set_entropy_literal(cov, st, mo, STATIC_STD, MOBILE_STD)
    return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count()
    ctx = cpu_reference_setup()
    t1 = cpu_score_sample(ctx, 0, 1)              # calibration (also the first warm-up)
    budget = 150.0
    per_step = max(1, int(budget / max(1e-3, t1) / max(1, args.steps + args.warmup)))
    per_step = min(per_step, 64)
    pos = 1
    for _ in range(args.warmup):
        cpu_score_sample(ctx, pos, per_step)
        pos += per_step
    t = 0.0
    for _ in range(args.steps):
        t += cpu_score_sample(ctx, pos, per_step)
        pos += per_step
    value = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 slogdet over f32 kernel matrix", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d candidate sets per step x %d steps of the 65536 (literal fancy-index + "
                                   "np.linalg.slogdet of a 4104x4104 matrix per set, OpenBLAS threads)" % (per_step, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict():
    return {"workload": "configs[2]: 65536 candidate sets (k=8: 1 static + 7 mobile) per GPU vs the N=4096 factor of a "
                        "128x128 mixture-of-Gaussians field (n=16384 locations, d=2, RBF, ls=8, s2=1, noise=1e-2)",
            "candidates_per_gpu": N_CAND, "set_size": K_SET, "n_train": N_BASE, "n_locations": FIELD * FIELD,
            "l2_policy": "inputs larger than L2: each step streams the 537 MB W^T matrix (126 MB L2)",
            "parallelism": "candidates sharded, factor replicated, per-rank winners exchanged through NVLink peer-memory "
                           "mailboxes from the argmax kernel (algp_b200.dist.sharded_best; NCCL all-gather as the fall-back)"}


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def i8_surviving_work(torch, km, lm, S, MT, NT, kchunks):
    """The MMA work that survives the occupancy masks of the INT8 digit variance GEMM (csrc/i8.cu plan_word /
    issue_sparse, restated on the masks): for every (128-row tile of K(X*,X), 64-row tile of Linv, 32-byte k chunk) inside
    the triangular k-range, the A planes [pmin, S-1-qmin] meet the B planes [qmin, min(S-p, qmax+1)).  Returns
    (surviving chunks, in-range chunks, MMA instructions, accumulator columns issued)."""
    mld = (kchunks + 7) // 8 * 8
    a = km.view(-1)[: MT * mld].view(MT, mld)[:, :kchunks].to(torch.int32)
    b = lm.view(-1)[: NT * mld].view(NT, mld)[:, :kchunks].to(torch.int32)

    def lowest(v):
        out = torch.full_like(v, 99)
        for p in range(S - 1, -1, -1):
            out = torch.where(((v >> p) & 1) > 0, torch.full_like(v, p), out)
        return out

    def highest(v):
        out = torch.full_like(v, -1)
        for p in range(S):
            out = torch.where(((v >> p) & 1) > 0, torch.full_like(v, p), out)
        return out
    pmin, qmin, qmax = lowest(a), lowest(b), highest(b)
    kend = (torch.arange(NT, device=b.device) * 64 + 64) // 32
    inrange = torch.arange(kchunks, device=b.device)[None, :] < kend[:, None]
    qmin = torch.where(inrange, qmin, torch.full_like(qmin, 99))
    cntA = torch.stack([(pmin == p).sum(0) for p in range(S)], 0).double()
    cntB = torch.zeros((S, S, kchunks), dtype=torch.float64, device=b.device)
    for q0 in range(S):
        for q1 in range(S):
            cntB[q0, q1] = ((qmin == q0) & (qmax == q1)).sum(0)
    H = torch.einsum("pk,qrk->pqr", cntA, cntB).cpu().numpy()
    chunks = mmas = cols = 0.0
    for p in range(S):
        for q0 in range(S):
            for q1 in range(q0, S):
                c = H[p, q0, q1]
                if c == 0 or p + q0 > S - 1:
                    continue
                chunks += c
                for pa in range(p, S - q0):
                    nplanes = min(S - pa, q1 + 1) - q0
                    for part in ([nplanes] if nplanes <= 4 else [4, nplanes - 4]):
                        mmas += c
                        cols += c * part * 64
    return chunks, float(MT) * float(inrange.sum().item()), mmas, cols


def med_of(v):
    return float(np.median(v)) if len(v) else None


def fit_predict_bench(torch, engine, n_train, grid_side, reps, peak_hbm):
    """GP fit+predict ms (second BASELINE metric): kbuild -> potrf -> trtri -> alpha -> fused mean ->
    variance TRMM with row norms.  Inputs resident; CUDA events; median of `reps`."""
    from algp_b200 import _lib
    rng = np.random.default_rng(1)
    x = rng.uniform(0, grid_side, size=(n_train, 2))
    yy, xx = np.meshgrid(np.arange(grid_side), np.arange(grid_side), indexing="ij")
    xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
    y = np.sin(x[:, 0] / 9.0) + np.cos(x[:, 1] / 7.0) + rng.normal(0, 0.1, n_train)
    hy = engine.Hyper(np.log([grid_side / 16.0] * 2), 0.0, np.log(1e-2), "rbf")
    xd, xsd = engine.to_dev(x), engine.to_dev(xs)
    y0 = engine.to_dev(y - y.mean())
    var = engine.to_dev(np.full(n_train, STATIC_STD ** 2))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    names = ["kbuild_train", "potrf", "trtri", "solve", "kbuild_cross_mean", "variance_trmm", "variance_tf32"]
    times = {k: [] for k in names + ["total"]}
    M = xs.shape[0]
    Npad = max(128, engine.pad_to(n_train))
    Mpad = max(128, engine.pad_to(M))
    N_, Mp_ = float(Npad), float(Mpad)
    dev = xd.device
    # buffers live across repetitions: the timed region is kernels only (inputs resident in HBM)
    A = torch.empty((Npad, Npad), dtype=torch.float64, device=dev)
    Linv = torch.empty((Npad, Npad), dtype=torch.float64, device=dev)
    Ks = torch.empty((Mpad, Npad), dtype=torch.float64, device=dev)
    work = torch.empty(max(2, _lib.lib.algp_trtri_work_doubles(Npad)), dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.__enter__()
    for rep in range(reps + 1):
        marks = [ev() for _ in range(len(names) + 1)]
        # keep the GPU busy while the host enqueues the first stages, so that event intervals are
        # kernel time and not Python launch latency (the queue never runs dry afterwards)
        Ks[: min(Mpad, 8192)].zero_()
        marks[0].record()
        engine.kbuild(hy, xd, None, Npad, Npad, var, hy.noise, True, out=A)
        marks[1].record()
        _lib.call("algp_potrf", _lib.ptr(A), Npad, Npad, _lib.ptr(Linv), Npad, _lib.ptr(info), _lib.stream())
        marks[2].record()
        _lib.call("algp_trtri", _lib.ptr(A), Npad, Npad, _lib.ptr(Linv), Npad, _lib.ptr(work), 1, _lib.stream())
        marks[3].record()
        f = engine.GPFactor.__new__(engine.GPFactor)
        f.hyper, f.x, f.N, f.Npad, f.L, f.Linv, f.info = hy, xd, n_train, Npad, A, Linv, info
        alpha, beta = f.solve(y0)
        marks[4].record()
        _, part = engine.kbuild(hy, xsd, xd, Mpad, Npad, out=Ks, dot_vec=alpha)
        mu = engine.rowsum(part, 1.0, float(y.mean()), rows=M)
        marks[5].record()
        _, rn = f.whiten(Ks, want_V=False)
        v = engine.rowsum(rn, -1.0, hy.outputscale, None, rows=M)
        marks[6].record()
        marks[7].record()
        torch.cuda.synchronize()
        if int(info.item()) != 0:
            raise RuntimeError("fit_predict bench: matrix not positive definite")
        if rep == 0:
            continue           # warm-up
        for i, k in enumerate(names[:-1]):
            times[k].append(marks[i].elapsed_time(marks[i + 1]))
        times["total"].append(marks[0].elapsed_time(marks[6]))
    # comparator ("cuSOLVER potrf as a sanity ceiling"): torch.linalg.cholesky on the same padded matrix with
    # the cuSOLVER backend; library code, not on the product path
    cusolver_ms = None
    try:
        torch.backends.cuda.preferred_linalg_library("cusolver")
        Ac, _ = engine.kbuild(hy, xd, None, Npad, Npad, var, hy.noise, True)
        Asym = torch.tril(Ac) + torch.tril(Ac, -1).T
        del Ac
        torch.linalg.cholesky(Asym)
        torch.cuda.synchronize()
        cs = []
        for _ in range(max(2, reps)):
            t0, t1 = ev(), ev()
            t0.record()
            torch.linalg.cholesky(Asym)
            t1.record()
            torch.cuda.synchronize()
            cs.append(t0.elapsed_time(t1))
        cusolver_ms = float(np.median(cs))
        del Asym
    except Exception as e:
        cusolver_ms = repr(e)
    # TF32 mode of the variance step, timed in its own loop (a run uses one mode or the other; interleaving
    # would let the tensor-core power draw of this stage throttle the next repetition's fp64 factorisation):
    # split K(X*,X) and Linv into fp32 hi/lo planes + tcgen05 split-TF32 TRMM
    times["variance_tf32"] = []
    for rep in range(reps + 1):
        f._linv_tf32 = None
        t0, t1 = ev(), ev()
        t0.record()
        rn32 = f.whiten_norm_tf32(Ks)
        v32 = engine.rowsum(rn32, -1.0, hy.outputscale, None, rows=M)
        t1.record()
        torch.cuda.synchronize()
        if rep:
            times["variance_tf32"].append(t0.elapsed_time(t1))
        del rn32
    tf32_err = float((v32 - v).abs().max().item())
    # INT8 digit mode (fp64 tier), staged the way precision "i8" runs it through the API: from N = 2048 both point sets
    # are sorted along a Z curve (results do not depend on the order; far-apart tiles become all-zero digit tiles
    # that the GEMM skips), from N = 8192 the factor comes from the recursive digit factorisation.
    reorder = n_train >= engine.I8_REORDER_MIN
    use_i8_factor = Npad >= engine.I8_FACTOR_MIN
    for k in ("reorder_i8", "kbuild_train_i8", "factor_i8", "solve_i8", "kbuild_cross_mean_i8", "variance_i8"):
        times[k] = []
    L_ref, Linv_ref = torch.tril(A), Linv.clone()
    for rep in range(reps + 1):
        m8 = [ev() for _ in range(7)]
        Ks[: min(Mpad, 8192)].zero_()
        m8[0].record()
        if reorder:
            perm, lo, hi = engine.morton_perm(xd)
            tperm, _, _ = engine.morton_perm(xsd, lo, hi)
            x8, y8, xs8 = xd[perm].contiguous(), y0[perm].contiguous(), xsd[tperm].contiguous()
        else:
            x8, y8, xs8 = xd, y0, xsd
        m8[1].record()
        engine.kbuild(hy, x8, None, Npad, Npad, var, hy.noise, True, out=A)       # var is constant: no permutation needed
        m8[2].record()
        if use_i8_factor:
            engine.potrf_inv_i8(A, Linv, info)
        else:
            _lib.call("algp_potrf", _lib.ptr(A), Npad, Npad, _lib.ptr(Linv), Npad, _lib.ptr(info), _lib.stream())
            _lib.call("algp_trtri", _lib.ptr(A), Npad, Npad, _lib.ptr(Linv), Npad, _lib.ptr(work), 1, _lib.stream())
        m8[3].record()
        f8 = engine.GPFactor.__new__(engine.GPFactor)
        f8.hyper, f8.x, f8.N, f8.Npad, f8.L, f8.Linv, f8.info = hy, x8, n_train, Npad, A, Linv, info
        alpha8, _ = f8.solve(y8)
        m8[4].record()
        _, part8 = engine.kbuild(hy, xs8, x8, Mpad, Npad, out=Ks, dot_vec=alpha8)
        mu8 = engine.rowsum(part8, 1.0, float(y.mean()), rows=M)
        m8[5].record()
        rn8 = f8.whiten_norm_i8(Ks)
        v8 = engine.rowsum(rn8, -1.0, hy.outputscale, None, rows=M)
        m8[6].record()
        torch.cuda.synchronize()
        if rep:
            for i, k in enumerate(("reorder_i8", "kbuild_train_i8", "factor_i8", "solve_i8", "kbuild_cross_mean_i8", "variance_i8")):
                times[k].append(m8[i].elapsed_time(m8[i + 1]))
        del rn8
    # the faster 7-plane setting of the same step (49 bits below each row's scale: fp64 tier relative to the prior
    # scale only), on the operands of the last repetition
    times["variance_i8_7planes"] = []
    f8._linv_i8 = None
    for rep in range(reps + 1):
        t0, t1 = ev(), ev()
        t0.record()
        rn7 = f8.whiten_norm_i8(Ks, nslices=7)
        v7 = engine.rowsum(rn7, -1.0, hy.outputscale, None, rows=M)
        t1.record()
        torch.cuda.synchronize()
        if rep:
            times["variance_i8_7planes"].append(t0.elapsed_time(t1))
        del rn7
    f8._linv_i8 = None
    # the same product with the occupancy masks OFF (the dense digit-tile schedule: what a field whose length-scale is
    # comparable to its extent costs), and the MMA work that survives the masks on this field
    i8_dense_ms = i8_survive = None
    if reorder:
        try:
            S8 = engine.I8_SLICES
            dts = []
            for rep in range(2):
                t0, t1 = ev(), ev()
                t0.record()
                rnd = f8.whiten_norm_i8(Ks, nslices=S8, use_masks=False)
                t1.record()
                torch.cuda.synchronize()
                dts.append(t0.elapsed_time(t1))
                del rnd
            i8_dense_ms = float(min(dts))
            _, _, km8 = f8.split_i8(Ks, S8, 128, want_mask=True)
            _, _, lm8 = f8._linv_digits(S8)
            ch, inr, mmas, cols = i8_surviving_work(torch, km8, lm8, S8, Mpad // 128, Npad // 64, Npad // 32)
            i8_survive = {"surviving_chunks": ch, "in_range_chunks": inr, "mma_instructions": mmas,
                          "surviving_int8_ops": 2.0 * 128 * 32 * cols, "dense_int8_ops": S8 * (S8 + 1) / 2 * 2.0 * N_ * N_ * Mp_ / 2}
            del km8, lm8
            f8._linv_i8 = None
        except Exception as e:
            i8_survive = {"error": repr(e)}
    if reorder:
        inv = torch.empty_like(tperm)
        inv[tperm] = torch.arange(M, device=dev)
        v8, mu8, v7 = v8[inv], mu8[inv], v7[inv]
    i8_err = float((v8 - v).abs().max().item())
    i8_rel = float(((v8 - v).abs() / v).max().item())
    i7_err = float((v7 - v).abs().max().item())
    i7_rel = float(((v7 - v).abs() / v).max().item())
    tf32_rel = float(((v32 - v).abs() / v).max().item())
    i8_mu_err = float(((mu8 - mu).abs().max() / mu.abs().max()).item())
    # cond_2(K) estimate ("record cond(K) estimates in the output") from the fp64 factor: largest eigenvalue of
    # K = L L^T and of K^-1 = Linv^T Linv by 30 power iterations each (torch matrix-vector products: a bench statistic,
    # not a product path)
    cond_est = None
    try:
        g = torch.Generator(device=dev).manual_seed(0)
        vv = torch.randn(Npad, dtype=torch.float64, device=dev, generator=g)
        ww = vv.clone()
        lam_max = lam_inv = 0.0
        for _ in range(30):
            vv = torch.mv(L_ref, torch.mv(L_ref.T, vv))
            lam_max = float(vv.norm().item())
            vv /= lam_max
            ww = torch.mv(Linv_ref.T, torch.mv(Linv_ref, ww))
            lam_inv = float(ww.norm().item())
            ww /= lam_inv
        cond_est = {"lambda_max": lam_max, "lambda_min": 1.0 / lam_inv, "cond_2": lam_max * lam_inv,
                    "how": "30 power iterations on L L^T and on Linv^T Linv (padded identity rows included)"}
        del vv, ww
    except Exception as e:
        cond_est = {"error": repr(e)}
    if reorder:
        i8_factor_err = None                        # a different (permuted) factor: compared through mean / variance
    else:
        i8_factor_err = {"max_abs_dL": float((torch.tril(A) - L_ref).abs().max().item()),
                         "max_abs_dLinv_rel": float(((Linv - Linv_ref).abs().max() / Linv_ref.abs().max()).item())}
    del L_ref, Linv_ref
    sampler.__exit__()
    # end to end through the reference-facing call with HOST arrays (utils.py:293): H2D of x / y / var / grid,
    # kernel build + factor + solve + variance, D2H of mean and variance, fresh factor every call
    import algp_b200
    e2e = {}
    gp = algp_b200.GPR(kernel_params={'type': 'rbf'})
    var_h = np.full(n_train, STATIC_STD ** 2)
    gp.reset(x, y, var_h)
    with torch.no_grad():
        gp.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(hy.log_ls).view(1, 1, -1))
        gp.model.kernel_covar_module.log_outputscale.fill_(hy.log_os)
        gp.likelihood.log_noise.fill_(hy.log_noise)
    ref64 = None
    for mode in ("fp64", "i8", "i8fast", "tf32"):
        gp.precision = mode
        ts = []
        for rep in range(reps + 1):
            gp._cache.clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mu_h, var_out = algp_b200.predictive_distribution(gp, x, y, xs, var_h, return_var=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        e2e[mode] = float(np.median(ts[1:]))
        if mode == "fp64":
            ref64 = (mu_h.copy(), var_out.copy())
        else:
            # accuracy of each mode through the same public call, against the fp64 (DMMA) mode
            e2e[mode + "_max_abs_var_diff_vs_fp64"] = float(np.abs(var_out - ref64[1]).max())
            e2e[mode + "_max_rel_var_diff_vs_fp64"] = float((np.abs(var_out - ref64[1]) / ref64[1]).max())
            e2e[mode + "_max_abs_mean_diff_over_max_abs_mean"] = float(np.abs(mu_h - ref64[0]).max() / np.abs(ref64[0]).max())
    # the same call with the Matern-1.5 kernel at the same theta ("also a Matern-1.5 run with the same theta")
    gpm = algp_b200.GPR(kernel_params={'type': 'matern'})
    gpm.reset(x, y, var_h)
    with torch.no_grad():
        gpm.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(hy.log_ls).view(1, 1, -1))
        gpm.model.kernel_covar_module.log_outputscale.fill_(hy.log_os)
        gpm.likelihood.log_noise.fill_(hy.log_noise)
    for mode in ("fp64", "i8"):
        gpm.precision = mode
        ts = []
        for rep in range(min(reps, 2) + 1):
            gpm._cache.clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mu_m, var_m = algp_b200.predictive_distribution(gpm, x, y, xs, var_h, return_var=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        e2e["matern_" + mode] = float(np.median(ts[1:]))
        if mode == "fp64":
            refm = var_m.copy()
        else:
            e2e["matern_i8_max_abs_var_diff_vs_fp64"] = float(np.abs(var_m - refm).max())
    e2e["h2d_bytes"] = int(x.nbytes + y.nbytes + var_h.nbytes + xs.nbytes)
    e2e["d2h_bytes"] = int(mu_h.nbytes + var_out.nbytes)
    med = {k: (float(np.median(v)) if len(v) else None) for k, v in times.items()}
    # the Z-order stage is a dozen small torch kernels: its event interval is dominated by host-side allocator stalls in
    # some repetitions (cudaMalloc of the temporaries after the 8 GB buffers of the previous stage were freed), which are
    # not device work -- take the fastest repetition for it
    if times.get("reorder_i8"):
        med["reorder_i8"] = float(np.min(times["reorder_i8"]))
    N = float(max(128, engine.pad_to(n_train)))
    Mp = float(max(128, engine.pad_to(M)))
    out = {"n_train": n_train, "n_test": M, "ms": med["total"],
           "ms_i8_mode": sum(med[k] for k in ("reorder_i8", "kbuild_train_i8", "factor_i8", "solve_i8",
                                                "kbuild_cross_mean_i8", "variance_i8")),
           "i8_mode": {"z_order": bool(reorder), "digit_factorisation": bool(use_i8_factor),
                       "factor_digit_planes": engine.I8_FACTOR_SLICES, "factor_base_rows": engine.I8_FACTOR_BASE,
                       "factor_vs_dmma": i8_factor_err, "max_abs_mean_diff_over_max_abs_mean_vs_fp64": i8_mu_err,
                       "factor_fp64_equiv_tflops": 2 * N ** 3 / 3 / med["factor_i8"] / 1e9},
           "ms_tf32_mode": med["total"] - med["variance_trmm"] + med["variance_tf32"],
           "e2e_ms_host_arrays": e2e, "clocks": sampler.summary(),
           "tf32_max_abs_var_diff_vs_fp64": tf32_err, "i8_max_abs_var_diff_vs_fp64": i8_err,
           "variance_error_vs_fp64_dmma": {
               "what": "kernel level, same factor: max |dvar| (prior variance s^2 = 1) and max |dvar| / var over the grid",
               "i8_%d_planes" % engine.I8_SLICES: {"abs": i8_err, "rel": i8_rel},
               "i8_7_planes": {"abs": i7_err, "rel": i7_rel, "ms": med_of(times["variance_i8_7planes"])},
               "split_tf32_kernel": {"abs": tf32_err, "rel": tf32_rel,
                                     "note": "precision='tf32' uses this kernel up to N = %d and the 4-plane digit GEMM beyond" % engine.TF32_MAX_N}},
           "tolerance_contract": "fp64 tier: |dmean| <= 1e-9 max|mean|, |dvar| <= 1e-9 s^2 (prior scale; SURVEY 7 'variance "
                                 "cancellation'), and with the default 8 digit planes also |dvar| <= 1e-9 var (relative); "
                                 "1e-4 tier: 1e-4 of the same scales",
           "i8_digit_planes": engine.I8_SLICES, "ms_by_stage": med, "cond_K_estimate": cond_est,
           "var_min": float(v.min().item()), "var_max": float(v.max().item()),
           "rooflines": {
               "kbuild_train": {"bound": "hbm", "achieved": 8 * N * N / med["kbuild_train"] / 1e6, "peak": peak_hbm, "unit": "GB/s"},
               "kbuild_cross_mean": {"bound": "hbm", "achieved": 8 * N * Mp / med["kbuild_cross_mean"] / 1e6, "peak": peak_hbm, "unit": "GB/s"},
               "solve": {"bound": "hbm", "achieved": 8 * N * N / med["solve"] / 1e6, "peak": peak_hbm, "unit": "GB/s",
                         "what": "alpha = Linv^T (Linv y): two triangular gemv passes, 8 N^2 / 2 bytes each (SURVEY 8d)"},
               "potrf": {"bound": "fp64 tensor (DMMA)", "achieved": N ** 3 / 3 / med["potrf"] / 1e9, "unit": "TFLOP/s",
                         "ms": med["potrf"], "cusolver_potrf_ms_same_matrix": cusolver_ms,
                         "note": "algp_potrf also inverts every 128 x 128 diagonal block (needed by the panel GEMM and by trtri)"},
               "trtri": {"bound": "fp64 tensor (DMMA)", "achieved": N ** 3 / 3 / med["trtri"] / 1e9, "unit": "TFLOP/s"},
               "variance_trmm": {"bound": "fp64 tensor (DMMA)", "achieved": N * N * Mp / med["variance_trmm"] / 1e9, "unit": "TFLOP/s"},
               "variance_i8": {"bound": "int8 tensor (tcgen05 kind::i8), S(S+1)/2 exact digit GEMMs per product, incl. the digit split "
                                        "passes; digit tiles that are all zero (Z-ordered points) are skipped, so the dense-equivalent "
                                        "rate can exceed the dense GEMM peak",
                               "achieved": engine.I8_SLICES * (engine.I8_SLICES + 1) / 2 * N * N * Mp / med["variance_i8"] / 1e9,
                               "unit": "TOP/s(int8)", "effective_fp64_equiv_tflops": N * N * Mp / med["variance_i8"] / 1e9},
               "variance_i8_surviving": None if not i8_survive or "error" in i8_survive else {
                   "bound": "int8 tensor (tcgen05 kind::i8) on the MMA work that SURVIVES the occupancy masks (GEMM + split passes timed)",
                   "achieved": i8_survive["surviving_int8_ops"] / med["variance_i8"] / 1e9, "unit": "TOP/s(int8)",
                   "surviving_fraction_of_dense_schedule": i8_survive["surviving_int8_ops"] / i8_survive["dense_int8_ops"],
                   "surviving_chunks": i8_survive["surviving_chunks"], "in_range_chunks": i8_survive["in_range_chunks"],
                   "mma_instructions": i8_survive["mma_instructions"],
                   "dense_digit_tiles_ms_masks_off": i8_dense_ms},
               "variance_tf32": {"bound": "tf32 tensor (tcgen05), 3 MMAs per product, incl. the hi/lo split passes",
                                 "achieved": 3 * N * N * Mp / med["variance_tf32"] / 1e9, "unit": "TFLOP/s(tf32)",
                                 "effective_fp64_equiv_tflops": N * N * Mp / med["variance_tf32"] / 1e9},
           }}
    out["rooflines"] = {k: v for k, v in out["rooflines"].items() if v is not None}
    for r in out["rooflines"].values():
        if "peak" in r:
            r["frac"] = r["achieved"] / r["peak"]
    return out


def episode_bench(torch, engine, side=200, n_pilot=1024, acquisitions=500, per_batch=4, n_paths=256, path_len=16,
                  distributed=False, dist=None, rank=0, world=1, dev=None, oracle_batches=0):
    """BASELINE configs[4]: active-sampling episode on a side x side field: `acquisitions` static picks in batches
    of `per_batch` (greedy, rank-1 appends), each batch followed by scoring `n_paths` candidate paths of `path_len`
    mobile readings and committing the winner.  Path enumeration (env.py) is the planner's job: paths here are
    synthetic straight runs starting at the batch's picks.  distributed=True (N > 1): every rank holds the replicated
    posterior state, scores a contiguous block of the paths (algp_b200.dist.sharded_best, winners over the NVLink
    mailboxes) and applies the same commits; rank 0 then repeats the episode alone and the chosen indices must agree."""
    from algp_b200.episode import run_episode
    hyper, grid, static, mobile, path_fn = episode_problem(engine, side, n_pilot, n_paths, path_len)
    n = len(grid)
    batches = acquisitions // per_batch

    Xd = engine.to_dev(grid, device=dev)
    run_episode(hyper, Xd, static, mobile, STATIC_STD, MOBILE_STD, 2, per_batch, path_fn, distributed=distributed)   # warm-up
    if distributed:
        dist.barrier()
    res = run_episode(hyper, Xd, static, mobile, STATIC_STD, MOBILE_STD, batches, per_batch, path_fn, distributed=distributed)
    ms_b, ms_a = res["ms_per_batch"], res["ms_per_acquisition"]
    out = {"field": "%dx%d" % (side, side), "n_locations": n, "pilot_samples": n_pilot, "acquisitions": batches * per_batch,
           "paths_per_batch": n_paths, "path_len": path_len, "mobile_committed": int(res["mobile"].sum()),
           "entropy_first_last": [res["H"][0], res["H"][-1]]}
    if distributed:
        tt = torch.tensor([ms_b, ms_a], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_b, ms_a = float(tt[0].item()), float(tt[1].item())
        out["gpus"] = world
        out["paths_per_gpu"] = n_paths // world
        picks, paths = res["picks"], res["best_paths"]
        del res
        # chosen-index agreement: every rank against rank 0, and rank 0's sharded run against its own single-GPU run
        mine = torch.tensor([p for b in picks for p in b] + paths, dtype=torch.int64, device=dev)
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([int(torch.equal(mine, ref))], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["all_ranks_chose_the_same_indices"] = bool(same.item())
        if rank == 0:
            solo = run_episode(hyper, Xd, static, mobile, STATIC_STD, MOBILE_STD, batches, per_batch, path_fn, distributed=False)
            out["same_picks_and_paths_as_one_gpu"] = bool(solo["picks"] == picks and solo["best_paths"] == paths)
            out["ms_per_acquisition_one_gpu_same_run"] = solo["ms_per_acquisition"]
            out["entropy_last_one_gpu"] = solo["H"][-1]
        dist.barrier()
    out["ms_per_acquisition"] = ms_a
    out["ms_per_batch"] = ms_b
    if not distributed and oracle_batches > 0:
        # chosen-index agreement with the fp64 oracle at this scale (BASELINE.md, configs[4] row): the first batches
        # against oracle.LeanEpisode, the restructured episode that never forms an n x n matrix
        import oracle as O
        t0 = time.perf_counter()
        th = O.Theta(hyper.log_ls.copy(), hyper.log_os, hyper.log_noise, "rbf")
        ep = O.LeanEpisode(th, grid, static, mobile, STATIC_STD, MOBILE_STD)
        same, dH = True, 0.0
        for b in range(min(oracle_batches, batches)):
            picks = ep.greedy(per_batch)
            paths = path_fn(b, picks)
            sc = ep.score_paths(paths)
            best = int(np.argmax(sc))
            same = same and picks == res["picks"][b] and best == res["best_paths"][b]
            dH = max(dH, abs(float(sc[best]) - res["H"][b]) / abs(float(sc[best])))
            ep.commit_path(paths[best], sc[best])
        out["oracle_agreement"] = {"batches_checked": min(oracle_batches, batches), "same_picks_and_paths": bool(same),
                                   "max_rel_entropy_diff": dH, "cpu_s": time.perf_counter() - t0}
    return out


def episode_problem(engine, side=200, n_pilot=1024, n_paths=256, path_len=16):
    """The synthetic episode of configs[4]: field grid, pilot flags, hyper-parameters and the per-batch candidate paths
    (straight runs of `path_len` locations starting at the batch's picks).  Shared with tests/test_gpu_fullsize.py."""
    from algp_b200.utils import generate_gaussian_data
    grid, _ = generate_gaussian_data(side, side, seed=1)
    grid = grid.astype(np.float64)
    n = len(grid)
    rng = np.random.default_rng(3)
    static = np.zeros(n, bool)
    static[rng.choice(n, n_pilot, replace=False)] = True
    mobile = np.zeros(n, bool)
    hyper = engine.Hyper(np.log([side / 16.0] * 2), 0.0, np.log(1e-2), "rbf")

    def path_fn(b, picks):
        prng = np.random.default_rng(1000 + b)
        starts = np.array(picks)[prng.integers(0, len(picks), n_paths)]
        r, c = starts // side, starts % side
        dr, dc = prng.integers(-1, 2, n_paths), prng.integers(-1, 2, n_paths)
        steps = np.arange(1, path_len + 1)[None, :]
        rr = np.clip(r[:, None] + dr[:, None] * steps, 0, side - 1)
        cc = np.clip(c[:, None] + dc[:, None] * steps, 0, side - 1)
        return (rr * side + cc).astype(np.int32)

    return hyper, grid, static, mobile, path_fn


def patched_reference_style_agent():
    """A class with the reference Agent's sample bookkeeping (agent.py:47-82) and NO hot-path methods of its own,
    run through algp_b200.patch(): the drop-in route of INTEGRATION.md (the reference's own agent.py cannot be imported
    on the GPU box, so its bookkeeping is restated here)."""
    import algp_b200

    class RefAgent(object):
        def __init__(self, env, gp, static_std, mobile_std, rng):
            self.env, self.gp, self.static_std, self.mobile_std, self.rng = env, gp, static_std, mobile_std, rng
            self.criterion = 'entropy'
            self.reset()

        def reset(self):
            self.collected = {'ind': [], 'std': [], 'y': []}
            self.static_data = [[] for _ in range(self.env.num_samples)]
            self.mobile_data = [[] for _ in range(self.env.num_samples)]

        def _add_samples(self, indices, stds):
            all_y = [None] * len(indices)
            for i in range(len(indices)):
                idx = indices[i]
                if idx == -1:
                    continue
                y = float(self.rng.normal(0.5, stds[i]))            # env.collect_samples: a noisy reading
                all_y[i] = y
                (self.static_data if stds[i] == self.static_std else self.mobile_data)[idx].append(y)
            self.collected['ind'] += list(indices)
            self.collected['std'] += list(stds)
            self.collected['y'] += all_y

    return algp_b200.patch(RefAgent)


def default_scale_bench(torch, engine, cpu_sample=24):
    """BASELINE configs[0] scale (the reference's own CPU-runnable case): a planning step on an 860-location field with
    d = 6 inputs and a Matern-1.5 kernel -- enumerate the candidate paths through 5 waypoints on the reference's
    planning graph (tests/golden/ref_paths.npz, case 4: 945 paths of up to 55 mobile readings), then
    Agent.greedy(4) + Agent.best_path(paths) through the reference-facing API with host lists.  The CPU side is the
    oracle's literal reference loops (agent.py:295-403) on a sample of the paths, scaled to all of them."""
    import algp_b200
    from algp_b200 import paths as P
    import oracle as O
    gpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "ref_paths.npz")
    if not os.path.exists(gpath):
        return {"skipped": "tests/golden/ref_paths.npz not found"}
    g = np.load(gpath)
    k = 4
    c = {name: g["c%d_%s" % (k, name)] for name in ("rc", "adj_ptr", "adj", "eptr", "eidx", "start", "heading", "waypoints",
                                                     "least_cost", "slack")}
    nodes = [tuple(r) for r in c["rc"].tolist()]
    enum = lambda: P.enumerate_paths_arrays(nodes, c["rc"], c["adj_ptr"], c["adj"], c["eptr"], c["eidx"], int(c["start"]),
                                            tuple(c["heading"].tolist()), c["waypoints"], float(c["least_cost"]), float(c["slack"]))
    enum()
    t0 = time.perf_counter()
    ps = enum()
    slots = ps.slots()
    enum_ms = (time.perf_counter() - t0) * 1e3
    lists = ps.indices()
    n, d = 860, 6
    rng = np.random.default_rng(5)
    cells = rng.choice(30 * 30, n, replace=False)
    X = np.column_stack([cells // 30, (cells % 30) * 2.0, rng.integers(0, 2, (n, 4))]).astype(np.float64)
    static = rng.choice(n, 300, replace=False)
    mobile = rng.choice(np.setdiff1d(np.arange(n), static), 60, replace=False)

    class Env(object):
        pass
    env = Env()
    env.X, env.test_X, env.num_samples = X, X[:40], n
    ag = patched_reference_style_agent()(env, None, STATIC_STD, MOBILE_STD, np.random.default_rng(6))
    ag._add_samples(list(static), [STATIC_STD] * len(static))
    ag._add_samples(list(mobile) + list(mobile), [MOBILE_STD] * (2 * len(mobile)))
    ls, os_, noise = [3.0, 6.0, 1.5, 1.5, 1.5, 1.5], 1.0, 1e-2
    ind, yv, var = ag.get_sampled_dataset()
    ag.gp = algp_b200.GPR(kernel_params={'type': 'matern'})
    ag.gp.reset(X[ind], yv, var)
    with torch.no_grad():
        ag.gp.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(np.log(ls)).view(1, 1, -1))
        ag.gp.model.kernel_covar_module.log_outputscale.fill_(float(np.log(os_)))
        ag.gp.likelihood.log_noise.fill_(float(np.log(noise)))
    ag._post_update()

    def step(paths_arg):
        ag._hot_state = None                       # a fresh base set every planning step, as in run.py
        picks = ag.greedy(4)
        return picks, ag.best_path(paths_arg, picks)

    step(lists)
    torch.cuda.synchronize()
    ts_l, ts_a = [], []
    for _ in range(5):
        t0 = time.perf_counter(); picks, best = step(lists); torch.cuda.synchronize(); ts_l.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter(); picks_a, best_a = step(slots); torch.cuda.synchronize(); ts_a.append((time.perf_counter() - t0) * 1e3)
    # CPU: the literal loops of the reference on the same inputs
    th = O.Theta.from_values(ls, os_, noise, "matern")
    cov = O.OracleGP(th, "ref32").cov_mat(X, add_likelihood_var=True)
    st = np.zeros(n, bool); st[static] = True
    mo = np.zeros(n, bool); mo[mobile] = True
    t0 = time.perf_counter()
    cpu_picks = O.greedy_literal(cov, st, mo, STATIC_STD, MOBILE_STD, 4)
    cpu_greedy_ms = (time.perf_counter() - t0) * 1e3
    sample = lists[:cpu_sample]
    t0 = time.perf_counter()
    O.best_path_literal(cov, st, mo, STATIC_STD, MOBILE_STD, sample, cpu_picks)
    cpu_paths_ms = (time.perf_counter() - t0) * 1e3 * len(lists) / len(sample)
    # three more iterations of the run_ipp body (agent.py:133-201 with update=False) on the same patched agent: the
    # picks and the winning path's readings are added through the reference's _add_samples, so every later greedy /
    # best_path extends the cached posterior state instead of re-factorising
    ipp_ms, same_state = [], []
    ag._hot_state = None
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pk = ag.greedy(4)
        bp = ag.best_path(lists, pk)
        torch.cuda.synchronize()
        ipp_ms.append((time.perf_counter() - t0) * 1e3)
        same_state.append(id(ag._hot_state["state"]))
        seq = [int(j) for j in lists[bp]] + [int(j) for j in pk]
        ag._add_samples(seq, [MOBILE_STD] * len(lists[bp]) + [STATIC_STD] * len(pk))
    return {"field_locations": n, "d": d, "kernel": "matern", "n_base": int(len(ind)), "paths": len(lists),
            "agent": "reference-style class (agent.py:47-82 bookkeeping) through algp_b200.patch()",
            "run_ipp_iterations_ms": [round(t, 3) for t in ipp_ms],
            "posterior_state_extended_not_refactorised": bool(len(set(same_state)) == 1),
            "longest_path": int(slots.shape[1]), "path_enumeration_ms": enum_ms,
            "greedy4_plus_best_path_ms": {"lists": float(np.median(ts_l)), "slot_array": float(np.median(ts_a))},
            "picks": [int(p) for p in picks], "best_path": int(best), "same_choice_array_form": bool(best == best_a and picks == picks_a),
            "cpu_literal_loops_ms": {"greedy4": cpu_greedy_ms, "best_path_scaled": cpu_paths_ms, "cores": os.cpu_count(),
                                     "sample": "%d of %d paths" % (len(sample), len(lists)),
                                     "same_picks": bool([int(p) for p in cpu_picks] == [int(p) for p in picks])}}


def fit_loop_bench(torch, engine, no_cpu=False):
    """GPR.fit as a measured loop (reference models.py:145-158: 200 Adam iterations on the exact MLL): the public
    call, end to end -- per iteration one kernel build, factorisation, inverse, alpha, MLL gradient pass, a (d+5)-double
    D2H and the host Adam / ReduceLROnPlateau step over d + 2 scalars; x, y, var and all N^2 buffers stay resident
    (algp_b200.mll.MLLWorkspace).  Config E (n = 645, d = 6, Matern, lr 0.1 as arguments.py:11) and N = 4096 (d = 2,
    RBF).  Beside each: the oracle's NumPy loss + gradient evaluation on the host cores."""
    import io
    import contextlib
    import algp_b200
    out = {}
    for name, n, d, kind, iters in (("config_E_n645_d6_matern", 645, 6, "matern", 200), ("n4096_d2_rbf", 4096, 2, "rbf", 200)):
        rng = np.random.default_rng(7)
        if d == 6:
            cells = rng.choice(30 * 30, n, replace=False)
            x = np.column_stack([cells // 30, (cells % 30) * 2.0, rng.integers(0, 2, (n, 4))]).astype(np.float64)
        else:
            x = rng.uniform(0, 64, size=(n, d))
        y = np.sin(x[:, 0] / 5.0) + np.cos(x[:, 1] / 7.0) + rng.normal(0, 0.1, n)
        var = np.full(n, STATIC_STD ** 2)
        gp = algp_b200.GPR(lr=0.1, max_iterations=iters, kernel_params={'type': kind})
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            gp.max_iter = 3
            gp.fit(x, y, var)                                     # warm-up (module load, allocator)
            gp.max_iter = iters
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gp.fit(x, y, var)
            torch.cuda.synchronize()
            fit_s = time.perf_counter() - t0
        row = {"n": n, "d": d, "kernel": kind, "iterations": iters, "fit_ms": fit_s * 1e3, "ms_per_iteration": fit_s * 1e3 / iters,
               "log": buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else None}
        if not no_cpu:
            import oracle as O
            hy = gp.hyper()
            th = O.Theta(hy.log_ls.copy(), hy.log_os, hy.log_noise, hy.kind_name)
            reps = 3 if n < 1000 else 1
            t0 = time.perf_counter()
            for _ in range(reps):
                O.mll_loss(th, x, y, var)
                O.mll_loss_grad(th, x, y, var)
            row["cpu_oracle_ms_per_iteration"] = (time.perf_counter() - t0) * 1e3 / reps
            row["cpu_cores"] = os.cpu_count()
        out[name] = row
    return out


def mi_greedy_bench(torch, engine, picks=4):
    """Mutual-information greedy at scale (agent.py:330-339): 128 x 128 field (n = 16384), 4096 static
    samples, `picks` MI picks.  The two complement factorizations (12288^2 and 16384^2) are built ONCE per greedy call
    and follow every pick by rank-1 updates of their inverse diagonals and log-dets (MIContext.commit); round 1
    re-factorised both per pick, which is what `ms_context_build` costs."""
    grid, _ = field_grid()
    n = len(grid)
    rng = np.random.default_rng(1)
    base = np.sort(rng.choice(n, N_BASE, replace=False))
    hyper = engine.Hyper(np.log([FIELD / 16.0] * 2), 0.0, np.log(1e-2), "rbf")
    pi = np.zeros(n)
    pi[base] = 1.0 / STATIC_STD ** 2
    Xd = engine.to_dev(grid)
    state = engine.PosteriorState(hyper, Xd, base, pi, is_static=pi > 0, capacity=picks + 16, cov_mode="never")
    d = 1.0 / STATIC_STD ** 2
    ev = lambda: torch.cuda.Event(enable_timing=True)
    engine.MIContext(hyper, Xd, pi).check()                          # warm-up
    e0, e1 = ev(), ev()
    e0.record()
    ctx = engine.MIContext(hyper, Xd, pi)
    e1.record()
    torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    pair = torch.empty(2, dtype=torch.int64, device=Xd.device)
    chosen, per_pick = [], []
    for _ in range(picks):
        a, b = ev(), ev()
        a.record()
        ent_a = state.H_base_dev + state.greedy_utilities(d)
        ut = ctx.greedy_utilities(ent_a, STATIC_STD, MOBILE_STD).contiguous()
        state.argmax(ut, 0, out=pair)
        state.append(pair[1:2], d, mark_static=True)
        j = int(pair[1].item())
        state.H_base_dev.copy_(ent_a[j:j + 1])
        state._H_base = None
        ctx.commit(j, STATIC_STD, MOBILE_STD)
        b.record()
        torch.cuda.synchronize()
        per_pick.append(a.elapsed_time(b))
        chosen.append(j)
        pi[j] += d
    # the maintained quantities against a context factored from scratch on the final flags
    fresh = engine.MIContext(hyper, Xd, pi)
    err = {"logdet_unsampled": float((ctx.ld2 - fresh.ld2).abs().item()), "logdet_all": float((ctx.ld3 - fresh.ld3).abs().item()),
           "inv_diag_all_rel": float(((ctx.diag3 - fresh.diag3).abs() / fresh.diag3).max().item())}
    return {"n_locations": n, "n_static": N_BASE, "picks": chosen, "ms_context_build": build_ms,
            "ms_per_pick_rank1_maintenance": [round(t, 3) for t in per_pick],
            "ms_per_pick_refactorising_round1": build_ms, "after_%d_picks_vs_refactorisation" % picks: err}


def sharded_fit_predict_bench(torch, dist, engine, dev, world, n_train=16384, grid_side=256, reps=2):
    """GP fit+predict ms at N = 16384 over the 256 x 256 grid with the test rows sharded over the ranks: kernel build +
    factor + alpha on rank 0, NCCL broadcast of Linv, every rank its block of rows, all-gather.  Max over ranks."""
    from algp_b200 import dist as adist
    rng = np.random.default_rng(1)
    x = rng.uniform(0, grid_side, size=(n_train, 2))
    yy, xx = np.meshgrid(np.arange(grid_side), np.arange(grid_side), indexing="ij")
    xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
    y = np.sin(x[:, 0] / 9.0) + np.cos(x[:, 1] / 7.0) + rng.normal(0, 0.1, n_train)
    hy = engine.Hyper(np.log([grid_side / 16.0] * 2), 0.0, np.log(1e-2), "rbf")
    xd, xsd = engine.to_dev(x, device=dev), engine.to_dev(xs, device=dev)
    y0 = engine.to_dev(y - y.mean(), device=dev)
    var = engine.to_dev(np.full(n_train, STATIC_STD ** 2), device=dev)
    ts = []
    for rep in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mu, v = adist.sharded_mean_var(hy, xd, var, y0, float(y.mean()), xsd, precision="i8")
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rep:
            ts.append(float(tt.item()))
    return {"n_train": n_train, "n_test": int(xs.shape[0]), "mode": "INT8 digit mode, test rows sharded over %d ranks" % world,
            "ms": float(np.median(ts)), "var_min": float(v.min().item()), "var_max": float(v.max().item()),
            "linv_broadcast_bytes": int(8 * n_train * n_train)}


def dgemm_peak(torch, n=8192, reps=3, sustained_s=1.5):
    """cuBLAS fp64 GEMM rate on this box: the DMMA roofline denominator (not in MEASURED_PEAKS.json).
    Returns (burst, sustained): best single call, and the mean over ~sustained_s of back-to-back calls
    (what a kernel inside a long fp64 phase can expect under the power cap)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    calls = max(4, int(sustained_s * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    sustained = 2.0 * n ** 3 * calls / e0.elapsed_time(e1) / 1e9
    return 2.0 * n ** 3 / best / 1e9, sustained


def tf32_gemm_peak(torch, n=8192, reps=3):
    """cuBLAS TF32 GEMM rate (fp32 operands, allow_tf32): the tcgen05 roofline denominator."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, dtype=torch.float32, device="cuda")
        b = torch.randn(n, n, dtype=torch.float32, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / best / 1e9
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def int8_gemm_peak(torch, n=8192, reps=3):
    """cuBLASLt int8 x int8 -> int32 GEMM throughput (TOP/s) through torch._int_mm: the roofline denominator of the
    INT8 digit mode (the driver's MEASURED_PEAKS.json has no int8 figure).  None if the library path is unavailable."""
    try:
        a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
        torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / best / 1e9
    except Exception:
        return None


def resident_cov_bench(torch, engine, algp_b200, hyper, Xd, base, pi0, is_static, idx, idx_d, delta_d, H_base,
                       stream_scores, stream_ms, steps, make_agent):
    """Config B with the posterior covariance P of the base set resident in HBM ("If the build
    precomputes P (allowed), the amortised SYRK cost is reported and included in a second end-to-end figure").
    A candidate then gathers 36 entries of P instead of streaming 8 rows of W^T.  Timed with 8 rotating candidate
    batches so the gathered sectors of one step (~75 MB) are not the L2 contents left by the previous one."""
    from algp_b200 import _lib
    dev = Xd.device
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    states = {}
    for prec in ("fp64", "i8"):
        st = engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, precision=prec, cov_mode="never")
        st.build_cov()
        st.drop_cov()
        torch.cuda.synchronize()
        e0.record()
        st.build_cov()
        e1.record()
        torch.cuda.synchronize()
        out["build_ms_" + prec] = e0.elapsed_time(e1)
        states[prec] = st
    out["build_i8_max_abs_dP"] = float((torch.tril(states["i8"].P) - torch.tril(states["fp64"].P)).abs().max().item())
    del states["i8"]
    st = states["fp64"]
    n = st.n
    scores = torch.empty(N_CAND, dtype=torch.float64, device=dev)
    st.score_sets(idx_d, delta_d, H_base=H_base, out=scores)
    out["max_abs_score_diff_vs_streaming"] = float((scores - stream_scores).abs().max().item())
    out["same_winner_as_streaming"] = bool(int(scores.argmax().item()) == int(stream_scores.argmax().item()))
    # rotating batches: uniform draws over the non-base locations (a repeated location inside a set counts once)
    rest = torch.nonzero(torch.as_tensor(pi0 == 0, device=dev)).view(-1).to(torch.int32)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    batches = [rest[torch.randint(0, rest.numel(), (N_CAND, K_SET), device=dev, generator=g)].contiguous() for _ in range(8)]
    pair = torch.empty(2, dtype=torch.int64, device=dev)
    def step(i):
        st.score_sets(batches[i % 8], delta_d, H_base=H_base, out=scores)
        st.argmax(scores, idx_offset=0, out=pair)
    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0.record()
    for i in range(steps):
        kev[i][0].record()
        st.score_sets(batches[i % 8], delta_d, H_base=H_base, out=scores)
        kev[i][1].record()
        st.argmax(scores, idx_offset=0, out=pair)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    out["ms_per_step"] = ms
    out["value"] = N_CAND / (ms / 1e3)
    out["unit"] = UNIT
    out["kernel"] = "score_cov_k8_kernel"
    out["kernel_ms"] = kms
    # algorithmic bytes per candidate: 36 gathered doubles + 8 int32 slots + 8 deltas + the score
    algo = N_CAND * (36 * 8 + K_SET * 4 + K_SET * 8 + 8)
    out["roofline"] = {"bound": "hbm", "achieved": algo / (kms / 1e3) / 1e9, "unit": "GB/s",
                       "algorithmic_bytes_per_launch": algo,
                       "note": "random 8-byte gathers: every entry costs a 32-byte DRAM sector, so at most a quarter of "
                               "the HBM rate is reachable on algorithmic bytes"}
    bms = out["build_ms_fp64"]
    out["incl_build_amortised"] = {
        "what": "candidates/s with the one-off P build (fp64 DMMA SYRK; i8 in brackets) charged to B batches of 65536",
        "batches_1": N_CAND / ((bms + ms) / 1e3), "batches_10": 10 * N_CAND / ((bms + 10 * ms) / 1e3),
        "batches_100": 100 * N_CAND / ((bms + 100 * ms) / 1e3),
        "batches_10_i8_build": 10 * N_CAND / ((out["build_ms_i8"] + 10 * ms) / 1e3),
        "streaming_for_comparison": N_CAND / (stream_ms / 1e3),
        "break_even_batches_fp64_build": bms / max(stream_ms - ms, 1e-9),
        "break_even_batches_i8_build": out["build_ms_i8"] / max(stream_ms - ms, 1e-9)}
    # commits keep P alive: 4 greedy picks (rank-1 appends) + the 16 mobile readings of a path (one block append), then
    # the new columns are folded into P by rank-k downdates and the next batch is scored from the same P
    try:
        st_c = engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, capacity=64, cov_mode="always")
        st_c.score_sets(batches[0], delta_d, H_base=H_base, out=scores)
        torch.cuda.synchronize()
        d_s, d_m = 1.0 / STATIC_STD ** 2, 1.0 / MOBILE_STD ** 2
        e0.record()
        picks = st_c.greedy(4, d_s)
        path = [int(v) for v in rest[:4096:256].cpu().tolist() if int(v) not in picks][:16]
        st_c.append_block(path, d_m, mark_static=False)
        e1.record()
        torch.cuda.synchronize()
        commit_ms = e0.elapsed_time(e1)
        e0.record()
        st_c._sync_cov()
        e1.record()
        torch.cuda.synchronize()
        sync_ms = e0.elapsed_time(e1)
        e0.record()
        sc_p = st_c.score_sets(batches[1], delta_d).clone()
        e1.record()
        torch.cuda.synchronize()
        score_ms = e0.elapsed_time(e1)
        st_c.cov_mode, keepP, st_c.P = "never", st_c.P, None
        sc_s = st_c.score_sets(batches[1], delta_d)
        out["across_commits"] = {
            "what": "4 greedy picks + one 16-reading block append on a state with P resident; P is kept and brought up to date "
                    "by algp_cov_downdate (one pass over the lower triangle per <= %d new columns: %d for these 20) instead of a "
                    "rebuild" % (_lib.lib.algp_cov_downdate_max_cols(), -(-20 // _lib.lib.algp_cov_downdate_max_cols())),
            "commit_ms": commit_ms, "downdate_ms": sync_ms, "rebuild_ms_for_comparison": bms, "score_from_P_ms": score_ms,
            "max_abs_score_diff_vs_streaming_after_commits": float((sc_p - sc_s).abs().max().item()),
            "downdate_gbs": -(-20 // _lib.lib.algp_cov_downdate_max_cols()) * 2 * 8.0 * st_c.n_pad * (st_c.n_pad + 64) / 2
                            / (sync_ms / 1e3) / 1e9}
        del st_c, keepP
    except Exception as e:
        out["across_commits_error"] = repr(e)
    del st, states
    # through the reference-facing call with the default policy (cov_mode "auto"): the state streams until the
    # streamed work would have paid for the build, builds P inside one call, and gathers from then on
    ag = make_agent("auto")
    per_call = []
    for _ in range(40):
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        ag.best_path(idx, [])
        torch.cuda.synchronize()
        per_call.append((time.perf_counter() - w0) * 1e3)
    out["e2e_auto_policy"] = {"api": "Agent.best_path(ndarray[65536,8], []) x 40 on an unchanged state, host arrays",
                              "ms_per_call": [round(t, 3) for t in per_call],
                              "steady_value": N_CAND / (float(np.median(per_call[-6:])) / 1e3), "unit": UNIT,
                              "value_over_all_40_calls": 40 * N_CAND / (sum(per_call) / 1e3)}
    return out


def timed_steps(torch, dist, world, dev, fn, steps, warmup, sampler=None):
    """W untimed + K timed calls of fn(i) bracketed by barrier + synchronize on both sides; CUDA events on the
    current stream; the MAX over ranks of the elapsed time in ms."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn(None)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.__enter__()
    barrier()
    t0.record()
    last = None
    for i in range(steps):
        last = fn(i)
    t1.record()
    barrier()
    if sampler is not None:
        sampler.__exit__()
    ms = t0.elapsed_time(t1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms, last


def decode_winner(torch, res):
    """(score, index) from what algp_b200.dist.sharded_best returned (device tensor or host tuple)."""
    if torch.is_tensor(res):
        h = res.cpu()
        if h.numel() > 2 and int(h[2]) != 0:
            raise RuntimeError("winner exchange timed out")
        return float(h[0:1].view(torch.float64).item()), int(h[1])
    return float(res[0]), int(res[1])


def l2_probe(torch, mbytes=(32, 64)):
    """L2 -> SM delivery rate measured in this run (algp_probe_l2_read): an L2-resident buffer read 40 times with the
    scoring kernel's load instruction, every CTA a different part.  TB/s per buffer size."""
    from algp_b200 import _lib
    out = {}
    sink = torch.zeros(1, dtype=torch.float64, device="cuda")
    for mb in mbytes:
        buf = torch.ones(mb << 20, dtype=torch.uint8, device="cuda")
        best = 0.0
        for ctas in (2, 4, 8):
            _lib.call("algp_probe_l2_read", _lib.ptr(buf), buf.numel(), 4, ctas, _lib.ptr(sink), _lib.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("algp_probe_l2_read", _lib.ptr(buf), buf.numel(), 40, ctas, _lib.ptr(sink), _lib.stream())
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 40.0 * buf.numel() / (e0.elapsed_time(e1) / 1e3) / 1e12)
        out["%dMB" % mb] = best
        del buf
    return out


def ncu_dram_traffic(kernel_substr, files=("r02_prof_score_summary.csv", "r01_prof_score_summary.csv"), reduce="mean"):
    """dram__bytes_read.sum + dram__bytes_write.sum of a kernel from the committed ncu summaries (profiles/*.csv, written
    by scripts/ncu_summary.py from an `ncu --set full` capture): (bytes | None, file | None, ms | None).  reduce="mean":
    per launch; reduce="sum": over all captured launches (the chunk launches of ONE chunked scoring call), with
    extra = {"launches", "dmma_pipe_pct" (time-weighted), "l2_hit_pct"}."""
    import csv
    units = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    for name in files:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, unit = rows[0], rows[1]
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        vals = [float(r[ir]) * units.get(unit[ir], 1.0) + float(r[iw]) * units.get(unit[iw], 1.0)
                for r in rows[2:] if r and kernel_substr in r[0]]
        if vals:
            ms = None
            sel = [r for r in rows[2:] if r and kernel_substr in r[0]]
            if "gpu__time_duration.sum" in hdr:
                it = hdr.index("gpu__time_duration.sum")
                tu = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit[it], 1.0)
                tms = [float(r[it]) * tu for r in sel]
                ms = float(np.sum(tms) if reduce == "sum" else np.mean(tms))
            if reduce != "sum":
                return float(np.mean(vals)), "profiles/" + name, ms
            extra = {"launches": len(vals)}
            for key, col in (("dmma_pipe_pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
                             ("l2_hit_pct", "lts__t_sector_hit_rate.pct")):
                if col in hdr and ms:
                    ic = hdr.index(col)
                    extra[key] = float(np.sum([float(r[ic]) * t for r, t in zip(sel, tms)]) / np.sum(tms))
            return float(np.sum(vals)), "profiles/" + name, ms, extra
    return (None, None, None) if reduce != "sum" else (None, None, None, None)


def strong_scaling_bench(torch, dist, adist, engine, state, idx0_d, delta0_d, H_base, rank, world, dev, steps, warmup):
    """configs[2] read literally: ONE batch of 65 536 candidate sets sharded over the ranks (8 192 per GPU at N = 8)
    through algp_b200.dist.sharded_best (winners exchanged over the NVLink mailboxes), next to the same batch on rank 0
    alone in the same run, and to the same sharded step with the NCCL all-gather carrying the winners."""
    out = {"scaling": "strong", "total_candidate_sets": N_CAND, "sets_per_gpu": N_CAND // world}
    ms, last = timed_steps(torch, dist, world, dev,
                           lambda i: adist.sharded_best(state, idx0_d, delta0_d, H_base=H_base, return_device=True),
                           steps, warmup)
    out["ms_per_step"] = ms / steps
    out["value"] = N_CAND * steps / (ms / 1e3)
    out["unit"] = UNIT
    out["winner"] = decode_winner(torch, last)[1]
    out["exchange"] = "nvlink mailboxes (csrc/p2p.cu)" if adist.peer_exchange() is not None else "nccl all-gather"
    # comparator: the round-1 step (two argmax kernels + NCCL all-gather of 16 bytes per rank)
    lo, hi = adist.shard_range(N_CAND, rank, world)
    scores = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    pair = torch.empty(2, dtype=torch.int64, device=dev)
    gathered = torch.empty(2 * world, dtype=torch.int64, device=dev)

    def nccl_step(i):
        state.score_sets(idx0_d[lo:hi], delta0_d[lo:hi], H_base=H_base, out=scores)
        state.argmax(scores, idx_offset=lo, out=pair)
        dist.all_gather_into_tensor(gathered, pair)
    ms_n, _ = timed_steps(torch, dist, world, dev, nccl_step, steps, warmup)
    out["ms_per_step_nccl_allgather"] = ms_n / steps
    # the same batch on ONE GPU in the same run (rank 0 works, the others wait at the barrier)
    full = torch.empty(N_CAND, dtype=torch.float64, device=dev)

    def single_step(i):
        if rank == 0:
            state.score_sets(idx0_d, delta0_d, H_base=H_base, out=full)
            state.argmax(full, idx_offset=0, out=pair)
    ms_1, _ = timed_steps(torch, dist, world, dev, single_step, steps, warmup)
    out["ms_per_step_one_gpu_same_run"] = ms_1 / steps
    out["efficiency_vs_one_gpu_same_run"] = (ms_1 / steps) / (world * ms / steps)
    if rank == 0:
        out["winner_one_gpu"] = int(pair[1].item())
    return out


def resident_sharded_bench(torch, dist, adist, engine, hyper, Xd, base, pi0, is_static, idx_all_d, delta_all_d, idx0_d,
                           delta0_d, H_base, world, dev, steps, warmup):
    """The resident-covariance path at N GPUs: every rank builds P once (replicated, like the factor) and scores its
    block by gathers; weak (65 536 sets per GPU) and strong (65 536 in total)."""
    st = engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, cov_mode="always")
    out = {}
    for name, ix, dl, total in (("weak", idx_all_d, delta_all_d, world * N_CAND), ("strong", idx0_d, delta0_d, N_CAND)):
        ms, last = timed_steps(torch, dist, world, dev,
                               lambda i: adist.sharded_best(st, ix, dl, H_base=H_base, return_device=True), steps, warmup)
        out[name] = {"ms_per_step": ms / steps, "value": total * steps / (ms / 1e3), "unit": UNIT,
                     "winner": decode_winner(torch, last)[1]}
    del st
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import algp_b200
    from algp_b200 import _lib, engine
    from algp_b200 import dist as adist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    peak_hbm, peak_src = load_peaks()
    steps, warmup = args.steps, max(3, args.warmup)

    # rank r's block of the weak-scaling run is the 65 536 sets of seed 2 + r; block 0 is configs[2]'s batch
    grid, y, base, idx0, delta, hy = workload(seed_sets=2)
    n = len(grid)
    rest = np.setdiff1d(np.arange(n), base)
    idx_all = np.concatenate([idx0] + [candidate_sets(rest, 2 + r) for r in range(1, world)])
    delta_all = np.tile(delta, (world, 1))
    hyper = engine.Hyper(np.log(hy["ls"]), np.log(hy["os"]), np.log(hy["noise"]), hy["kind"])
    pi0 = np.zeros(n)
    pi0[base] = 1.0 / STATIC_STD ** 2
    is_static = (pi0 > 0)

    # ---- one-off: factor + W build (reported separately, SURVEY.md 8d metric 1) ----
    Xd = engine.to_dev(grid, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static)       # warm-up (module load, attributes)
    torch.cuda.synchronize()
    e0.record()
    state = engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, cov_mode="never")   # headline: streaming
    e1.record()
    torch.cuda.synchronize()
    setup_ms = e0.elapsed_time(e1)
    setup_ms_i8 = setup_i8_dW = None
    if world == 1:
        # the same one-off build with precision "i8" (W^T through the INT8 digit GEMM)
        engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, precision="i8")
        torch.cuda.synchronize()
        e0.record()
        state8 = engine.PosteriorState(hyper, Xd, base, pi0, is_static=is_static, precision="i8")
        e1.record()
        torch.cuda.synchronize()
        setup_ms_i8 = e0.elapsed_time(e1)
        setup_i8_dW = float((state8.Wt - state.Wt).abs().max().item())
        del state8
    H_base = state.H_base

    idx_all_d = engine.to_dev(idx_all, dtype=torch.int32, device=dev)
    delta_all_d = engine.to_dev(delta_all, device=dev)
    idx0_d, delta0_d = idx_all_d[:N_CAND], delta_all_d[:N_CAND]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]

    # ---- headline: the product's sharded scoring step (algp_b200.dist.sharded_best): this rank's block of the global
    # candidate array scored against the replicated factor, np.argmax, winners exchanged between the GPUs ----
    def step(i):
        return adist.sharded_best(state, idx_all_d, delta_all_d, H_base=H_base, return_device=True,
                                  events=None if i is None else kev[i])

    for _ in range(warmup):
        step(None)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    clk = ClockSampler(local_rank)
    ms, last = timed_steps(torch, dist, world, dev, step, steps, 0, sampler=clk)
    gpu_launches = _lib.launch_count - launches0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    win_score, win = decode_winner(torch, last)
    value = world * N_CAND * steps / (ms / 1e3)
    exchange = None
    if world > 1:
        exchange = "nvlink mailboxes (csrc/p2p.cu)" if adist.peer_exchange() is not None else "nccl all-gather"

    # ---- end to end through the reference-facing call: Agent.best_path with HOST arrays ----
    class Env(object):
        pass

    def make_agent(cov_mode):
        env = Env()
        env.X, env.test_X, env.num_samples = grid, grid[:16], n
        ag = algp_b200.Agent.__new__(algp_b200.Agent)
        ag.env, ag.static_std, ag.mobile_std, ag.criterion = env, STATIC_STD, MOBILE_STD, 'entropy'
        ag.cov_mode = cov_mode
        ag.shard_candidates = world > 1          # one process per GPU: every rank scores its block of the paths
        ag.static_data = [[0.0] if s else [] for s in is_static]
        ag.mobile_data = [[] for _ in range(n)]
        ag.collected = {'ind': list(base), 'std': [STATIC_STD] * len(base), 'y': [0.0] * len(base)}
        ag.gp = algp_b200.GPR(kernel_params={'type': hy["kind"]})
        ag.gp.reset(grid[base], y[base], np.full(len(base), STATIC_STD ** 2))
        with torch.no_grad():
            ag.gp.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(np.log(hy["ls"])).view(1, 1, -1))
            ag.gp.model.kernel_covar_module.log_outputscale.fill_(float(np.log(hy["os"])))
            ag.gp.likelihood.log_noise.fill_(float(np.log(hy["noise"])))
        ag._post_update()
        return ag
    ag = make_agent("never")         # headline end-to-end number: the streaming path, as in `value`
    # the reference call scores paths of mobile readings on top of static waypoints: here every set is 8 mobile slots
    # and no new static waypoint (same kernel, same bytes per candidate).  At N > 1 the call receives the GLOBAL host
    # array; every rank copies and scores its own block and all ranks return the same winner.
    e2e_steps = max(3, min(steps, 10))
    for _ in range(3):
        ag.best_path(idx_all, [])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_win = ag.best_path(idx_all, [])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * N_CAND * e2e_steps / e2e_s

    extra = {}
    cpu_base = None
    if world > 1:
        for name, fn in (
                ("strong_scaling", lambda: strong_scaling_bench(torch, dist, adist, engine, state, idx0_d, delta0_d, H_base, rank,
                                                                 world, dev, steps, warmup)),
                ("resident_cov_sharded", lambda: resident_sharded_bench(torch, dist, adist, engine, hyper, Xd, base, pi0, is_static,
                                                                        idx_all_d, delta_all_d, idx0_d, delta0_d, H_base, world,
                                                                        dev, steps, warmup)),
                ("episode", lambda: episode_bench(torch, engine, distributed=True, dist=dist, rank=rank, world=world, dev=dev))):
            try:
                extra[name] = fn()
            except Exception as e:
                extra[name + "_error"] = repr(e)
                if "timed out" in repr(e) or "NCCL" in repr(e):
                    raise
    if world > 1 and not args.skip_large:
        # second metric at N > 1: the factorisation stays on rank 0, its inverse factor is broadcast, the 256 x 256
        # grid of test rows is sharded over the ranks (algp_b200.dist.sharded_mean_var), INT8 digit mode
        try:
            extra["fit_predict_sharded"] = sharded_fit_predict_bench(torch, dist, engine, dev, world)
        except Exception as e:
            extra["fit_predict_sharded_error"] = repr(e)
    # the fit+predict / fit-loop / CPU legs explain the N=1 line; at N>1 only the sharded metrics are measured
    # (the other ranks would idle at the barrier while rank 0 ran them)
    fp64_peak = None
    if rank == 0 and world == 1:
        try:
            extra["l2_to_sm_probe_tbs"] = l2_probe(torch)
        except Exception as e:
            extra["l2_probe_error"] = repr(e)
        try:
            extra["resident_cov"] = resident_cov_bench(torch, engine, algp_b200, hyper, Xd, base, pi0, is_static, idx0, idx0_d,
                                                       delta0_d, H_base, state.score_sets(idx0_d, delta0_d, H_base=H_base).clone(),
                                                       ms / steps, steps, make_agent)
        except Exception as e:
            extra["resident_cov_error"] = repr(e)
        try:
            fp64_peak, fp64_sustained = dgemm_peak(torch)
            extra["fp64_gemm_peak_tflops_cublas_8192"] = fp64_peak
            extra["fp64_gemm_sustained_tflops_cublas_8192"] = fp64_sustained
            fits = [fit_predict_bench(torch, engine, 4096, 64, 5, peak_hbm)]
            if not args.skip_large:
                fits.append(fit_predict_bench(torch, engine, 16384, 256, 2, peak_hbm))
            tf32_peak = tf32_gemm_peak(torch)
            extra["tf32_gemm_peak_tflops_cublas_8192"] = tf32_peak
            i8_peak = int8_gemm_peak(torch)
            extra["int8_gemm_peak_tops_cublaslt_8192"] = i8_peak
            for f in fits:
                for r in f["rooflines"].values():
                    if r["unit"] == "TFLOP/s":
                        # burst figure for the short N=4096 leg, sustained for the long N=16384 leg
                        pk = fp64_peak if f["n_train"] <= 4096 else fp64_sustained
                        r["peak"] = pk
                        r["peak_kind"] = "cuBLAS dgemm burst" if f["n_train"] <= 4096 else "cuBLAS dgemm sustained"
                        r["frac"] = r["achieved"] / pk
                    elif r["unit"].startswith("TFLOP/s(tf32)"):
                        r["peak"] = tf32_peak
                        r["frac"] = r["achieved"] / tf32_peak
                    elif r["unit"].startswith("TOP/s(int8)"):
                        r["peak"] = i8_peak if i8_peak else 4500.0
                        r["peak_kind"] = "cuBLASLt int8 GEMM burst" if i8_peak else "nominal dense int8"
                        r["frac"] = r["achieved"] / r["peak"]
            extra["fit_predict"] = fits
        except Exception as e:       # the headline line must still print
            extra["fit_predict_error"] = repr(e)
        try:
            extra["episode"] = episode_bench(torch, engine, oracle_batches=0 if args.no_cpu else 6)
        except Exception as e:
            extra["episode_error"] = repr(e)
        try:
            extra["mi_greedy"] = mi_greedy_bench(torch, engine)
        except Exception as e:
            extra["mi_greedy_error"] = repr(e)
        torch.cuda.empty_cache()
        try:
            extra["fit_loop"] = fit_loop_bench(torch, engine, no_cpu=args.no_cpu)
        except Exception as e:
            extra["fit_loop_error"] = repr(e)
        if not args.no_cpu:
            try:
                extra["default_scale_step"] = default_scale_bench(torch, engine)
            except Exception as e:
                extra["default_scale_error"] = repr(e)
        if not args.no_cpu:
            ctx = cpu_reference_setup()
            cpu_score_sample(ctx, 0, 1)
            cnt = 12
            t = cpu_score_sample(ctx, 1, cnt)
            # second metric on the CPU: the reference's predictive_distribution (float32 inv + two GEMMs,
            # utils.py:296-308) at N=4096 / 64x64 grid, once
            try:
                import oracle as O
                rng = np.random.default_rng(1)
                xc = rng.uniform(0, 64, size=(4096, 2))
                yyc, xxc = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
                xsc = np.stack([yyc.ravel(), xxc.ravel()], 1).astype(np.float64)
                yc = np.sin(xc[:, 0] / 9.0) + np.cos(xc[:, 1] / 7.0) + rng.normal(0, 0.1, 4096)
                thc = O.Theta.from_values([4.0, 4.0], 1.0, 1e-2, "rbf")
                t0 = time.perf_counter()
                O.predictive_distribution(O.OracleGP(thc, "ref32"), xc, yc, xsc, np.full(4096, STATIC_STD ** 2), return_var=True)
                extra["cpu_fit_predict_n4096"] = {"ms": (time.perf_counter() - t0) * 1e3, "cores": os.cpu_count(), "kind": "port",
                                                  "what": "oracle ref32 predictive_distribution(return_var), N=4096, M=4096"}
            except Exception as e:
                extra["cpu_fit_predict_error"] = repr(e)
            cpu_base = {"value": cnt / t, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": "%d of the 65536 candidate sets, literal reference loop (fancy-index + "
                                  "np.linalg.slogdet of a 4104x4104 matrix per set, OpenBLAS threads)" % cnt}
    if rank == 0:
        algo_bytes = 8.0 * N_BASE * K_SET * N_CAND           # s*N*k per candidate
        achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
        # the timed step scores its batch through algp_score_sets_tiled, which streams 17.2 GB and therefore runs the
        # persistent sweep (one launch: 768-column chunks, partial Grams in shared memory): the ncu capture of exactly
        # that launch (profiles/r02_prof_score_resident_summary.csv) is the traffic of a step; the captures of the
        # earlier forms (one launch per 1024-column chunk; the plain single launch) are kept beside it
        launches_per_step = max(1, _lib.lib.algp_score_sets_tiled_launches(K_SET, N_CAND, state.ncols, state.n_pad))
        traffic, traffic_src, ncu_ms, ncu_extra = ncu_dram_traffic("score_sets_k8_resident_kernel",
                                                                  files=("r02_prof_score_resident_summary.csv",), reduce="sum")
        chunked = ncu_dram_traffic("score_sets_k8_kernel", files=("r02_prof_score_tiled_summary.csv",), reduce="sum")
        single = ncu_dram_traffic("score_sets_k8_kernel")
        kernel_name = "score_sets_k8_resident_kernel"
        if traffic is None or (ncu_extra or {}).get("launches") != launches_per_step:
            traffic, traffic_src, ncu_ms = single            # no capture of the shipped form for this shape
            ncu_extra = None
            kernel_name = "score_sets_k8_kernel"
        gram_flops = 2.0 * N_BASE * K_SET * K_SET * N_CAND   # the full 8 x 8 Gram the DMMA tiles compute
        roof = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak_hbm,
                "unit": "GB/s", "frac": achieved / peak_hbm, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": kernel_ms, "kernel_launches_per_step": launches_per_step,
                "algorithmic_bytes_per_launch": algo_bytes,
                "note": "one step = ONE algp_score_sets_tiled call = %d launch(es) of the persistent sweep (column chunks sized "
                        "for the L2, the candidates' partial Grams resident in shared memory).  `achieved` is SURVEY 8d's "
                        "no-reuse byte count over the kernel time: it exceeds the DRAM peak because the L2 serves most of the "
                        "rows (sector hits in the capture: `resources.dram.l2_sector_hit_pct_ncu`), so `frac` is not a "
                        "fraction of a roof.  The resources the kernel actually loads are under `resources`; none is saturated: "
                        "all algorithmic bytes cross the L2 -> SM path exactly once, and the DMMA pipe is the busiest unit." % launches_per_step}
        res = {}
        if traffic is not None:
            res["dram"] = {"bytes_per_step_ncu": traffic, "achieved_gbs": traffic / (kernel_ms / 1e3) / 1e9, "peak_gbs": peak_hbm,
                           "frac": traffic / (kernel_ms / 1e3) / 1e9 / peak_hbm,
                           "ncu_kernel_ms": ncu_ms,
                           "frac_within_the_ncu_capture": (traffic / (ncu_ms / 1e3) / 1e9 / peak_hbm) if ncu_ms else None,
                           "note": "dram__bytes_read + write of the step's launches in the committed ncu capture (each replayed "
                                   "from flushed caches) over the kernel time of this run; over the capture's own time: "
                                   "`frac_within_the_ncu_capture`"}
            if ncu_extra:
                res["dram"]["l2_sector_hit_pct_ncu"] = ncu_extra.get("l2_hit_pct")
            if chunked[0] is not None and chunked[1] != traffic_src:
                res["dram"]["per_chunk_launches_capture"] = {
                    "bytes": chunked[0], "ncu_kernel_ms": chunked[2], "source": chunked[1],
                    "l2_sector_hit_pct_ncu": (chunked[3] or {}).get("l2_hit_pct"),
                    "note": "the form shipped before the persistent sweep: 4 launches of 1024-column chunks, accumulator "
                            "fragments parked in global memory between them"}
            if single[0] is not None and single[1] != traffic_src:
                res["dram"]["plain_single_launch_capture"] = {
                    "bytes": single[0], "ncu_kernel_ms": single[2], "source": single[1],
                    "frac_within_the_ncu_capture": single[0] / (single[2] / 1e3) / 1e9 / peak_hbm if single[2] else None,
                    "note": "algp_score_sets as ONE launch over all 4096 columns (what an isolated call ran before the chunked "
                            "form): DRAM-bound there, 1.55 ms"}
        probe = extra.get("l2_to_sm_probe_tbs")
        if probe:
            pk = max(probe.values())
            res["l2_to_sm"] = {"achieved_tbs": achieved / 1e3, "peak_tbs_measured_in_run": pk, "frac": achieved / 1e3 / pk}
        if fp64_peak:
            res["fp64_tensor_dmma"] = {"achieved_tflops": gram_flops / (kernel_ms / 1e3) / 1e12, "peak_tflops_cublas_dgemm": fp64_peak,
                                       "frac": gram_flops / (kernel_ms / 1e3) / 1e12 / fp64_peak}
            if ncu_extra and ncu_extra.get("dmma_pipe_pct") is not None:
                res["fp64_tensor_dmma"]["pipe_active_pct_ncu"] = ncu_extra["dmma_pipe_pct"]
        if res:
            roof["resources"] = res
            roof["binding"] = max(res, key=lambda k: res[k]["frac"])
            # the fraction of the busiest measured resource, all three over the kernel time of THIS run
            roof["frac_of_binding_roof"] = res[roof["binding"]]["frac"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(idx0.nbytes), "d2h_bytes_per_step": 16 if world == 1 else 24,
                    "api": "algp_b200.Agent.best_path(ndarray[%d,8], []) with host arrays%s" %
                           (world * N_CAND, "" if world == 1 else ", shard_candidates=True: every rank copies and scores its block"),
                    "steps": e2e_steps, "winner": int(e2e_win)},
            "gpu_launches": int(gpu_launches),
            "roofline": roof,
            "setup_ms_factor_and_W": setup_ms, "setup_ms_factor_and_W_i8": setup_ms_i8, "setup_i8_max_abs_dW": setup_i8_dW,
            "winner": win, "winner_score": win_score, "H_base": H_base,
        }
        if exchange is not None:
            line["winner_exchange"] = exchange
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        adist.shutdown()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-large", action="store_true", help="skip the N=16384 fit+predict measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
