"""Candidate scoring sharded over the GPUs of one box.

The candidates (greedy: the n locations; best_path: the P paths; config B: the 65 536 sets) are
independent given the base factor, so each rank holds a replicated factor, scores one contiguous
block of candidates, and the only exchange is one all-gather of a 16-byte (score, global index)
pair per rank followed by a local reduction with np.argmax's first-maximum rule (agent.py:349,402):
highest score, ties to the lowest global index.  Contiguous blocks keep that rule trivial.
There is no data-path collective: the training-set factorisation stays on each GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` candidates owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_pairs(values, indices):
    """First-max over per-rank winners: max value, then min global index.  Ranks with an empty shard
    report -inf."""
    values = np.asarray(values, dtype=np.float64)
    indices = np.asarray(indices, dtype=np.int64)
    order = np.lexsort((indices, -values))
    return float(values[order[0]]), int(indices[order[0]])


def allgather_argmax(pair, group=None):
    """pair: int64[2] tensor {bit pattern of the float64 score, global index} as algp_argmax writes it
    (CUDA tensor under NCCL, CPU tensor under gloo).  Returns (score, index) identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        g = pair.reshape(1, 2)
    else:
        out = torch.empty(2 * world, dtype=torch.int64, device=pair.device)
        dist.all_gather_into_tensor(out, pair.contiguous(), group=group)
        g = out.view(world, 2)
    g = g.cpu()
    vals = g[:, 0].contiguous().view(torch.float64).numpy()
    return combine_pairs(vals, g[:, 1].numpy())


class PeerExchange(object):
    """The ranks' winner mailboxes mapped into each other over cudaIpc (csrc/p2p.cu): the exchange of the per-rank
    {score, global index} pairs becomes plain NVLink stores from the last argmax kernel instead of an NCCL all-gather
    launch.  One instance per process group; every rank must call argmax() the same number of times."""

    def __init__(self, group=None, timeout_ms=20000.0):
        import ctypes as C
        from . import _lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.timeout_ms = float(timeout_ms)
        self.epoch = 0
        self._lib = _lib
        self.dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = _lib.lib.algp_p2p_mailbox_bytes(self.world)
        if nbytes <= 0:
            raise RuntimeError("PeerExchange supports up to 16 ranks")
        # every rank takes part in both collectives below whatever happens locally: a rank whose export or mapping
        # failed reports it through the MIN all-reduce instead of leaving the others waiting
        self.local = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        ok = 1
        rc = _lib.lib.algp_p2p_create(nbytes, C.byref(self.local), C.cast(handle, C.c_void_p))
        if rc != 0:
            ok = 0
            self.local = None
            self.error = "algp_p2p_create: " + _lib.lib.algp_last_cuda_error().decode()
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy().reshape(self.world, 64)
        self.opened = []
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(self.local.value if self.local is not None else 0)
                continue
            if not allh[r].any():                 # that rank could not export its mailbox
                ok = 0
                ptrs.append(0)
                continue
            buf = (C.c_ubyte * 64)(*allh[r].tolist())
            p = C.c_void_p()
            rc = _lib.lib.algp_p2p_open(C.cast(buf, C.c_void_p), C.byref(p))
            if rc != 0:
                ok = 0
                self.error = _lib.lib.algp_last_cuda_error().decode()
                ptrs.append(0)
            else:
                self.opened.append(p)
                ptrs.append(p.value)
        # all ranks or none: a rank that could not map a peer must not leave the others waiting for its stores
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError("cudaIpc peer mapping unavailable: " + getattr(self, "error", "on another rank"))
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.dev)
        self.out3 = torch.empty(3, dtype=torch.int64, device=self.dev)
        self.work = torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device=self.dev)

    def argmax(self, x, idx_offset=0):
        """x: this rank's score block (1-D float64 CUDA tensor, or None / empty for an empty shard).
        Returns the device tensor {score bits, global index, status}, identical on every rank."""
        self.epoch += 1
        n = 0 if x is None else int(x.shape[0])
        self._lib.call("algp_argmax_exchange", self._lib.ptr(x) if n else None, n, int(idx_offset), self._lib.ptr(self.work),
                       self._lib.ptr(self.peers), self.rank, self.world, self.epoch, self.timeout_ms,
                       self._lib.ptr(self.out3), self._lib.stream())
        return self.out3

    def result(self, out3=None):
        """(score, index) on the host; raises if a peer's pair did not arrive (24-byte D2H, synchronises)."""
        h = (self.out3 if out3 is None else out3).cpu()
        if int(h[2]) != 0:
            raise RuntimeError("PeerExchange: a peer's winner did not arrive within %.0f ms (rank %d, epoch %d)"
                               % (self.timeout_ms, self.rank, self.epoch))
        return float(h[0:1].view(torch.float64).item()), int(h[1])

    def close(self):
        for p in getattr(self, "opened", []):
            self._lib.lib.algp_p2p_close(p)
        self.opened = []
        if getattr(self, "local", None) is not None and self.local.value:
            self._lib.lib.algp_p2p_destroy(self.local)
            self.local = None


_exchanges = {}


def peer_exchange(group=None):
    """The PeerExchange of `group` (created on first use, collectively), or None when peer mapping is unavailable
    (then the NCCL all-gather carries the pairs).  ALGP_P2P=0 forces the NCCL path."""
    import os
    key = id(group) if group is not None else None
    if key not in _exchanges:
        ex = None
        if os.environ.get("ALGP_P2P", "1") != "0":
            try:
                ex = PeerExchange(group)
            except RuntimeError as e:
                import warnings
                warnings.warn("algp_b200.dist: falling back to the NCCL all-gather for the winner exchange: %s" % e)
        _exchanges[key] = ex
    return _exchanges[key]


def shutdown():
    """Unmap / free the mailboxes (call before destroy_process_group)."""
    for ex in _exchanges.values():
        if ex is not None:
            ex.close()
    _exchanges.clear()


def pack_pair(value, index, device="cpu"):
    """Host-side constructor of the {score bits, index} pair (tests, empty shards)."""
    v = torch.tensor([value], dtype=torch.float64).view(torch.int64)
    return torch.cat([v, torch.tensor([index], dtype=torch.int64)]).to(device)


def sharded_best(state, idx_all, delta_all=None, delta_scalar=0.0, skip=None, group=None, H_base=None, return_device=False,
                 events=None, check=False):
    """Score this rank's block of the global candidate array idx_all [B,k] (host int32 array, or an int32 device
    tensor replicated on every rank) against the replicated PosteriorState and return the global (score, index) winner
    on every rank.  The per-rank winners travel through the NVLink mailboxes of PeerExchange (NCCL all-gather when peer
    mapping is unavailable, gloo on CPU tensors).  return_device=True (PeerExchange only) returns the device tensor
    {score bits, index, status} without synchronising.  events = (start, end) CUDA events recorded around the scoring
    kernel (bench.py's per-kernel timing).  check=True (host idx_all only): this rank's block is range-checked on
    the device before it is scored and IndexError is raised, after the exchange, on the ranks whose block held a slot
    outside [-1, n) (agent.py:377 indexes NumPy arrays with the paths)."""
    from . import engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(idx_all), rank, world)
    dev = state.X.device
    scores = bad = None
    if hi > lo:
        if torch.is_tensor(idx_all):                  # candidates already on the device: this rank's block is a view
            idx = idx_all[lo:hi]
            dl = None if delta_all is None else delta_all[lo:hi]
        else:
            idx = engine.to_dev(np.ascontiguousarray(idx_all[lo:hi]), dtype=torch.int32, device=dev)
            dl = None if delta_all is None else engine.to_dev(np.ascontiguousarray(delta_all[lo:hi]), device=dev)
            if check:
                bad = state.check_indices(idx)
        if events is not None:
            events[0].record()
        scores = state.score_sets(idx, dl, delta_scalar=delta_scalar, skip=skip, H_base=H_base)
        if events is not None:
            events[1].record()
    def raise_if_bad(result):
        if bad is not None and int(bad.item()) != 0:
            raise IndexError("sharded_best: %d slot(s) of rows %d..%d of the candidate array are outside [-1, %d)"
                             % (int(bad.item()), lo, hi - 1, state.n))
        return result

    ex = peer_exchange(group) if (world > 1 and dev.type == "cuda") else None
    if ex is not None:
        out3 = ex.argmax(scores, idx_offset=lo)       # argmax + NVLink mailbox exchange in one kernel
        return out3 if return_device else raise_if_bad(ex.result(out3))
    if scores is not None:
        pair = state.argmax(scores, idx_offset=lo)
    else:
        pair = pack_pair(-np.inf, np.iinfo(np.int64).max, dev)
    if return_device and world == 1:
        return pair                                   # {score bits, index} on the device, no synchronisation
    return raise_if_bad(allgather_argmax(pair, group))


def row_block(total, rank, world):
    """Equal padded blocks for one all-gather: (rows per rank, lo, hi) of `rank`'s block of `total` rows."""
    per = (int(total) + int(world) - 1) // int(world)
    return per, min(total, rank * per), min(total, (rank + 1) * per)


def gather_rows(local, total, group=None):
    """All-gather the ranks' padded row blocks (1-D tensors of equal length) and cut the padding."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:total]
    out = torch.empty(local.shape[0] * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:total]


def sharded_mean_var(hyper, train_x, train_var, y0, ymean, xs, test_var=None, precision="i8", group=None, src=0):
    """Posterior mean / latent variance over the test rows `xs`, sharded across the ranks of one box.

    The training-set factorisation stays on ONE GPU (rank `src`), as for the scoring path; its inverse factor and
    the weights alpha are broadcast once (NCCL over NVLink: 8 N^2 bytes), then every rank evaluates a contiguous
    block of test rows -- they are independent given the factor -- and the blocks are all-gathered.  With
    precision "i8" and N >= engine.I8_REORDER_MIN the points are sorted along a Z curve first, so a rank's block is
    also spatially compact.  All arguments are device tensors on this rank's GPU and identical on every rank;
    returns (mu[M], var[M]) on every rank, in the caller's row order."""
    from . import engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = xs.device
    N, M = train_x.shape[0], xs.shape[0]
    reorder = engine.uses_digits(precision, N) and N >= engine.I8_REORDER_MIN
    tperm = None
    if reorder:
        perm, lo, hi = engine.morton_perm(train_x)
        train_x = train_x.index_select(0, perm).contiguous()
        train_var = None if train_var is None else train_var.index_select(0, perm).contiguous()
        y0 = y0.index_select(0, perm)
        tperm, _, _ = engine.morton_perm(xs, lo, hi)
        xs = xs.index_select(0, tperm).contiguous()
        test_var = None if test_var is None else test_var.index_select(0, tperm).contiguous()
    Npad = max(engine.BLK, engine.pad_to(N))
    if rank == src:
        fkind, fslices = engine.factor_plan(precision, N)
        f = engine.GPFactor(hyper, train_x, diag_add=train_var, diag_scalar=hyper.noise, factor=fkind, factor_slices=fslices)
        alpha, _ = f.solve(y0)
        head = torch.cat([alpha, f.info.to(torch.float64)])
    else:
        f = engine.GPFactor.__new__(engine.GPFactor)
        f.hyper, f.x, f.N, f.Npad, f.L = hyper, train_x, N, Npad, None
        f.Linv = torch.empty((Npad, Npad), dtype=torch.float64, device=dev)
        f.info = torch.zeros(1, dtype=torch.int32, device=dev)
        f.perm = f.box = None
        head = torch.empty(Npad + 1, dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(f.Linv, src=src, group=group)
        dist.broadcast(head, src=src, group=group)
    alpha = head[:Npad].contiguous()
    f.info = head[Npad:].to(torch.int32)
    # equal padded shards so that one all-gather returns everything
    per, lo_r, hi_r = row_block(M, rank, world)
    mu_loc = torch.zeros(per, dtype=torch.float64, device=dev)
    var_loc = torch.zeros(per, dtype=torch.float64, device=dev)
    if hi_r > lo_r:
        Ks, part = f.cross(xs[lo_r:hi_r], alpha)
        mu_loc[:hi_r - lo_r] = engine.rowsum(part, 1.0, ymean, rows=hi_r - lo_r)
        if precision in engine.I8_FAMILY and f.Npad <= engine.I8_MAX_K:
            rn = f.whiten_norm_i8(Ks, nslices=engine.I8_FAST_SLICES if precision == "i8fast" else engine.I8_SLICES)
        elif precision == "tf32" and f.Npad <= engine.TF32_MAX_N:
            rn = f.whiten_norm_tf32(Ks)
        elif precision == "tf32":
            rn = f.whiten_norm_i8(Ks, nslices=engine.I8_FAST_SLICES)
        else:
            _, rn = f.whiten(Ks, want_V=False)
        tv = None if test_var is None else test_var[lo_r:hi_r].contiguous()
        var_loc[:hi_r - lo_r] = engine.rowsum(rn, -1.0, hyper.outputscale, tv, rows=hi_r - lo_r)
    mu, var = gather_rows(mu_loc, M, group), gather_rows(var_loc, M, group)
    if tperm is not None:
        inv = torch.empty_like(tperm)
        inv[tperm] = torch.arange(M, device=dev)
        mu, var = mu.index_select(0, inv), var.index_select(0, inv)
    f.check()
    return mu, var
