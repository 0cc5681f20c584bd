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


def pack_pair(value, index, device="cpu"):
    """Host-side constructor of the {score bits, index} pair (tests, empty shards)."""
    v = torch.tensor([value], dtype=torch.float64).view(torch.int64)
    return torch.cat([v, torch.tensor([index], dtype=torch.int64)]).to(device)


def sharded_best(state, idx_all, delta_all=None, delta_scalar=0.0, skip=None, group=None):
    """Score this rank's block of the global candidate array idx_all [B,k] (host int32) against the
    replicated PosteriorState and return the global (score, index) winner on every rank."""
    from . import engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(idx_all), rank, world)
    dev = state.X.device
    if hi > lo:
        idx = engine.to_dev(np.ascontiguousarray(idx_all[lo:hi]), dtype=torch.int32, device=dev)
        dl = None if delta_all is None else engine.to_dev(np.ascontiguousarray(delta_all[lo:hi]), device=dev)
        scores = state.score_sets(idx, dl, delta_scalar=delta_scalar, skip=skip)
        pair = state.argmax(scores, idx_offset=lo)
    else:
        pair = pack_pair(-np.inf, np.iinfo(np.int64).max, dev)
    return allgather_argmax(pair, group)
