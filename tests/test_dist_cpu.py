"""Host-side multi-GPU logic on CPU: world_size-2 gloo, contiguous sharding and the first-max
all-gather reduction (np.argmax semantics across ranks)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from algp_b200.dist import allgather_argmax, combine_pairs, pack_pair, shard_range


def test_shard_range_partitions_contiguously():
    for total in (0, 1, 7, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_combine_pairs_first_max():
    assert combine_pairs([1.0, 3.0, 3.0, -np.inf], [5, 90, 40, 2 ** 62]) == (3.0, 40)
    assert combine_pairs([-np.inf, -np.inf], [7, 3]) == (-np.inf, 3)


def _worker(rank, world, port, scores, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(len(scores), rank, world)
        if hi > lo:
            j = int(np.argmax(scores[lo:hi]))                 # what algp_argmax returns for the shard
            pair = pack_pair(float(scores[lo + j]), lo + j)
        else:
            pair = pack_pair(-np.inf, np.iinfo(np.int64).max)
        q.put((rank, allgather_argmax(pair)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["unique", "tie_across_ranks", "short"])
def test_allgather_argmax_world2_gloo(case):
    rng = np.random.default_rng(0)
    if case == "unique":
        scores = rng.normal(size=1001)
    elif case == "tie_across_ranks":
        scores = rng.normal(size=1000)
        scores[[100, 900]] = 10.0                              # same max in both shards: lowest index wins
    else:
        scores = np.array([2.5])                               # rank 1 gets an empty shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scores, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (float(scores.max()), int(np.argmax(scores)))
    for _, got in res:
        assert got == want


def _rows_worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from algp_b200.dist import gather_rows, row_block
        per, lo, hi = row_block(total, rank, world)
        local = torch.full((per,), -1.0, dtype=torch.float64)
        local[:hi - lo] = torch.arange(lo, hi, dtype=torch.float64) * 2.0     # "variance" of row r = 2 r
        q.put((rank, gather_rows(local, total).numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [1, 5, 64, 1001])
def test_sharded_rows_world2_gloo(total):
    """The row sharding of dist.sharded_mean_var: equal padded blocks, one all-gather, padding cut, every rank
    ends with all rows in order (an odd total leaves the last block short; total = 1 leaves rank 1 empty)."""
    from algp_b200.dist import row_block
    blocks = [row_block(total, r, 2) for r in range(2)]
    assert blocks[0][1] == 0 and blocks[-1][2] == total and blocks[0][2] == blocks[1][1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    want = np.arange(total) * 2.0
    for r in range(2):
        np.testing.assert_array_equal(got[r], want)
