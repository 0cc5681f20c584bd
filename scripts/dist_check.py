"""Multi-GPU check of algp_b200.dist under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 scripts/dist_check.py

The winners of sharded_best through the NVLink mailboxes (csrc/p2p.cu) against the NCCL all-gather path and against
np.argmax of the full score vector computed on every rank, over many steps (epoch parity, ties, empty shards), and the
per-step cost of the two exchanges.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from algp_b200 import dist as adist, engine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import datetime  # noqa: E402
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

grid, y, base, idx, delta, hy = bench.workload()
hyper = engine.Hyper(np.log(hy["ls"]), np.log(hy["os"]), np.log(hy["noise"]), hy["kind"])
pi0 = np.zeros(len(grid))
pi0[base] = 1.0 / bench.STATIC_STD ** 2
state = engine.PosteriorState(hyper, engine.to_dev(grid, device=dev), base, pi0, is_static=pi0 > 0, cov_mode="never")
H = state.H_base
idx_d, delta_d = engine.to_dev(idx, dtype=torch.int32, device=dev), engine.to_dev(delta, device=dev)
full = state.score_sets(idx_d, delta_d, H_base=H).cpu().numpy()

ex = adist.peer_exchange()
report = {"world": world, "peer_exchange": ex is not None}
ok = True
rng = np.random.default_rng(0)                      # same stream on every rank
for trial in range(60):
    B = int(rng.choice([1, 2, world - 1 if world > 1 else 1, 7, 300, 4096, 65536]))
    start = int(rng.integers(0, 65536 - B + 1))
    sl = slice(start, start + B)
    if trial % 7 == 3:
        # ties across ranks: the same set repeated B times -> the first one must win
        ii = idx_d[start:start + 1].expand(B, 8).contiguous()
        dd = delta_d[start:start + 1].expand(B, 8).contiguous()
        want = (float(full[start]), 0)
    else:
        ii, dd = idx_d[sl], delta_d[sl]
        j = int(np.argmax(full[sl]))
        want = (float(full[sl][j]), j)
    got = adist.sharded_best(state, ii, dd, H_base=H)                      # mailboxes (or NCCL when unavailable)
    host = adist.sharded_best(state, idx[sl] if trial % 7 != 3 else np.repeat(idx[start:start + 1], B, 0),
                              delta[sl] if trial % 7 != 3 else np.repeat(delta[start:start + 1], B, 0), H_base=H)
    if got[1] != want[1] or abs(got[0] - want[0]) > 1e-12 or host != got:
        ok = False
        report.setdefault("mismatch", []).append([trial, B, list(got), list(host), list(want)])
report["winners_match_numpy_argmax_on_every_step"] = ok

# NCCL path explicitly (what round 1 did), same winners
lo, hi = adist.shard_range(65536, rank, world)
sc = state.score_sets(idx_d[lo:hi], delta_d[lo:hi], H_base=H)
pair = state.argmax(sc, idx_offset=lo)
nccl = adist.allgather_argmax(pair)
p2p = adist.sharded_best(state, idx_d, delta_d, H_base=H)
report["nccl_vs_mailbox_same_winner"] = bool(nccl[1] == p2p[1] and nccl[0] == p2p[0] and p2p[1] == int(np.argmax(full)))


def timed(fn, steps=200):
    for _ in range(10):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# exchange cost alone: argmax of a small resident score vector + exchange, no scoring kernel
small = sc[:1024].contiguous()
gathered = torch.empty(2 * world, dtype=torch.int64, device=dev)
pr = torch.empty(2, dtype=torch.int64, device=dev)


def nccl_ex():
    state.argmax(small, idx_offset=lo, out=pr)
    dist.all_gather_into_tensor(gathered, pr)


report["us_per_exchange_nccl_allgather"] = 1e3 * timed(nccl_ex)
if ex is not None:
    report["us_per_exchange_nvlink_mailbox"] = 1e3 * timed(lambda: ex.argmax(small, idx_offset=lo))
    ex.result()
report["us_argmax_only"] = 1e3 * timed(lambda: state.argmax(small, idx_offset=lo, out=pr))
if rank == 0:
    print(json.dumps(report))
dist.barrier()
adist.shutdown()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
