"""Where a batch of the configs[4] episode goes (200 x 200 field, Wt 40 000 x ~2300 columns half-way through): device
times of the kernels of one batch, each repeated back to back, and the wall time of a batch with its two host reads."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from algp_b200 import engine  # noqa: E402
from algp_b200.episode import run_episode  # noqa: E402

hyper, grid, static, mobile, path_fn = bench.episode_problem(engine, 200, 1024, 256, 16)
Xd = engine.to_dev(grid)
d_s, d_m = 1.0 / bench.STATIC_STD ** 2, 1.0 / bench.MOBILE_STD ** 2
half = run_episode(hyper, Xd, static, mobile, bench.STATIC_STD, bench.MOBILE_STD, 62, 4, path_fn, capacity=62 * 68 + 16 + 4000,
                   distributed=False)
state, mob = half["state"], half["mobile"]
print("state after 62 batches: n = %d, ncols = %d, ldw = %d; episode so far %.3f ms per batch" % (state.n, state.ncols, state.ldw,
                                                                                                half["ms_per_batch"]))


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ut = torch.empty(state.n, dtype=torch.float64, device="cuda")
pair = torch.empty(2, dtype=torch.int64, device="cuda")
print("greedy_utilities      %.4f ms" % timed(lambda: state.greedy_utilities(d_s, out=ut)))
print("argmax over n         %.4f ms" % timed(lambda: state.argmax(ut, 0, out=pair)))
paths = np.ascontiguousarray(path_fn(0, [100, 5000, 20000, 39000]), dtype=np.int32)
pd = engine.to_dev(paths, dtype=torch.int32)
skip = engine.to_dev(mob.astype(np.uint8), dtype=torch.uint8)
print("score 256 paths x 16  %.4f ms" % timed(lambda: state.score_sets(pd, None, delta_scalar=d_m, skip=skip)))
gb = 8.0 * state.n * state.ncols / 1e9
nc0 = state.ncols
jdev = torch.tensor([12345], dtype=torch.int64, device="cuda")


def app():
    state.append(jdev, d_s, mark_static=False)
    state.ncols = nc0                      # rewind: time the same pass again (diagP drifts, irrelevant here)


t = timed(app)
print("append (rank 1)       %.4f ms  = %.2f TB/s over the %.2f GB of Wt" % (t, gb / t, gb))
blk = torch.tensor(np.arange(3000, 3016), dtype=torch.int64, device="cuda")


def appb():
    state.append_block(blk, d_m, mark_static=False)
    state.ncols = nc0


from algp_b200 import _lib  # noqa: E402
for scalar in (1, 0):
    _lib.lib.algp_set_append_block_scalar(scalar)
    state.Wt[:, nc0:nc0 + 32].zero_()          # the rewinds above left columns behind; the DMMA pass reads a zero tail
    t = timed(appb)
    print("append_block (16) %s %.4f ms  = %.2f TB/s" % ("scalar" if scalar else "DMMA  ", t, gb / t))
full = run_episode(hyper, Xd, static, mobile, bench.STATIC_STD, bench.MOBILE_STD, 125, 4, path_fn, distributed=False)
print("episode, 125 batches: %.3f ms per batch, %.4f ms per acquisition" % (full["ms_per_batch"], full["ms_per_acquisition"]))
t0 = time.perf_counter()
for _ in range(200):
    pair.cpu()
print("24-byte host read     %.4f ms" % ((time.perf_counter() - t0) * 1e3 / 200))
