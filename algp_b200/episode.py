"""Active-sampling episode on the device (BASELINE.json configs[4]; the loop of Agent.run_ipp,
reference agent.py:125-229, reduced to its GP arithmetic).

Per batch: `greedy` picks `per_batch` static locations (agent.py:141), the candidate paths through
them are scored by joint entropy (agent.py:168), the winner's mobile readings are committed
(agent.py:179-192).  Every commit is a rank-1 downdate of the posterior stored as one appended
column of Wt, so nothing is re-factorised; a path's readings are committed as one block append
(Wt read once per 16 readings).  Callers pass the candidate paths per batch (`path_fn`);
algp_b200.paths enumerates them on the reference's planning graph.
With torch.distributed initialised, each rank scores a contiguous block of the paths and every
rank applies the same commits (algp_b200.dist).
"""
import numpy as np
import torch

from . import dist as adist
from . import engine


def run_episode(hyper, X, static_flags, mobile_flags, static_std, mobile_std, batches, per_batch, path_fn,
                capacity=None, return_scores=False, distributed=True):
    """Returns dict(picks=[[...] per batch], best_paths=[...], H=[entropy after each batch], ms_per_batch).

    static_flags / mobile_flags: boolean per location (agent.py:298,302).
    path_fn(batch, picks) -> int array [P, k] of mobile sampling locations per candidate path (-1 = empty)."""
    d_s, d_m = 1.0 / static_std ** 2, 1.0 / mobile_std ** 2
    is_static = np.array(static_flags, dtype=bool)
    mobile = np.array(mobile_flags, dtype=bool)
    pi0 = is_static * d_s + mobile * d_m
    base_idx = np.nonzero(pi0 > 0)[0]
    if capacity is None:
        capacity = batches * (per_batch + 64) + 16
    state = engine.PosteriorState(hyper, X, base_idx, pi0, is_static=is_static, capacity=capacity)
    dev = X.device
    skip = engine.to_dev(mobile.astype(np.uint8), dtype=torch.uint8, device=dev)
    jbuf = torch.empty(1, dtype=torch.int64, device=dev)
    out = dict(picks=[], best_paths=[], H=[], scores=[])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for b in range(batches):
        picks = state.greedy(per_batch, d_s)
        out["picks"].append(picks)
        paths = np.ascontiguousarray(path_fn(b, picks), dtype=np.int32)
        if distributed:
            score, best = adist.sharded_best(state, paths, None, delta_scalar=d_m, skip=skip)
        else:
            sc = state.score_sets(engine.to_dev(paths, dtype=torch.int32, device=dev), None, delta_scalar=d_m, skip=skip)
            pair = state.argmax(sc).cpu()
            score, best = float(pair[0:1].view(torch.float64).item()), int(pair[1].item())
        if len(paths) == 1:                                       # agent.py:362-363
            best = 0
        out["best_paths"].append(int(best))
        if return_scores:
            out["scores"].append(score)
        # commit the winner's NEW mobile readings (the flag is boolean: repeats add nothing)
        seen = []
        for j in paths[best]:
            j = int(j)
            if j < 0 or mobile[j] or j in seen:
                continue
            seen.append(j)
            mobile[j] = True
        if seen:
            # one block append: Wt is read once for the whole path instead of once per reading
            state.append_block(seen, d_m, mark_static=False)
            skip = engine.to_dev(mobile.astype(np.uint8), dtype=torch.uint8, device=dev)
        # H(B + path) is exactly the winning score
        state.H_base_dev.fill_(score)
        state._H_base = float(score)
        out["H"].append(float(score))
    ev1.record()
    torch.cuda.synchronize()
    if state.factor is not None:
        state.factor.check()
    out["ms_per_batch"] = ev0.elapsed_time(ev1) / max(1, batches)
    out["ms_per_acquisition"] = ev0.elapsed_time(ev1) / max(1, batches * per_batch)
    out["state"] = state
    out["mobile"] = mobile
    return out
