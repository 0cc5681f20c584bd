"""Where the end-to-end time of Agent.best_path(ndarray[65536, 8], []) goes (configs[2], host arrays in, winner out):
wall time per call, the device time of the scoring + argmax kernels inside it, and a cProfile of the host side."""
import cProfile
import pstats
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import algp_b200
from algp_b200 import engine

torch.cuda.set_device(0)
grid, y, base, idx0, delta, hy = bench.workload(seed_sets=2)
n = len(grid)
is_static = np.zeros(n, bool)
is_static[base] = True


class Env(object):
    pass


env = Env()
env.X, env.test_X, env.num_samples = grid, grid[:16], n
ag = algp_b200.Agent.__new__(algp_b200.Agent)
ag.env, ag.static_std, ag.mobile_std, ag.criterion = env, bench.STATIC_STD, bench.MOBILE_STD, 'entropy'
ag.cov_mode = "never"
ag.static_data = [[0.0] if s else [] for s in is_static]
ag.mobile_data = [[] for _ in range(n)]
ag.collected = {'ind': list(base), 'std': [bench.STATIC_STD] * len(base), 'y': [0.0] * len(base)}
ag.gp = algp_b200.GPR(kernel_params={'type': hy["kind"]})
ag.gp.reset(grid[base], y[base], np.full(len(base), bench.STATIC_STD ** 2))
with torch.no_grad():
    ag.gp.model.kernel_covar_module.base_kernel.log_lengthscale.copy_(torch.tensor(np.log(hy["ls"])).view(1, 1, -1))
    ag.gp.model.kernel_covar_module.log_outputscale.fill_(float(np.log(hy["os"])))
    ag.gp.likelihood.log_noise.fill_(float(np.log(hy["noise"])))
ag._post_update()
for _ in range(3):
    w = ag.best_path(idx0, [])
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    w = ag.best_path(idx0, [])
    ts.append((time.perf_counter() - t0) * 1e3)
print("best_path wall ms per call: min %.3f median %.3f  (winner %d)" % (min(ts), float(np.median(ts)), w))
# the device part alone: same state, device-resident candidates
state = ag._hot_state["state"]
idx_d = engine.to_dev(idx0, dtype=torch.int32)
torch.cuda.synchronize()
dm = 1.0 / bench.MOBILE_STD ** 2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ks = []
for _ in range(10):
    e0.record()
    sc = state.score_sets(idx_d, None, delta_scalar=dm, skip=state._skip)
    state.argmax(sc)
    e1.record()
    torch.cuda.synchronize()
    ks.append(e0.elapsed_time(e1))
print("score + argmax device ms (isolated calls): min %.3f median %.3f" % (min(ks), float(np.median(ks))))
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    d = engine.to_dev(idx0, dtype=torch.int32)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("to_dev(idx 2 MB) + sync wall ms: min %.3f median %.3f" % (min(ts), float(np.median(ts))))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    ag.best_path(idx0, [])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
