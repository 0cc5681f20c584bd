"""Per-tile fixed cost of the INT8 digit GEMM: the variance launch with an all-zero occupancy mask for K (every k chunk
skipped), i.e. TMEM allocation, accumulator zero-init, barrier set-up, epilogue and CTA launch for all tiles."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
N, M, S = 16384, 65536, 7
rng = np.random.default_rng(1)
x = engine.to_dev(rng.uniform(0, 256, size=(N, 2)))
hy = engine.Hyper(np.log([16.0, 16.0]), 0.0, np.log(1e-2), "rbf")
f = engine.GPFactor(hy, x, diag_add=engine.to_dev(np.full(N, 0.01)), factor="i8")
Ks = torch.zeros((M, N), dtype=torch.float64, device=x.device)       # all-zero operand: every digit plane empty
kt, ks, km = f.split_i8(Ks, S, 128, want_mask=True)
lt, ls, lm = f._linv_digits(S)
assert int(km.sum()) == 0
rn = torch.empty((M, N // 64), dtype=torch.float64, device=x.device)
def run():
    engine.call("algp_trmm_rt_i8", engine.ptr(kt), engine.ptr(ks), engine.ptr(km), M, engine.ptr(lt), engine.ptr(ls), engine.ptr(lm),
                N, S, engine.ptr(rn), engine.stream())
run(); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
tiles = (M // 128) * (N // 64)
print("all chunks skipped: %.2f ms for %d tiles = %.2f us per tile per SM" % (best, tiles, best * 1e3 / (tiles / 148)))
assert float(rn.abs().max()) == 0.0
