"""Small run of the newer kernels for compute-sanitizer memcheck (INT8 digit split / GEMM with masks, recursive
factorisation, block append, long-path scoring)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
rng = np.random.default_rng(0)
N, M = 640, 384
x = np.sort(rng.uniform(0, 300, N))[:, None]
xs = np.sort(rng.uniform(0, 300, M))[:, None]
hy = engine.Hyper(np.log([2.0]), 0.0, np.log(0.05), "rbf")
f = engine.GPFactor(hy, engine.to_dev(x), diag_add=None)
f.check()
Ks, _ = f.cross(engine.to_dev(xs))
_, rn = f.whiten(Ks, want_V=False)
for S in (7, 8):
    a = f.whiten_norm_i8(Ks, nslices=S, use_masks=True)
    b = f.whiten_norm_i8(Ks, nslices=S, use_masks=False)
    assert torch.equal(a, b)
    print("i8 S=%d max diff vs dmma %.2e" % (S, float((a.sum(1) - rn.sum(1)).abs().max())))
A, _ = engine.kbuild(hy, engine.to_dev(x), None, 640, 640, None, hy.noise, True)
Linv = torch.empty_like(A); info = torch.zeros(1, dtype=torch.int32, device=A.device)
engine.potrf_inv_i8(A, Linv, info, nslices=8, base=128)
print("potrf_inv_i8 info", int(info.item()), "max|L-Lref|", float((torch.tril(A) - torch.tril(f.L)).abs().max()))
X2 = rng.uniform(0, 20, (500, 2))
hy2 = engine.Hyper(np.log([2.0, 2.0]), 0.0, np.log(0.05), "matern")
base = rng.choice(500, 120, replace=False)
pi0 = np.zeros(500); pi0[base] = 100.0
st = engine.PosteriorState(hy2, engine.to_dev(X2), base, pi0, is_static=pi0 > 0, capacity=64)
free = np.setdiff1d(np.arange(500), base)
st.append_block([int(j) for j in rng.choice(free, 23, replace=False)], 1.0)
idx = np.full((6, 300), -1, dtype=np.int32)
for r in range(6):
    L = int(rng.integers(130, 300)); idx[r, :L] = rng.choice(500, L, replace=True)
sc = st.score_sets(engine.to_dev(idx, dtype=torch.int32), None, delta_scalar=1.0)
print("long-path scores", sc.cpu().numpy()[:3])
torch.cuda.synchronize()
print("done")
