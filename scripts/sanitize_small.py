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

# ---- round 2 kernels: resident covariance across commits, split-candidate scoring, chunked scoring, MI rank-1
# maintenance, single-launch argmax, the NVLink mailbox exchange (two "ranks" as two streams of this device)
import ctypes as C
from algp_b200 import _lib
st2 = engine.PosteriorState(hy2, engine.to_dev(X2), base, pi0, is_static=pi0 > 0, capacity=64, cov_mode="always")
idx8 = rng.choice(free, (700, 8)).astype(np.int32)
idx8_d = engine.to_dev(idx8, dtype=torch.int32)
s_cov = st2.score_sets(idx8_d, None, delta_scalar=1.0).clone()
picks = st2.greedy(2, 100.0)
st2.append_block([int(j) for j in free[:18] if int(j) not in picks], 1.0)
s_cov2 = st2.score_sets(idx8_d, None, delta_scalar=1.0).clone()           # cov_downdate: 2 + 18 columns
st2.cov_mode, keep, st2.P = "never", st2.P, None
for mode, tile in (("stream", 0), ("tiled", 0), ("tiled", 64), ("tiled", -1)):
    st2.score_mode = mode
    _lib.lib.algp_set_score_tile_cols(tile)
    s = st2.score_sets(idx8_d, None, delta_scalar=1.0)
    print("scores %s/%d vs resident covariance after commits: %.2e" % (mode, tile, float((s - s_cov2).abs().max())))
_lib.lib.algp_set_score_tile_cols(0)
pi2 = st2.pi.cpu().numpy()
ctx = engine.MIContext(hy2, engine.to_dev(X2), pi2)
for j in (int(free[40]), int(free[41])):
    ctx.commit(j, 0.1, 1.0)
print("MI rank-1: ld2 %.6f ld3 %.6f" % (float(ctx.ld2), float(ctx.ld3)))
out = torch.empty(2, dtype=torch.int64, device="cuda")
aw = torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device="cuda")
for n in (5, 3000, 40000):
    v = engine.to_dev(rng.normal(size=n))
    _lib.call("algp_argmax", _lib.ptr(v), n, 0, _lib.ptr(out), _lib.ptr(aw), _lib.stream())
boxes, handles = [], []
for _ in range(2):
    p = C.c_void_p(); h = (C.c_ubyte * 64)()
    _lib.call("algp_p2p_create", _lib.lib.algp_p2p_mailbox_bytes(2), C.byref(p), C.cast(h, C.c_void_p))
    boxes.append(p)
peers = torch.tensor([b.value for b in boxes], dtype=torch.int64, device="cuda")
outs = [torch.zeros(3, dtype=torch.int64, device="cuda") for _ in range(2)]
works = [torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
xs2 = [engine.to_dev(rng.normal(size=900)), engine.to_dev(rng.normal(size=20000))]
torch.cuda.synchronize()
for epoch in (1, 2, 3):
    for r in range(2):
        with torch.cuda.stream(streams[r]):
            _lib.call("algp_argmax_exchange", _lib.ptr(xs2[r]), xs2[r].shape[0], r * 900, _lib.ptr(works[r]), _lib.ptr(peers), r, 2,
                      epoch, 3000.0, _lib.ptr(outs[r]), _lib.stream())
    torch.cuda.synchronize()
print("exchange", outs[0].cpu().tolist()[1:], outs[1].cpu().tolist()[1:])
for b in boxes:
    _lib.call("algp_p2p_destroy", b)
print("round-2 kernels done")
