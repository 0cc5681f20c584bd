"""Time / check the recursive INT8 factorisation against DMMA potrf + trtri."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine, _lib
from algp_b200._lib import call, ptr, stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bases = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2048]
slices = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8]
rng = np.random.default_rng(1)
side = 256 if N > 4096 else 64
x = engine.to_dev(rng.uniform(0, side, size=(N, 2)))
if len(sys.argv) > 4 and sys.argv[4] == "morton":
    x = x[engine.morton_perm(x)[0]].contiguous()
hy = engine.Hyper(np.log([side / 16.0, side / 16.0]), 0.0, np.log(1e-2), "rbf")
var = engine.to_dev(np.full(N, 0.01))
Npad = engine.pad_to(N)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
A0, _ = engine.kbuild(hy, x, None, Npad, Npad, var, hy.noise, True)
info = torch.zeros(1, dtype=torch.int32, device=A0.device)
work = torch.empty(max(2, _lib.lib.algp_trtri_work_doubles(Npad)), dtype=torch.float64, device=A0.device)
L = torch.empty_like(A0); Linv = torch.empty_like(A0)
def dmma():
    L.copy_(A0)
    call("algp_potrf", ptr(L), Npad, Npad, ptr(Linv), Npad, ptr(info), stream())
    call("algp_trtri", ptr(L), Npad, Npad, ptr(Linv), Npad, ptr(work), 1, stream())
tc, _ = timed(lambda: L.copy_(A0))
t, _ = timed(dmma)
print("DMMA potrf+trtri N=%d: %.2f ms (copy %.2f ms subtracted)" % (N, t - tc, tc))
Lr, Lir = L.clone(), Linv.clone()
ld_ref = float(torch.log(torch.diagonal(Lr)).sum() * 2)
L8 = torch.empty_like(A0); Linv8 = torch.empty_like(A0)
for S in slices:
    for base in bases:
        def i8():
            L8.copy_(A0)
            engine.potrf_inv_i8(L8, Linv8, info, nslices=S, base=base)
        t8, _ = timed(i8)
        ok = int(info.item())
        dL = float((torch.tril(L8) - torch.tril(Lr)).abs().max())
        dLi = float((Linv8 - Lir).abs().max()) / float(Lir.abs().max())
        ld8 = float(torch.log(torch.diagonal(L8)).sum() * 2)
        print("i8 S=%d base=%d: %.2f ms  info=%d  max|dL|=%.2e  max|dLinv|/max|Linv|=%.2e  rel dlogdet=%.2e" % (
            S, base, t8 - tc, ok, dL, dLi, abs(ld8 - ld_ref) / abs(ld_ref)))
