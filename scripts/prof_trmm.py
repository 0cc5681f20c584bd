"""The fp64 variance TRMM alone (V = Ks Linv^T with fused row norms) on synthetic operands."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algp_b200 import _lib  # noqa: E402

M, N = 8192, 4096
g = torch.Generator(device="cuda").manual_seed(0)
Ks = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
Linv = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g).tril()
rn = torch.empty(M, N // 64, dtype=torch.float64, device="cuda")
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("algp_trmm_rt", _lib.ptr(Ks), M, N, _lib.ptr(Linv), N, N, None, 0, _lib.ptr(rn), _lib.stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rep %d: %.3f ms  %.2f TFLOP/s" % (rep, ms, M * N * N / ms / 1e9))
ref = ((Ks[:256] @ Linv.T) ** 2).view(256, N // 64, 64).sum(-1)
print("max rel err", float(((rn[:256] - ref).abs() / ref).max()))
