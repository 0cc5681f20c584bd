"""algp_cov_downdate: time of one pass over the lower triangle of the resident posterior covariance (n = 16384) for
k = 1, 16, 20, 32 appended columns, and the result against P - W W^T computed with torch."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200._lib import call, ptr, stream, lib
n, ldw = 16384, 4160
g = torch.Generator(device="cuda").manual_seed(0)
P0 = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
Wt = torch.randn(n, ldw, dtype=torch.float64, device="cuda", generator=g)
P = torch.empty_like(P0)
print("max columns per pass:", lib.algp_cov_downdate_max_cols())
for k in (1, 16, 20, 32):
    if k > lib.algp_cov_downdate_max_cols():
        continue
    ts = []
    for rep in range(5):
        P.copy_(P0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("algp_cov_downdate", ptr(P), n, n, ptr(Wt), ldw, 4096, k, stream())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    W = Wt[:, 4096:4096 + k]
    rows = slice(5000, 5512)
    want = P0[rows] - W[rows] @ W.T
    got = P[rows]
    mask = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))[rows]
    err = ((got - want).abs() * mask).max().item()
    gb = 2 * 8.0 * n * (n + 64) / 2 / 1e9
    print("k=%2d: %.3f ms (min of 5)  %.0f GB/s on the lower triangle read + written  max|err| on 512 rows %.2e" %
          (k, min(ts), gb / (min(ts) / 1e3), err), flush=True)
