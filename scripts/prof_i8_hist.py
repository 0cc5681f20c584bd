"""Histogram of the per-chunk MMA schedules of the masked INT8 digit GEMM (variance step, Z-ordered points):
how many (tile, k-chunk) pairs survive, how many MMAs of which width they issue, and a cycle model
(max(60, N/2) cycles per MMA, scripts/micro/mma_i8_rate.cu) against the measured kernel time."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
S = int(sys.argv[3]) if len(sys.argv) > 3 else 7
rng = np.random.default_rng(1)
side = int(np.sqrt(M))
x = rng.uniform(0, side, size=(N, 2))
yy, xx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
hy = engine.Hyper(np.log([side / 16.0, side / 16.0]), 0.0, np.log(1e-2), "rbf")
xd, xsd = engine.to_dev(x), engine.to_dev(xs)
p1, lo, hi = engine.morton_perm(xd)
p2, _, _ = engine.morton_perm(xsd, lo, hi)
x, xs = xd[p1].cpu().numpy(), xsd[p2].cpu().numpy()
f = engine.GPFactor(hy, engine.to_dev(x), diag_add=engine.to_dev(np.full(N, 0.01)))
f.check()
Ks, _ = f.cross(engine.to_dev(xs))
kp, ks, km = f.split_i8(Ks, S, 128, want_mask=True)
lp, ls, lm = f._linv_digits(S)
kchunks = f.Npad // 32
MT, NT = Ks.shape[0] // 128, f.Npad // 64
mld = (kchunks + 7) // 8 * 8
a = km.view(-1)[: MT * mld].view(MT, mld)[:, :kchunks].to(torch.int32)          # [MT, kchunks]
mldb = mld
b = lm.view(-1)[: NT * mldb].view(NT, mldb)[:, :kchunks].to(torch.int32)        # [NT, kchunks]
def ffs(v):   # lowest set bit index, 99 if none
    out = torch.full_like(v, 99)
    for p in range(S - 1, -1, -1):
        out = torch.where(((v >> p) & 1) > 0, torch.full_like(v, p), out)
    return out
def fls(v):
    out = torch.full_like(v, -1)
    for p in range(S):
        out = torch.where(((v >> p) & 1) > 0, torch.full_like(v, p), out)
    return out
pmin = ffs(a)                      # [MT, kc]
qmin, qmax = ffs(b), fls(b)        # [NT, kc]
kend = (torch.arange(NT, device=b.device) * 64 + 64) // 32
inrange = torch.arange(kchunks, device=b.device)[None, :] < kend[:, None]
qmin = torch.where(inrange, qmin, torch.full_like(qmin, 99))
# histogram over (pmin, qmin, qmax) triples: count pairs via per-chunk outer products of one-hot counts
cntA = torch.stack([(pmin == p).sum(0) for p in range(S)], 0).double()                      # [S, kc]
cntB = torch.zeros((S, S, kchunks), dtype=torch.float64, device=b.device)
for q0 in range(S):
    for q1 in range(S):
        cntB[q0, q1] = ((qmin == q0) & (qmax == q1)).sum(0)
H = torch.einsum("pk,qrk->pqr", cntA, cntB).cpu().numpy()                                   # pairs per (pmin,qmin,qmax)
total_pairs = float(MT) * float(inrange.sum())
chunks = mmas = 0.0
nbytes = 0.0
cols = 0.0
cyc = 0.0
byN = {}
bycount = {}
for p in range(S):
    for q0 in range(S):
        for q1 in range(q0, S):
            c = H[p, q0, q1]
            if c == 0 or p + q0 > S - 1:
                continue
            chunks += c
            nbytes += c * ((S - q0 - p) * 4096 + (min(S - p, q1 + 1) - q0) * 2048)
            nm = 0
            for pa in range(p, S - q0):
                n = min(S - pa, q1 + 1) - q0
                for part in ([n] if n <= 4 else [4, n - 4]):
                    byN[part] = byN.get(part, 0) + c
                    mmas += c
                    nm += 1
                    cols += c * part * 64
                    cyc += c * max(60.0, part * 32.0)
            bycount[nm] = bycount.get(nm, 0) + c
print("tiles %d x %d, k chunks %d; in-range (tile, chunk) pairs %.3e, surviving %.3e (%.1f %%)" % (MT, NT, kchunks, total_pairs, chunks, 100 * chunks / total_pairs))
print("MMAs %.3e (%.2f per surviving chunk), columns %.3e; dense schedule would be %.3e columns" % (mmas, mmas / chunks, cols, total_pairs * S * (S + 1) / 2 * 64))
print("MMAs by planes-per-instruction:", {k: "%.3e" % v for k, v in sorted(byN.items())})
print("chunks by MMA count:", {k: "%.3e" % v for k, v in sorted(bycount.items())})
clk = 1.9e9
print("operand bytes through L2->SM: %.1f GB (dense schedule: %.1f GB); at 10.3 TB/s: %.1f ms" % (nbytes / 1e9, total_pairs * S * 6144 / 1e9, nbytes / 10.3e12 * 1e3))
print("tensor floor (N/2 cycles): %.1f ms; with the 60-cycle minimum: %.1f ms (148 SMs, %.1f GHz)" % (cols / 2 / 148 / clk * 1e3, cyc / 148 / clk * 1e3, clk / 1e9))
for ov in (50, 100, 150, 200, 300):
    print("  + %d cycles per surviving chunk: %.1f ms" % (ov, (cyc + ov * chunks) / 148 / clk * 1e3))
