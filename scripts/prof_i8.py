"""Time / check the INT8 digit variance path against the DMMA path (dev + ncu target)."""
import sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
slices = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8]
check = (len(sys.argv) <= 4) or sys.argv[4] != "nocheck"
morton = len(sys.argv) > 5 and sys.argv[5] == "morton"
rng = np.random.default_rng(1)
side = int(np.sqrt(M))
x = rng.uniform(0, side, size=(N, 2))
yy, xx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
hy = engine.Hyper(np.log([side / 16.0, side / 16.0]), 0.0, np.log(1e-2), "rbf")
if morton:
    xd, xsd = engine.to_dev(x), engine.to_dev(xs)
    p1, lo, hi = engine.morton_perm(xd)
    p2, _, _ = engine.morton_perm(xsd, lo, hi)
    x, xs = xd[p1].cpu().numpy(), xsd[p2].cpu().numpy()
f = engine.GPFactor(hy, engine.to_dev(x), diag_add=engine.to_dev(np.full(N, 0.01)))
f.check()
Ks, _ = f.cross(engine.to_dev(xs))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
flops = float(f.Npad) ** 2 * Ks.shape[0]
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader",
                        "-lms", "50"], stdout=open("gpurun_out/i8_smi.log", "w"))
if check:
    t64, (_, rn64) = timed(lambda: f.whiten(Ks, want_V=False), reps=2)
    ref = rn64.sum(1)
    print("fp64 DMMA: %.2f ms  %.1f TFLOP/s" % (t64, flops / t64 / 1e9))
for S in slices:
    ts, _ = timed(lambda: f.split_i8(Ks, S, 128))
    f._linv_i8 = None
    t8, rn8 = timed(lambda: f.whiten_norm_i8(Ks, nslices=S))
    kp, ks, km = f.split_i8(Ks, S, 128, want_mask=True)
    lp, ls, lm = f._linv_digits(S)
    def mm(masks=True):
        rn = torch.empty((Ks.shape[0], f.Npad // 64), dtype=torch.float64, device=Ks.device)
        engine.call("algp_trmm_rt_i8", engine.ptr(kp), engine.ptr(ks), engine.ptr(km) if masks else None, Ks.shape[0],
                    engine.ptr(lp), engine.ptr(ls), engine.ptr(lm) if masks else None, f.Npad, S, engine.ptr(rn), engine.stream())
        return rn
    tm, _ = timed(mm)
    tm0, _ = timed(lambda: mm(False))
    occ = [float(((km >> p) & 1).float().mean()) for p in range(S)]
    print("   plane occupancy of K tiles", [round(o, 3) for o in occ], " mma without masks %.2f ms" % tm0)
    ops = flops * S * (S + 1) / 2
    msg = "i8 S=%d: total %.2f ms (split K %.2f ms, mma %.2f ms = %.0f TOP/s int8, %.1f TFLOP/s fp64-equivalent)" % (
        S, t8, ts, tm, ops / tm / 1e9, flops / tm / 1e9)
    if check:
        msg += "  max|d rn| = %.3e" % float((rn8.sum(1) - ref).abs().max())
    print(msg)
smi.terminate()
import collections
rows = [l.strip().split(", ") for l in open("gpurun_out/i8_smi.log") if l.strip()]
busy = [r for r in rows if float(r[1].split()[0]) > 500]
if busy:
    mhz = sorted(int(r[0].split()[0]) for r in busy)
    print("under load (>500 W): %d samples, SM MHz median %d min %d, power max %.0f W, reasons %s" % (
        len(busy), mhz[len(mhz) // 2], mhz[0], max(float(r[1].split()[0]) for r in busy), collections.Counter(r[2] for r in busy)))
