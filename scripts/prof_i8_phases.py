"""NOTE: needs the -DI8_PROFILE hooks of profiles/r01_i8_lookahead_and_phase_hooks.patch applied to csrc/i8.cu.
Per-phase cycle sums of gemm_i8_kernel (library built with -DI8_PROFILE, see scripts/gpu_i8_limits.sh)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine, _lib
N, M, S = 16384, 65536, 7
rng = np.random.default_rng(1)
side = int(np.sqrt(M))
x = rng.uniform(0, side, size=(N, 2))
yy, xx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
xs = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.float64)
hy = engine.Hyper(np.log([side / 16.0, side / 16.0]), 0.0, np.log(1e-2), "rbf")
xd, xsd = engine.to_dev(x), engine.to_dev(xs)
p1, lo, hi = engine.morton_perm(xd)
p2, _, _ = engine.morton_perm(xsd, lo, hi)
f = engine.GPFactor(hy, xd[p1].contiguous(), diag_add=engine.to_dev(np.full(N, 0.01)))
f.check()
Ks, _ = f.cross(xsd[p2].contiguous())
kp, ks, km = f.split_i8(Ks, S, 128, want_mask=True)
lp, ls, lm = f._linv_digits(S)
rn = torch.empty((Ks.shape[0], f.Npad // 64), dtype=torch.float64, device=Ks.device)
lib = _lib.lib
buf = (C.c_ulonglong * 16)()
for masks in (True, False):
    def mm():
        engine.call("algp_trmm_rt_i8", engine.ptr(kp), engine.ptr(ks), engine.ptr(km) if masks else None, Ks.shape[0],
                    engine.ptr(lp), engine.ptr(ls), engine.ptr(lm) if masks else None, f.Npad, S, engine.ptr(rn), engine.stream())
    mm(); torch.cuda.synchronize()
    lib.algp_i8_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mm(); e1.record(); torch.cuda.synchronize()
    lib.algp_i8_prof_read(buf, 1)
    v = [int(b) for b in buf]
    n = max(v[6], 1)
    print("masks=%s: %.2f ms, %d CTAs, k-range chunks per CTA %.1f, surviving %.1f" % (masks, e0.elapsed_time(e1), v[6], v[9] / n, v[10] / n))
    names = {0: "prologue (alloc, barriers, TMEM zero)", 1: "producer loop", 2: "issuer loop", 3: "epilogue warps waiting for the accumulators",
             4: "epilogue (tcgen05.ld + Horner)", 5: "whole CTA", 7: "issuer waiting on full barriers", 8: "producer waiting on empty barriers",
             11: "issuer: next-chunk look-ahead (walker)", 12: "issuer: elect + MMA issue", 13: "issuer: commit"}
    for k in (5, 0, 1, 8, 2, 7, 11, 12, 13, 3, 4):
        print("   %-46s %9.0f cycles per CTA" % (names[k], v[k] / n))
