"""Device time of the diagonal-block kernel (potrf of a single 128 x 128 block) for each rank variant."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
from algp_b200._lib import call, ptr, stream
rng = np.random.default_rng(0)
x = engine.to_dev(rng.uniform(0, 10, size=(128, 2)))
hy = engine.Hyper(np.log([2.0, 2.0]), 0.0, np.log(1e-2), "rbf")
A0, _ = engine.kbuild(hy, x, None, 128, 128, None, hy.noise, True)
A = A0.clone(); Linv = torch.empty_like(A); info = torch.zeros(1, dtype=torch.int32, device=A.device)
ranks = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 2, 4]
for r in ranks:
    call("algp_set_potf2_rank", r)
    ts = []
    for rep in range(6):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call("algp_potrf", ptr(A), 128, 128, ptr(Linv), 128, ptr(info), stream())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    print("rank %d: %.1f us per potrf(128) (memset + kernel)" % (r, min(ts)))
