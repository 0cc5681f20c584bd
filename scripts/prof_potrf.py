"""algp_potrf alone (CUDA events, best of several) for each diagonal-block kernel variant and matrix size, with the
result checked against the rank-2 factor."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from algp_b200 import engine
from algp_b200._lib import call, ptr, stream
sizes = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [640, 1024, 4096, 16384]
ranks = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2, 1]
rng = np.random.default_rng(0)
for N in sizes:
    side = int(np.sqrt(N)) * 1.0
    x = engine.to_dev(rng.uniform(0, side, size=(N, 2)))
    hy = engine.Hyper(np.log([side / 16.0] * 2), 0.0, np.log(1e-2), "rbf")
    Np = engine.pad_to(N)
    A0, _ = engine.kbuild(hy, x, None, Np, Np, None, hy.noise, True)
    A = A0.clone(); Linv = torch.empty_like(A); info = torch.zeros(1, dtype=torch.int32, device=A.device)
    ref = None
    for r in ranks:
        call("algp_set_potf2_rank", r)
        ts = []
        for rep in range(5):
            A.copy_(A0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call("algp_potrf", ptr(A), Np, Np, ptr(Linv), Np, ptr(info), stream())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        L = torch.tril(A)
        msg = ""
        if ref is None:
            ref = (L.clone(), torch.stack([Linv[i:i + 128, i:i + 128] for i in range(0, Np, 128)]).clone())
        else:
            dl = (L - ref[0]).abs().max().item()
            di = (torch.stack([Linv[i:i + 128, i:i + 128] for i in range(0, Np, 128)]) - ref[1]).abs().max().item()
            msg = "  max|dL| %.2e max|dLinv_diag_blocks| %.2e vs rank %d" % (dl, di, ranks[0])
        print("N=%6d rank %2d: potrf %.3f ms (min of 5), info %d%s" % (N, r, min(ts), int(info.item()), msg), flush=True)
call("algp_set_potf2_rank", 2)
