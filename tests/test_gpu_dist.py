"""The NVLink mailbox exchange of the per-rank winners (csrc/p2p.cu) on ONE device: two "ranks" are two streams of
this process with two mailboxes in the same memory, so the protocol (payload, fence, epoch flag, parity double
buffering, first-maximum reduction, timeout status) is exercised without a second GPU.  The multi-GPU run of the same
entry point is scripts/dist_check.py under torchrun (profiles/r02_dist_check_*.log)."""
import ctypes as C

import numpy as np
import pytest
import torch

from algp_b200 import _lib
from gpu_helpers import dev

pytestmark = pytest.mark.gpu


class _Box(object):
    def __init__(self, world):
        self.world = world
        self.ptrs = []
        for _ in range(world):
            p = C.c_void_p()
            h = (C.c_ubyte * 64)()
            _lib.call("algp_p2p_create", _lib.lib.algp_p2p_mailbox_bytes(world), C.byref(p), C.cast(h, C.c_void_p))
            self.ptrs.append(p)
        self.peers = torch.tensor([p.value for p in self.ptrs], dtype=torch.int64, device="cuda")
        self.work = [torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device="cuda") for _ in range(world)]
        self.out = [torch.zeros(3, dtype=torch.int64, device="cuda") for _ in range(world)]
        self.streams = [torch.cuda.Stream() for _ in range(world)]

    def launch(self, rank, x, offset, epoch, timeout_ms=3000.0):
        n = 0 if x is None else x.shape[0]
        with torch.cuda.stream(self.streams[rank]):
            _lib.call("algp_argmax_exchange", _lib.ptr(x) if n else None, n, offset, _lib.ptr(self.work[rank]),
                      _lib.ptr(self.peers), rank, self.world, epoch, timeout_ms, _lib.ptr(self.out[rank]), _lib.stream())

    def close(self):
        torch.cuda.synchronize()
        for p in self.ptrs:
            _lib.call("algp_p2p_destroy", p)


def _decode(out3):
    h = out3.cpu()
    return float(h[0:1].view(torch.float64).item()), int(h[1]), int(h[2])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_mailbox_exchange_matches_numpy_argmax(world):
    rng = np.random.default_rng(world)
    box = _Box(world)
    try:
        torch.cuda.synchronize()
        for epoch in range(1, 8):
            sizes = rng.integers(0 if epoch % 3 == 0 else 1, 5000, world)
            if epoch == 5:
                sizes[:] = 300
            if epoch == 6:
                sizes[0] = 30000                          # beyond the one-CTA scan: block partials + exchange
            blocks, offs, off = [], [], 0
            for r in range(world):
                x = rng.normal(size=int(sizes[r]))
                if epoch == 5:
                    x[:] = np.round(x, 1)                 # many ties, also across ranks: the lowest global index wins
                blocks.append(x)
                offs.append(off)
                off += len(x)
            if sum(len(b) for b in blocks) == 0:
                blocks[0] = rng.normal(size=7)
            xs = [dev(b) if len(b) else None for b in blocks]
            torch.cuda.synchronize()
            for r in range(world):
                box.launch(r, xs[r], offs[r], epoch)
            torch.cuda.synchronize()
            allx = np.concatenate(blocks)
            want_i = int(np.argmax(allx))
            for r in range(world):
                v, i, status = _decode(box.out[r])
                assert status == 0
                assert i == want_i and v == allx[want_i], (epoch, r)
    finally:
        box.close()


def test_mailbox_exchange_reports_a_missing_peer():
    box = _Box(2)
    try:
        x = dev(np.arange(10.0))
        box.launch(0, x, 0, 1, timeout_ms=30.0)             # rank 1 never shows up
        torch.cuda.synchronize()
        assert _decode(box.out[0])[2] == 1
    finally:
        box.close()


def test_exchange_rejects_bad_arguments():
    assert _lib.lib.algp_p2p_mailbox_bytes(17) == 0
    assert _lib.lib.algp_argmax_exchange(None, 5, 0, None, None, 0, 1, 1, 10.0, None, None) == 1
    w = torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device="cuda")
    o = torch.zeros(3, dtype=torch.int64, device="cuda")
    assert _lib.lib.algp_argmax_exchange(None, 0, 0, _lib.ptr(w), _lib.ptr(o), 0, 1, 0, 10.0, _lib.ptr(o), None) == 1   # epoch 0
