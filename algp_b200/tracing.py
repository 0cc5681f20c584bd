"""Per-phase tracing of the hot path.

The reference prints ``time.time()`` deltas per phase of ``run_ipp`` (agent.py:138,152,157,164,171,199-219) and has no
profiler hooks.  Here every C-ABI call can be wrapped in an NVTX range named after its entry point (``algp_kbuild``,
``algp_potrf``, ``algp_score_sets_tiled`` ...: visible in ``ncu --nvtx`` / Nsight Systems) and bracketed by CUDA
events on the launching stream, so a caller gets the device time per kernel family without a profiler:

    import algp_b200.tracing as tracing
    with tracing.trace() as t:
        agent.greedy(4); agent.best_path(paths, [])
    print(t.summary())          # {"algp_score_sets_tiled": {"calls": 1, "ms": 1.52}, ...}

Disabled (the default) the cost is one ``is None`` test per call.  ``ALGP_TRACE=1`` in the environment enables it at
import and prints the summary at interpreter exit (what the reference's prints gave, per kernel instead of per phase).
"""
import atexit
import os

from . import _lib


class Tracer(object):
    MAX_PENDING = 4096

    def __init__(self, nvtx=True):
        self.nvtx = nvtx
        self.totals = {}            # name -> [calls, ms]
        self.pending = []           # (name, start event, end event), in launch order

    def around(self, name, fn):
        import torch
        if self.nvtx:
            torch.cuda.nvtx.range_push(name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            fn()
        finally:
            e1.record()
            if self.nvtx:
                torch.cuda.nvtx.range_pop()
            self.pending.append((name, e0, e1))
            if len(self.pending) > self.MAX_PENDING:
                self._drain(block=False)

    def _drain(self, block):
        """Fold finished event pairs into the totals (all of them when block=True: synchronises)."""
        import torch
        if block and self.pending:
            torch.cuda.synchronize()
        keep = []
        for name, e0, e1 in self.pending:
            if block or e1.query():
                t = self.totals.setdefault(name, [0, 0.0])
                t[0] += 1
                t[1] += e0.elapsed_time(e1)
            else:
                keep.append((name, e0, e1))
        self.pending = keep

    def summary(self):
        """{entry point: {"calls": n, "ms": device milliseconds between the call's first and last kernel}}, slowest first."""
        self._drain(block=True)
        return {k: {"calls": v[0], "ms": v[1]} for k, v in sorted(self.totals.items(), key=lambda kv: -kv[1][1])}

    def reset(self):
        self._drain(block=True)
        self.totals = {}


def enable(nvtx=True):
    """Start tracing every C-ABI call; returns the Tracer (also reachable as algp_b200._lib._trace)."""
    _lib._trace = Tracer(nvtx=nvtx)
    return _lib._trace


def disable():
    t, _lib._trace = _lib._trace, None
    return t


class trace(object):
    """Context manager: tracing on inside the block, the previous state restored after it."""

    def __init__(self, nvtx=True):
        self.nvtx = nvtx

    def __enter__(self):
        self.prev = _lib._trace
        return enable(self.nvtx)

    def __exit__(self, *exc):
        _lib._trace = self.prev
        return False


def format_summary(summary):
    rows = ["%-32s %8s %12s" % ("entry point", "calls", "device ms")]
    rows += ["%-32s %8d %12.3f" % (k, v["calls"], v["ms"]) for k, v in summary.items()]
    return "\n".join(rows)


if os.environ.get("ALGP_TRACE", "") not in ("", "0"):
    _t = enable()

    def _report():
        try:
            print(format_summary(_t.summary()))
        except Exception:            # the CUDA context may already be gone at exit
            pass
    atexit.register(_report)
