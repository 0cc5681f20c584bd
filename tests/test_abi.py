"""The C-ABI library builds, loads and exports every symbol include/algp_b200.h declares.
No compute calls: runs without a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "algp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(algp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from algp_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_binding_table_matches_header():
    from algp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.algp_version() >= 100
    assert _lib.lib.algp_strerror(3) == b"matrix is not positive definite"
    # workspace-size helpers are pure host functions
    assert _lib.lib.algp_trtri_work_doubles(4096) == 4096 * 4096 // 4
    assert _lib.lib.algp_trtri_work_doubles(384) >= 256 * 256
    assert _lib.lib.algp_kbuild_col_tiles(4096, 0) == 16


def test_invalid_arguments_are_rejected_without_a_gpu():
    from algp_b200 import _lib
    # null pointers / unpadded sizes return ALGP_ERR_INVALID before any CUDA call
    assert _lib.lib.algp_potrf(None, 128, 128, None, 128, None, None) == 1
    assert _lib.lib.algp_trtri(None, 100, 100, None, 100, None, 0, None) == 1
    assert _lib.lib.algp_score_sets(None, 0, 0, None, 2, None, 0.0, 0, 0.0, None, None, None, 0.0, None, 8, 1, 0.0, None, None) == 1


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "algp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_tracing_toggles_without_a_gpu():
    """algp_b200.tracing installs / removes the per-call hook of _lib.call; nothing touches CUDA until a call is made."""
    from algp_b200 import _lib, tracing
    assert _lib._trace is None
    t = tracing.enable(nvtx=False)
    assert _lib._trace is t and t.totals == {} and t.pending == []
    assert tracing.disable() is t and _lib._trace is None
    with tracing.trace() as t2:
        assert _lib._trace is t2
    assert _lib._trace is None
    text = tracing.format_summary({"algp_potrf": {"calls": 3, "ms": 2.5}})
    assert "algp_potrf" in text and "2.500" in text
