"""ctypes binding of libalgp_b200.so (include/algp_b200.h).

There is NO CPU fallback: if the library is missing the import of any product
module fails loudly, and every entry point raises on a non-zero status.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALGP_B200_LIB") or os.path.join(_HERE, "lib", "libalgp_b200.so")   # env override: tuning builds

_p, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double

# name -> (restype, argtypes); mirrors include/algp_b200.h one to one
SIGNATURES = {
    "algp_version": (C.c_int, []),
    "algp_strerror": (C.c_char_p, [C.c_int]),
    "algp_last_cuda_error": (C.c_char_p, []),
    "algp_kbuild": (C.c_int, [_p, _i64, _p, _i64, _i32, _p, _f64, _i32, _p, _f64, _i32,
                              _p, _i64, _i64, _i64, _i32, _p, _p, _p]),
    "algp_kbuild_col_tiles": (C.c_int, [_i64, _i32]),
    "algp_rowsum": (C.c_int, [_p, _i64, _i32, _f64, _f64, _p, _p, _p]),
    "algp_scatter_add": (C.c_int, [_p, _i64, _p, _i64, _f64, _p]),
    "algp_potrf": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p]),
    "algp_set_potf2_rank": (C.c_int, [_i32]),
    "algp_trtri": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _i32, _p]),
    "algp_trtri_work_doubles": (_i64, [_i64]),
    "algp_gemv_lower": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "algp_gemv_lower_t": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p]),
    "algp_gemv_work_doubles": (_i64, [_i64]),
    "algp_logdet_sumsq": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "algp_trmm_rt": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _p]),
    "algp_gemm_nt": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _f64, _f64, _i32, _p]),
    "algp_split_tf32": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "algp_trmm_rt_tf32": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i64, _i64, _p, _p]),
    "algp_split_i8": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
    "algp_i8_mask_bytes": (_i64, [_i64, _i64, _i32]),
    "algp_trmm_rt_i8": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p, _p]),
    "algp_trmm_rt_store_i8": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p, _i64, _p, _p]),
    "algp_gemm_nt_i8": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _i64, _i32, _f64, _f64, _p, _i64, _i32, _i32, _p]),
    "algp_potrf_inv_i8": (C.c_int, [_p, _i64, _i64, _p, _i64, _i32, _i64, _p, _i64, _p, _p]),
    "algp_potrf_inv_i8_work_bytes": (_i64, [_i64, _i32, _i64]),
    "algp_potri_lower": (C.c_int, [_p, _i64, _i64, _p, _i64, _p]),
    "algp_mll_grad": (C.c_int, [_p, _i64, _i32, _p, _f64, _i32, _f64, _p, _p, _i64, _p, _p, _p]),
    "algp_mll_grad_work_doubles": (_i64, [_i64]),
    "algp_colsumsq_lower": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "algp_colsumsq_work_doubles": (_i64, [_i64]),
    "algp_inv_rank1_update": (C.c_int, [_p, _i64, _p, _i64, _p, _i32, _i64, _i32, _f64, _p, _p, _p]),
    "algp_mi_terms": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _i32, _i64, _p, _f64, _f64, _p, _p]),
    "algp_prefix_reduce": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i32, _p, _p]),
    "algp_score_sets": (C.c_int, [_p, _i64, _i64, _p, _i32, _p, _f64, _i32, _f64, _p, _p, _p, _f64, _p,
                                  _i32, _i64, _f64, _p, _p]),
    "algp_score_sets_large": (C.c_int, [_p, _i64, _i64, _p, _i32, _p, _f64, _i32, _f64, _p, _p, _p, _f64, _p,
                                        _i32, _i64, _f64, _p, _p, _i64, _p]),
    "algp_score_sets_large_work_doubles": (_i64, [_i32, _i64]),
    "algp_score_sets_tiled": (C.c_int, [_p, _i64, _i64, _i64, _p, _i32, _p, _f64, _i32, _f64, _p, _p, _p, _f64, _p,
                                        _i32, _i64, _f64, _p, _p, _i64, _p]),
    "algp_score_sets_tiled_work_doubles": (_i64, [_i64]),
    "algp_score_sets_tiled_launches": (C.c_int, [_i32, _i64, _i64, _i64]),
    "algp_set_score_tile_cols": (C.c_int, [_i32]),
    "algp_set_score_resident": (C.c_int, [_i32]),
    "algp_score_sets_cov": (C.c_int, [_p, _i64, _p, _p, _p, _f64, _p, _i32, _i64, _f64, _p, _p]),
    "algp_mi_terms_large": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _i32, _i64, _p, _f64, _f64, _p, _p, _i64, _p]),
    "algp_mi_terms_large_work_doubles": (_i64, [_i32, _i64]),
    "algp_probe_l2_read": (C.c_int, [_p, _i64, _i32, _i32, _p, _p]),
    "algp_paths_enumerate": (C.c_int, [_i32, _p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _i32, _f64, _f64, _i64, _p]),
    "algp_paths_sizes": (C.c_int, [_p, _p, _p]),
    "algp_paths_fetch": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "algp_paths_fill_slots": (C.c_int, [_p, _p, _i64]),
    "algp_paths_free": (C.c_int, [_p]),
    "algp_cov_downdate": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _i32, _p]),
    "algp_cov_downdate_max_cols": (C.c_int, []),
    "algp_greedy_utilities": (C.c_int, [_p, _p, _p, _f64, _i64, _p, _p]),
    "algp_argmax": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "algp_argmax_work_bytes": (_i64, []),
    "algp_check_indices": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "algp_p2p_mailbox_bytes": (_i64, [_i32]),
    "algp_p2p_create": (C.c_int, [_i64, C.POINTER(C.c_void_p), _p]),
    "algp_p2p_open": (C.c_int, [_p, C.POINTER(C.c_void_p)]),
    "algp_p2p_close": (C.c_int, [_p]),
    "algp_p2p_destroy": (C.c_int, [_p]),
    "algp_argmax_exchange": (C.c_int, [_p, _i64, _i64, _p, _p, _i32, _i32, _i64, _f64, _p, _p]),
    "algp_append": (C.c_int, [_p, _i64, _i64, _p, _i64, _i32, _p, _f64, _i32, _f64, _p, _p, _p, _p, _f64,
                              _i32, _p, _p]),
    "algp_append_work_doubles": (_i64, [_i64]),
    "algp_append_block": (C.c_int, [_p, _i64, _i64, _p, _i64, _i32, _p, _f64, _i32, _f64, _p, _p, _p, _p, _i32, _p, _f64,
                                    _i32, _p, _p]),
    "algp_append_block_work_doubles": (_i64, []),
    "algp_set_append_block_scalar": (C.c_int, [_i32]),
}

ERR_NOT_PD = 3


class AlgpError(RuntimeError):
    def __init__(self, fn, code, msg):
        super().__init__("%s failed: %s (status %d)" % (fn, msg, code))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "algp_b200: %s is missing. Build it with `python algp_b200/build.py` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(name, code):
    if code != 0:
        msg = lib.algp_strerror(code).decode()
        if code == 2:
            msg += ": " + lib.algp_last_cuda_error().decode()
        raise AlgpError(name, code, msg)


# kernels launched per call of the multi-kernel entry points (everything else launches one; host-only entry points
# never go through call()); bench.py reads launch_count around its timed region for the "gpu_launches" it reports
KERNELS_PER_CALL = {"algp_argmax": 2, "algp_argmax_exchange": 2, "algp_append": 3, "algp_append_block": 3}   # argmax: 1 up to 2^14 values
launch_count = 0
_trace = None          # an algp_b200.tracing.Tracer while tracing is enabled (NVTX range + CUDA events around every call)


def call(name, *args, launches=None):
    """Invoke a status-returning entry point and raise on failure.  launches: kernels this call launches when it is not
    the fixed number of KERNELS_PER_CALL (the chunked scoring call)."""
    global launch_count
    launch_count += KERNELS_PER_CALL.get(name, 1) if launches is None else launches
    if _trace is None:
        check(name, getattr(lib, name)(*args))
    else:
        _trace.around(name, lambda: check(name, getattr(lib, name)(*args)))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def host_f64(a):
    """Host double array argument (hyper-parameters travel on the host)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
