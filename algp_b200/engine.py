"""Device-resident building blocks behind the reference-facing API.

``GPFactor``       A = s^2 k(X,X) + diag(var) + sigma_n^2 I = L L^T with L^-1: everything
                   ``predictive_distribution`` (reference utils.py:293-319) needs.
``PosteriorState`` the factored base set plus Wt = Sigma_{:,B} L^-T and diag(P): everything
                   ``Agent.greedy`` / ``Agent.best_path`` (agent.py:295-403) need.

PyTorch owns the memory and the stream; all arithmetic is in libalgp_b200.so
(hand-written sm_100a kernels, include/algp_b200.h).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

BLK = 128
MAX_SET_SMEM = 128     # slots per candidate that fit the shared-memory scoring kernel; longer paths use the global-scratch variant
MAX_SET = 2048         # slots per candidate algp_score_sets_large accepts
I8_FACTOR_SLICES = 8   # digit planes of the recursive INT8 factorisation (56 bits: fp64-grade trailing updates)
I8_FACTOR_BASE = 2048  # blocks of this many rows or fewer are factored by the DMMA kernels
I8_FACTOR_MIN = 8192   # factor="auto": smallest padded N that takes the INT8 factorisation
I8_REORDER_MIN = 8192  # precision "i8": from this many training points the posterior path sorts train / test points along a Z curve
I8_MAX_K = 32768       # k extent the digit GEMM accepts (8 pairs x 2^15 x 2^12 < 2^31)
# digit planes of the INT8 variance path.  8 planes (56 bits below each row's scale) keep the variance within 1e-14 s^2
# of the DMMA result, i.e. the north star's RELATIVE 1e-9 holds down to posterior variances of 1e-5 s^2; 7 planes
# (49 bits, 6e-13 s^2) are 24 % faster on the variance step and meet the tier relative to the prior scale only
I8_SLICES = 8
# precision "i8fast": the same digit path with fewer planes, for the 1e-4 tier of the north star (what the TF32 mode
# targets): 5 planes (35 bits) in the factorisation, 4 (28 bits) in the variance product -- more accurate than the
# split-TF32 path (fp32 accumulation) and 3-4x faster at N = 16384
I8_FAST_FACTOR_SLICES = 5
I8_FAST_SLICES = 4
I8_FAMILY = ("i8", "i8fast")
# precision "tf32" names the 1e-4 tier on the tensor cores.  Up to this many (padded) training points it is the
# split-TF32 tcgen05 variance kernel on an fp64 factor (fp32 accumulation: 4.7e-5 s^2 at N = 4096); beyond, fp32
# accumulation of N cancelling terms leaves the tier (1.7e-4 s^2 at N = 16384, measured), and the tier is served by the
# digit path of "i8fast" (5-plane factor, 4-plane variance: 5e-6 s^2, and 2.7x faster there)
TF32_MAX_N = 8192


def factor_plan(precision, n_train):
    """(factor kind, digit planes) of a precision mode: "i8" factors through 8-plane digit GEMMs where N is large enough
    to pay ("auto"), the 1e-4-tier modes through 5 planes -- "i8fast" always, "tf32" above TF32_MAX_N (where it is the
    same path as "i8fast") --, everything else through DMMA."""
    if precision == "i8":
        return "auto", None
    if precision == "i8fast" or (precision == "tf32" and pad_to(n_train) > TF32_MAX_N):
        return "auto", I8_FAST_FACTOR_SLICES
    return "dmma", None


def uses_digits(precision, n_train):
    """True when the O(N^2 M) variance step of this precision runs as INT8 digit GEMMs (so Z-ordering pays)."""
    return precision in I8_FAMILY or (precision == "tf32" and pad_to(n_train) > TF32_MAX_N)
CONST = 0.5 * np.log(2 * np.pi * np.exp(1))        # utils.py:10
KIND = {"rbf": 0, None: 0, "matern": 1}


def pad_to(n, m=BLK):
    return (int(n) + m - 1) // m * m


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("algp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Hyper(object):
    """theta as the reference stores it (logs; run.py:36-37, models.py:180)."""

    def __init__(self, log_lengthscale, log_outputscale, log_noise, kind="rbf"):
        self.log_ls = np.ascontiguousarray(np.asarray(log_lengthscale, dtype=np.float64).reshape(-1))
        self.log_os = float(log_outputscale)
        self.log_noise = float(log_noise)
        if kind not in KIND:
            raise NotImplementedError(kind)         # models.py:226-227
        self.kind = KIND[kind]
        self.kind_name = "rbf" if self.kind == 0 else "matern"

    @property
    def d(self):
        return self.log_ls.shape[0]

    @property
    def outputscale(self):
        return float(np.exp(self.log_os))

    @property
    def noise(self):
        return float(np.exp(self.log_noise))

    def key(self):
        return (self.log_ls.tobytes(), self.log_os, self.log_noise, self.kind)


class _Staging(object):
    """Reusable pinned staging buffers for host -> device copies: two per device, each guarded by the event of the
    last copy that read it, so a call neither allocates pinned memory nor waits for an unrelated earlier copy."""
    MIN_BYTES = 1 << 20

    def __init__(self):
        self.slots = {}

    def copy(self, t, device):
        nbytes = t.numel() * t.element_size()
        key = (device.index if device.index is not None else torch.cuda.current_device())
        ring = self.slots.setdefault(key, {"bufs": [None, None], "events": [None, None], "next": 0})
        i = ring["next"]
        ring["next"] = 1 - i
        buf = ring["bufs"][i]
        if buf is None or buf.numel() < nbytes:
            size = max(self.MIN_BYTES, 1 << (max(nbytes, 1) - 1).bit_length())
            buf = ring["bufs"][i] = torch.empty(size, dtype=torch.uint8, pin_memory=True)
            ring["events"][i] = None
        ev = ring["events"][i]
        if ev is not None:
            ev.synchronize()                      # the copy that last read this buffer has finished
        view = buf[:nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        out = view.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        ring["events"][i] = ev
        return out


_staging = _Staging()


def to_dev(a, dtype=torch.float64, device=None):
    """Host array -> device tensor through a reused pinned staging buffer (or pass a device tensor through)."""
    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.numel() == 0:
        return torch.empty(t.shape, dtype=dtype, device=device)
    return _staging.copy(t, device if isinstance(device, torch.device) else torch.device(device))


def kbuild(hyper, x1, x2=None, n1_pad=None, n2_pad=None, diag_add=None, diag_scalar=0.0, pad_identity=False,
           out=None, dtype=torch.float64, dot_vec=None):
    """GPR.cov_mat on the device (models.py:161-181).  Returns (out, dot_partial|None)."""
    n1, d = x1.shape
    n2 = n1 if x2 is None else x2.shape[0]
    if d != hyper.d:
        raise ValueError("x has %d dims, hyper-parameters have %d" % (d, hyper.d))
    n1_pad = n1 if n1_pad is None else n1_pad
    vec = 2 if dtype == torch.float64 else 4
    n2_pad = pad_to(n2, vec) if n2_pad is None else n2_pad
    if out is None:
        out = torch.empty((n1_pad, n2_pad), dtype=dtype, device=x1.device)
    ld = out.stride(0)
    partial = None
    odt = 0 if dtype == torch.float64 else 1
    if dot_vec is not None:
        nt = _lib.lib.algp_kbuild_col_tiles(n2_pad, odt)
        partial = torch.empty((n1_pad, nt), dtype=torch.float64, device=x1.device)
    ls, ls_p = _lib.host_f64(hyper.log_ls)
    call("algp_kbuild", ptr(x1), n1, ptr(x2), n2, d, ls_p, hyper.log_os, hyper.kind,
         ptr(diag_add), float(diag_scalar), int(bool(pad_identity)),
         ptr(out), n1_pad, n2_pad, ld, odt, ptr(dot_vec), ptr(partial), stream())
    return out, partial


def morton_perm(x, lo=None, hi=None):
    """Permutation (device int64) that sorts the rows of x [n, d] along a Z-order curve over the box [lo, hi].
    Spatially ordered points make kernel matrices (and their inverse factors) decay away from the diagonal, which
    the INT8 digit GEMM turns into skipped tiles (occupancy masks); results do not depend on the order."""
    n, d = x.shape
    lo = x.min(0).values if lo is None else lo
    hi = x.max(0).values if hi is None else hi
    bits = max(1, min(16, 62 // d))
    span = torch.clamp(hi - lo, min=1e-300)
    q = torch.clamp(((x - lo) / span * (1 << bits)).to(torch.int64), 0, (1 << bits) - 1)
    b = torch.arange(bits, device=x.device, dtype=torch.int64)
    shift = b[None, :] * d + torch.arange(d, device=x.device, dtype=torch.int64)[:, None]     # [d, bits]
    code = (((q[:, :, None] >> b[None, None, :]) & 1) << shift[None, :, :]).sum(dim=(1, 2))   # bit b of dim j -> bit b d + j
    return torch.argsort(code), lo, hi


def rowsum(partial, scale=1.0, bias=0.0, addvec=None, rows=None):
    rows = partial.shape[0] if rows is None else rows
    out = torch.empty(rows, dtype=torch.float64, device=partial.device)
    call("algp_rowsum", ptr(partial), rows, partial.shape[1], float(scale), float(bias), ptr(addvec), ptr(out), stream())
    return out


def potrf_inv_i8(A, Linv, info, nslices=None, base=None):
    """In place: lower triangle of A <- L, Linv <- L^-1, through the recursive INT8 digit factorisation."""
    nslices = I8_FACTOR_SLICES if nslices is None else nslices
    base = I8_FACTOR_BASE if base is None else base
    npad = A.shape[0]
    nbytes = _lib.lib.algp_potrf_inv_i8_work_bytes(npad, nslices, base)
    work = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    call("algp_potrf_inv_i8", ptr(A), npad, A.stride(0), ptr(Linv), Linv.stride(0), nslices, base, ptr(work), nbytes,
         ptr(info), stream())
    del work


class GPFactor(object):
    """Cholesky factor and explicit inverse factor of the training covariance."""

    def __init__(self, hyper, x, diag_add=None, diag_scalar=None, keep_linv=True, factor="dmma", factor_slices=None,
                 buffers=None):
        """factor="dmma": algp_potrf + algp_trtri (fp64 tensor cores); factor="i8": the recursive factorisation
        whose products run as exact INT8 digit GEMMs (same fp64 tier, faster from N ~ 8192 up); "auto" picks.
        factor_slices: digit planes of the "i8" factorisation (default I8_FACTOR_SLICES = 8, fp64-grade).
        buffers: dict of preallocated device buffers {"L", "Linv" [Npad x Npad], "info" int32[1], "trtri" work} that
        a caller re-factorising the same-sized matrix many times (GPR.fit) keeps across calls."""
        if factor not in ("dmma", "i8", "auto"):
            raise ValueError("factor must be 'dmma', 'i8' or 'auto'")
        self.hyper = hyper
        self.x = x
        self.N = x.shape[0]
        self.Npad = max(BLK, pad_to(self.N))
        dev = x.device
        buffers = buffers if buffers is not None else {}
        if diag_scalar is None:
            diag_scalar = hyper.noise                       # add_likelihood_var=True, models.py:179-180
        self.L, _ = kbuild(hyper, x, None, self.Npad, self.Npad, diag_add, diag_scalar, True, out=buffers.get("L"))
        self.Linv = buffers.get("Linv")
        if self.Linv is None:
            self.Linv = torch.empty((self.Npad, self.Npad), dtype=torch.float64, device=dev)
        self.info = buffers.get("info")
        if self.info is None:
            self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            self.info.zero_()
        if factor == "auto":
            factor = "i8" if I8_FACTOR_MIN <= self.Npad <= 2 * I8_MAX_K else "dmma"
        if factor == "i8" and self.Npad > I8_FACTOR_BASE:
            potrf_inv_i8(self.L, self.Linv, self.info, nslices=factor_slices)
        else:
            call("algp_potrf", ptr(self.L), self.Npad, self.Npad, ptr(self.Linv), self.Npad, ptr(self.info), stream())
            work = buffers.get("trtri")
            if work is None:
                work = torch.empty(max(2, _lib.lib.algp_trtri_work_doubles(self.Npad)), dtype=torch.float64, device=dev)
            call("algp_trtri", ptr(self.L), self.Npad, self.Npad, ptr(self.Linv), self.Npad, ptr(work), 1, stream())
            del work
        self._ld2 = None

    def check(self):
        """Raise if the matrix was not positive definite (one 4-byte D2H; synchronises)."""
        info = int(self.info.item())
        if info != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite (leading minor of order %d)" % info)

    def logdet_quad(self, beta=None):
        """tensor([log det A, |beta|^2]) on the device."""
        out = torch.empty(2, dtype=torch.float64, device=self.L.device)
        call("algp_logdet_sumsq", ptr(self.L), self.Npad, self.Npad, ptr(beta), ptr(out), stream())
        return out

    def solve(self, y0):
        """alpha = A^-1 y0 (and beta = L^-1 y0) for a length-N device vector."""
        dev = self.L.device
        yp = torch.zeros(self.Npad, dtype=torch.float64, device=dev)
        yp[:self.N] = y0
        beta = torch.empty_like(yp)
        alpha = torch.empty_like(yp)
        call("algp_gemv_lower", ptr(self.Linv), self.Npad, self.Npad, ptr(yp), ptr(beta), stream())
        work = torch.empty(_lib.lib.algp_gemv_work_doubles(self.Npad), dtype=torch.float64, device=dev)
        call("algp_gemv_lower_t", ptr(self.Linv), self.Npad, self.Npad, ptr(beta), ptr(alpha), ptr(work), stream())
        return alpha, beta

    def cross(self, xs, alpha=None):
        """Ks = s^2 k(xs, X) padded to [Mpad x Npad]; with alpha also the per-tile mean partials."""
        M = xs.shape[0]
        Mpad = max(BLK, pad_to(M))
        return kbuild(self.hyper, xs, self.x, Mpad, self.Npad, dot_vec=alpha)

    def whiten(self, Ks, want_V=True, want_norm=True, V_out=None):
        """V = Ks L^-T (rows = test points) and/or its squared row norms."""
        Mpad = Ks.shape[0]
        dev = Ks.device
        V = None
        if want_V:
            V = V_out if V_out is not None else torch.empty((Mpad, self.Npad), dtype=torch.float64, device=dev)
        rn = torch.empty((Mpad, self.Npad // 64), dtype=torch.float64, device=dev) if want_norm else None
        call("algp_trmm_rt", ptr(Ks), Mpad, Ks.stride(0), ptr(self.Linv), self.Npad, self.Npad,
             ptr(V), V.stride(0) if V is not None else 0, ptr(rn), stream())
        return V, rn

    def split_tf32(self, M, row_multiple=1):
        """fp64 device matrix -> (hi, lo) float32 planes for the tcgen05 TF32 path (allocation rows
        rounded up to row_multiple, extra rows zero)."""
        rows, cols = M.shape
        arows = pad_to(rows, row_multiple)
        alloc = torch.empty if arows == rows else torch.zeros
        hi = alloc((arows, cols), dtype=torch.float32, device=M.device)
        lo = alloc((arows, cols), dtype=torch.float32, device=M.device)
        call("algp_split_tf32", ptr(M), rows, cols, M.stride(0), ptr(hi), ptr(lo), cols, stream())
        return hi, lo

    def whiten_norm_tf32(self, Ks):
        """Squared row norms of V = Ks L^-T per 128-column tile through split-TF32 tcgen05 MMAs."""
        if getattr(self, "_linv_tf32", None) is None:
            self._linv_tf32 = self.split_tf32(self.Linv, row_multiple=256)
        lh, ll = self._linv_tf32
        kh, kl = self.split_tf32(Ks)
        Mpad = Ks.shape[0]
        rn = torch.empty((Mpad, self.Npad // BLK), dtype=torch.float64, device=Ks.device)
        call("algp_trmm_rt_tf32", ptr(kh), ptr(kl), Mpad, kh.stride(0), ptr(lh), ptr(ll), self.Npad, lh.stride(0),
             ptr(rn), stream())
        return rn

    def split_i8(self, M, nslices, tile_rows, want_mask=False):
        """fp64 device matrix -> (digit tiles int8 [rows*cols*nslices], row_scale [rows]) for the INT8
        tensor-core path (tile_rows = 128: left operand, 64: right operand); with want_mask also the
        plane-occupancy bytes per (row tile, 32-column chunk) that let the GEMM skip all-zero digit tiles."""
        rows, cols = M.shape
        tiles = torch.empty(nslices * rows * cols, dtype=torch.int8, device=M.device)
        scale = torch.empty(rows, dtype=torch.float64, device=M.device)
        mask = None
        if want_mask:
            mask = torch.empty(_lib.lib.algp_i8_mask_bytes(rows, cols, tile_rows), dtype=torch.uint8, device=M.device)
        call("algp_split_i8", ptr(M), rows, cols, M.stride(0), nslices, tile_rows, ptr(tiles), ptr(scale), ptr(mask), stream())
        if want_mask:
            return tiles, scale, mask
        return tiles, scale

    def _linv_digits(self, nslices):
        cache = getattr(self, "_linv_i8", None)
        if cache is None or cache[0] != nslices:
            self._linv_i8 = cache = (nslices,) + self.split_i8(self.Linv, nslices, 64, want_mask=True)
        return cache[1], cache[2], cache[3]

    def whiten_norm_i8(self, Ks, nslices=None, use_masks=True):
        """Squared row norms of V = Ks L^-T per 64-column tile through exact INT8 digit GEMMs (fp64 tier)."""
        nslices = I8_SLICES if nslices is None else nslices
        lt, ls, lm = self._linv_digits(nslices)
        kt, ks, km = self.split_i8(Ks, nslices, 128, want_mask=True)
        Mpad = Ks.shape[0]
        rn = torch.empty((Mpad, self.Npad // 64), dtype=torch.float64, device=Ks.device)
        call("algp_trmm_rt_i8", ptr(kt), ptr(ks), ptr(km) if use_masks else None, Mpad, ptr(lt), ptr(ls),
             ptr(lm) if use_masks else None, self.Npad, nslices, ptr(rn), stream())
        return rn

    def whiten_store_i8(self, Ks, V_out, nslices=I8_FACTOR_SLICES):
        """V_out[:, :Npad] = Ks L^-T through exact INT8 digit GEMMs; returns the row-norm partials [Mpad, Npad/64]."""
        lt, ls, lm = self._linv_digits(nslices)
        kt, ks, km = self.split_i8(Ks, nslices, 128, want_mask=True)
        Mpad = Ks.shape[0]
        rn = torch.empty((Mpad, self.Npad // 64), dtype=torch.float64, device=Ks.device)
        call("algp_trmm_rt_store_i8", ptr(kt), ptr(ks), ptr(km), Mpad, ptr(lt), ptr(ls), ptr(lm), self.Npad, nslices,
             ptr(V_out), V_out.stride(0), ptr(rn), stream())
        return rn

    def mean_var(self, xs, y0, ymean, test_var=None, want_var=True, max_rows=65536, precision="fp64"):
        """Posterior mean (and latent variance) at xs: utils.py:300-308 without the inverse.
        precision="tf32" runs the O(N^2 M) variance step on the tcgen05 tensor cores at the 1e-4 tier (split-TF32 MMAs up
        to TF32_MAX_N training points, the 4-plane digit GEMM beyond);
        precision="i8" runs it as exact INT8 digit GEMMs on the same tensor cores (fp64 tier), "i8fast" as the same
        digit GEMMs with 4 planes (1e-4 tier)."""
        if precision not in ("fp64", "tf32", "i8", "i8fast"):
            raise ValueError("precision must be 'fp64', 'i8', 'i8fast' or 'tf32'")
        if precision in I8_FAMILY and self.Npad > I8_MAX_K:
            precision = "fp64"        # beyond the exact-accumulation range of the digit GEMM: DMMA path
        alpha, _ = self.solve(y0)
        M = xs.shape[0]
        mu = torch.empty(M, dtype=torch.float64, device=xs.device)
        var = torch.empty(M, dtype=torch.float64, device=xs.device) if want_var else None
        for lo in range(0, M, max_rows):
            hi = min(M, lo + max_rows)
            Ks, part = self.cross(xs[lo:hi], alpha)
            mu[lo:hi] = rowsum(part, 1.0, ymean, rows=hi - lo)
            if want_var:
                if precision == "tf32" and self.Npad <= TF32_MAX_N:
                    rn = self.whiten_norm_tf32(Ks)
                elif precision == "tf32":
                    rn = self.whiten_norm_i8(Ks, nslices=I8_FAST_SLICES)      # same tier, see TF32_MAX_N
                elif precision == "i8":
                    rn = self.whiten_norm_i8(Ks)
                elif precision == "i8fast":
                    rn = self.whiten_norm_i8(Ks, nslices=I8_FAST_SLICES)
                else:
                    _, rn = self.whiten(Ks, want_V=False)
                tv = None if test_var is None else test_var[lo:hi].contiguous()
                var[lo:hi] = rowsum(rn, -1.0, self.hyper.outputscale, tv, rows=hi - lo)
            del Ks
        return mu, var


SCRATCH_FREE_FRACTION = 0.25      # long-path scratch (k > 128) takes at most this share of the free device memory


def scratch_for(owner, attr, preferred, minimum, device):
    """(doubles, tensor): a scratch cached on `owner`, `preferred` doubles if a quarter of the free memory allows,
    never less than `minimum` (one candidate's matrix; the kernels then run fewer CTAs)."""
    if preferred <= 0:
        return 0, None
    free, _ = torch.cuda.mem_get_info(device)
    cached = getattr(owner, attr, None)
    have = 0 if cached is None else cached.numel()
    budget = int(SCRATCH_FREE_FRACTION * (free + 8 * have)) // 8
    want = max(int(minimum), min(int(preferred), budget))
    if 8 * want > free + 8 * have:
        raise MemoryError("scoring paths of this length needs %.1f GB of scratch, %.1f GB are free" % (8e-9 * want, 1e-9 * free))
    if cached is None or cached.numel() < want:
        cached = None
        setattr(owner, attr, None)
        cached = torch.empty(want, dtype=torch.float64, device=device)
        setattr(owner, attr, cached)
    return cached.numel(), cached


def gemm_nt(A, B, C, alpha, beta, lower_only=False):
    call("algp_gemm_nt", ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(C), C.stride(0),
         C.shape[0], C.shape[1], A.shape[1], float(alpha), float(beta), int(lower_only), stream())
    return C


def chol_logdet(Apad, n):
    """log det of an SPD matrix held in a padded device buffer (destroys it): replaces slogdet (utils.py:193)."""
    npad = Apad.shape[0]
    dev = Apad.device
    scratch = torch.empty((npad, npad), dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    call("algp_potrf", ptr(Apad), npad, Apad.stride(0), ptr(scratch), npad, ptr(info), stream())
    out = torch.empty(2, dtype=torch.float64, device=dev)
    call("algp_logdet_sumsq", ptr(Apad), npad, Apad.stride(0), None, ptr(out), stream())
    res = out.cpu()
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite (leading minor of order %d)" % int(info.item()))
    return float(res[0])


class PosteriorState(object):
    """Factored base set B (locations with precision pi0 > 0) of a field of n locations.

    Holds Wt [n x ldw] (row loc = L^-1 Sigma_{B,loc}), diagP[n] = diag(Sigma - W^T W),
    H_base = H(B), the running precisions pi and static flags.  ``capacity`` extra
    columns are reserved for rank-1 appends (greedy commits)."""

    def __init__(self, hyper, X, base_idx, pi0, is_static=None, capacity=0, precision="fp64", cov_mode="auto"):
        """precision="i8": the factor (from N = 8192) and the W^T build (from N = 1024) run as exact INT8 digit
        GEMMs with 8 planes (fp64-grade) instead of DMMA.
        cov_mode: "never" scores candidate sets by streaming rows of Wt; "always" builds the resident posterior
        covariance P (build_cov) before the first scoring call; "auto" (default) builds it once the scoring work
        streamed against this state would have paid for the build (see score_sets).  A built P survives commits."""
        if cov_mode not in ("auto", "never", "always"):
            raise ValueError("cov_mode must be 'auto', 'never' or 'always'")
        dev = X.device
        self.precision = precision
        self.cov_mode = cov_mode
        self.P = None                  # [n_pad x n_pad] lower triangle of Sigma + noise I - Wt Wt^T, or None
        self._P_ncols = 0              # columns of Wt that P accounts for (appends are folded in lazily: _sync_cov)
        self._stream_s = 0.0           # estimated seconds of Wt streaming since the state last changed
        self.hyper = hyper
        self.X = X
        self.n = X.shape[0]
        self.n_pad = max(BLK, pad_to(self.n))
        base_idx = np.asarray(base_idx, dtype=np.int64)
        pi0 = np.asarray(pi0, dtype=np.float64)
        self.N0 = len(base_idx)
        self.pi = to_dev(pi0, device=dev)
        st = np.zeros(self.n, dtype=np.uint8) if is_static is None else np.asarray(is_static, dtype=np.uint8)
        self.is_static = to_dev(st, dtype=torch.uint8, device=dev)
        prior = hyper.outputscale + hyper.noise           # diag of cov_matrix (agent.py:90)
        if self.N0 == 0:
            self.Npad = 0
            self.ldw = pad_to(max(32, capacity), 32)
            self.Wt = torch.zeros((self.n_pad, self.ldw), dtype=torch.float64, device=dev)
            self.diagP = torch.full((self.n,), prior, dtype=torch.float64, device=dev)
            self.H_base_dev = torch.zeros(1, dtype=torch.float64, device=dev)
            self.factor = None
        else:
            bidx = to_dev(base_idx, dtype=torch.int64, device=dev)
            xb = X.index_select(0, bidx).contiguous()
            inv_pi = to_dev(1.0 / pi0[base_idx], device=dev)
            self.factor = GPFactor(hyper, xb, diag_add=inv_pi, diag_scalar=hyper.noise,
                                   factor="auto" if precision == "i8" else "dmma")
            self.Npad = self.factor.Npad
            self.ldw = pad_to(self.Npad + capacity, 32)      # multiple of the digit GEMM's k chunk (build_cov)
            Ks, _ = kbuild(hyper, X, xb, self.n_pad, self.Npad)
            call("algp_scatter_add", ptr(Ks), Ks.stride(0), ptr(bidx.to(torch.int32)), self.N0, hyper.noise, stream())
            self.Wt = torch.zeros((self.n_pad, self.ldw), dtype=torch.float64, device=dev)
            if precision == "i8" and 1024 <= self.Npad <= I8_MAX_K:
                rn = self.factor.whiten_store_i8(Ks, self.Wt)
            else:
                _, rn = self.factor.whiten(Ks, want_V=True, want_norm=True, V_out=self.Wt)
            del Ks
            self.diagP = rowsum(rn, -1.0, prior, rows=self.n)
            ldq = self.factor.logdet_quad()
            self.H_base_dev = (0.5 * ldq[0:1] + self.N0 * CONST)
        self.ncols = self.Npad                      # valid columns of Wt
        self._H_base = None
        self._argwork = torch.empty(_lib.lib.algp_argmax_work_bytes(), dtype=torch.uint8, device=dev)
        self._appwork = torch.empty(_lib.lib.algp_append_work_doubles(self.n), dtype=torch.float64, device=dev)

    @property
    def H_base(self):
        if self._H_base is None:
            if self.factor is not None:
                self.factor.check()
            self._H_base = float(self.H_base_dev.item())
        return self._H_base

    # ---- resident posterior covariance ("If the build precomputes P (allowed)") ----
    COV_MAX_K = 128                # slots per candidate algp_score_sets_cov accepts
    # priors of the rent-or-buy rule, replaced by measurements as soon as there are any: every streaming call is timed with
    # CUDA events (_harvest_stream_time) and every build calibrates the build rate (_measured_build_rate)
    STREAM_BYTES_PER_S = 10.0e12   # L2->SM delivery of the single-launch scoring kernel (profiles/r02_prof_score_summary.csv); the
                                   # persistent sweep of large batches reaches 13.9e12 (r02_prof_score_resident_summary.csv)
    COV_FLOPS_PER_S = {"fp64": 30.0e12, "i8": 100.0e12}   # lower-triangle SYRK rates of algp_gemm_nt / algp_gemm_nt_i8

    _measured_build_rate = {}      # kind -> flops/s of the last build_cov() in this process (replaces the prior below)

    def cov_build_seconds(self):
        """Expected build_cov() time: n_pad^2 x ncols flops (lower triangle of a SYRK) at the rate the last build in this
        process measured (CUDA events), or at the prior COV_FLOPS_PER_S (round-1 measurements) before the first one."""
        kind = "i8" if (self.precision == "i8" and 1024 <= pad_to(self.ncols, 32) <= I8_MAX_K) else "fp64"
        rate = PosteriorState._measured_build_rate.get(kind, self.COV_FLOPS_PER_S[kind])
        return float(self.n_pad) ** 2 * max(self.ncols, 1) / rate + 1.0e-3

    def _harvest_stream_time(self):
        """Add the MEASURED duration of the streaming score calls that have completed since the last look (CUDA events
        recorded around each call, read without synchronising) to the rent-or-buy account; calls still in flight stay
        on the books at their modelled cost."""
        pending = []
        for e0, e1, modelled in getattr(self, "_stream_events", []):
            if e1.query():
                self._stream_s += e0.elapsed_time(e1) * 1e-3 - modelled
            else:
                pending.append((e0, e1, modelled))
        self._stream_events = pending

    def build_cov(self):
        """P = Sigma + sigma_n^2 I - Wt Wt^T, lower triangle, [n_pad x n_pad] fp64 resident in HBM: the posterior
        covariance of the current base set over all field locations.  Candidate scoring then gathers k(k+1)/2
        entries per set (algp_score_sets_cov) instead of streaming k rows of Wt.  One kernel-matrix build plus one
        SYRK (DMMA, or exact INT8 digit GEMM with precision "i8").  append / append_block keep it: the new columns are
        folded in by _sync_cov (P -= w w^T) before the next scoring call that reads it."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        P, _ = kbuild(self.hyper, self.X, None, self.n_pad, self.n_pad, diag_scalar=self.hyper.noise)
        kpad = pad_to(self.ncols, 32)
        if self.ncols > 0:
            W = self.Wt[:, :kpad]
            if self.precision == "i8" and 1024 <= kpad <= I8_MAX_K:
                at, asc = GPFactor.split_i8(None, W, I8_FACTOR_SLICES, 128)
                bt, bsc = GPFactor.split_i8(None, W, I8_FACTOR_SLICES, 64)
                call("algp_gemm_nt_i8", ptr(at), ptr(asc), self.n_pad, ptr(bt), ptr(bsc), self.n_pad, kpad,
                     I8_FACTOR_SLICES, -1.0, 1.0, ptr(P), P.stride(0), 0, 1, stream())
                del at, bt
            else:
                call("algp_gemm_nt", ptr(self.Wt), self.ldw, ptr(self.Wt), self.ldw, ptr(P), P.stride(0), self.n_pad,
                     self.n_pad, kpad, -1.0, 1.0, 1, stream())
        e1.record()
        self.P = P
        self._P_ncols = self.ncols
        if self.ncols > 0 and self.n_pad >= 4096:
            # calibrate the rent-or-buy rule with this build (a build is rare and tens of ms: the wait is noise)
            e1.synchronize()
            kind = "i8" if (self.precision == "i8" and 1024 <= kpad <= I8_MAX_K) else "fp64"
            secs = max(e0.elapsed_time(e1) * 1e-3 - 1.0e-3, 1.0e-4)
            PosteriorState._measured_build_rate[kind] = float(self.n_pad) ** 2 * self.ncols / secs
        return P

    def drop_cov(self):
        self.P = None
        self._P_ncols = 0
        self._stream_s = 0.0
        self._stream_events = []

    def _sync_cov(self):
        """Bring the resident P up to date with the columns appended since it was built / last synchronised: every
        commit is a rank-1 downdate of P stored as a column of Wt, so P -= w w^T over those columns (one pass over the
        lower triangle per 32 columns) replaces the rebuild.  Lazy: runs before the next scoring call that reads P."""
        while self.P is not None and self._P_ncols < self.ncols:
            k = min(_lib.lib.algp_cov_downdate_max_cols(), self.ncols - self._P_ncols)
            call("algp_cov_downdate", ptr(self.P), self.P.stride(0), self.n_pad, ptr(self.Wt), self.ldw, self._P_ncols, k,
                 stream())
            self._P_ncols += k

    def _want_cov(self, B, k):
        """Rent-or-buy: stream rows of Wt until the streaming done against this unchanged state would have paid
        for the covariance build, then build it (at most twice the cost of the better choice in hindsight)."""
        if self.P is not None:
            if k > self.COV_MAX_K:
                return False
            self._sync_cov()
            return True
        if self.cov_mode == "never" or k > self.COV_MAX_K or self.ncols == 0:
            return False
        if self.cov_mode == "auto":
            self._harvest_stream_time()
            if self._stream_s < self.cov_build_seconds():
                return False
            free, total = torch.cuda.mem_get_info(self.X.device)
            if 8.0 * self.n_pad * self.n_pad * 2.5 > free:          # P plus the transient digit planes of Wt
                return False
        self.build_cov()
        return True

    # k <= 8: "auto" = algp_score_sets_tiled decides (split candidates for small batches; one persistent launch that
    # sweeps L2-sized column chunks with the partial Grams in shared memory for calls that stream >= ~2 GB), "stream" =
    # the plain single launch of algp_score_sets, "tiled" = same entry as "auto" (tests force the chunk and the form
    # through algp_set_score_tile_cols / algp_set_score_resident)
    score_mode = "auto"

    def _want_tiled(self, B, k):
        return self.score_mode != "stream"

    def score_sets(self, idx, delta=None, delta_scalar=0.0, H_base=None, out=None, skip=None):
        """scores[c] = H(S1_c) for candidate sets idx [B,k] (int32 device tensor, -1 = empty)."""
        B, k = idx.shape
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=idx.device)
        if B == 0:
            return out
        ls, ls_p = _lib.host_f64(self.hyper.log_ls)
        hb = self.H_base if H_base is None else H_base
        if self._want_cov(B, k):
            call("algp_score_sets_cov", ptr(self.P), self.P.stride(0), ptr(self.pi), ptr(idx), ptr(delta),
                 float(delta_scalar), ptr(skip), k, B, float(hb), ptr(out), stream())
            return out
        modelled = 8.0 * B * k * max(self.ncols, 1) / self.STREAM_BYTES_PER_S
        self._stream_s += modelled
        ev = None
        if self.cov_mode == "auto" and k <= self.COV_MAX_K and self.ncols > 0:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), modelled)
            ev[0].record()
        try:
            return self._score_streaming(idx, delta, delta_scalar, skip, k, B, hb, out, ls_p)
        finally:
            if ev is not None:
                ev[1].record()
                if not hasattr(self, "_stream_events"):
                    self._stream_events = []
                self._stream_events.append(ev)

    def _score_streaming(self, idx, delta, delta_scalar, skip, k, B, hb, out, ls_p):
        if k <= 8 and self._want_tiled(B, k):
            # small sets with a workspace: split tail / column chunks decided by the library (csrc/score.cu)
            nwork = _lib.lib.algp_score_sets_tiled_work_doubles(B)
            if getattr(self, "_tilework", None) is None or self._tilework.numel() < nwork:
                # zero-filled: the head of the buffer holds the arrival counters of the split-candidate kernel
                self._tilework = torch.zeros(nwork, dtype=torch.float64, device=idx.device)
            # kernel launches this call makes: one (the persistent sweep takes ~68 000 candidates per launch)
            nlaunch = max(1, _lib.lib.algp_score_sets_tiled_launches(k, B, self.ncols, self.n_pad))
            call("algp_score_sets_tiled", ptr(self.Wt), self.ldw, self.ncols, self.n_pad, ptr(self.X), self.hyper.d, ls_p,
                 self.hyper.log_os, self.hyper.kind, self.hyper.noise, ptr(self.pi), ptr(idx), ptr(delta),
                 float(delta_scalar), ptr(skip), k, B, float(hb), ptr(out), ptr(self._tilework), nwork, stream(),
                 launches=nlaunch)
            return out
        if k > MAX_SET_SMEM:
            # long paths: the k x k matrix of a candidate lives in a global scratch instead of shared memory
            kk = pad_to(k, 8)
            nwork, work = scratch_for(self, "_largework", _lib.lib.algp_score_sets_large_work_doubles(k, B), kk * (kk + 1), idx.device)
            call("algp_score_sets_large", ptr(self.Wt), self.ldw, self.ncols, ptr(self.X), self.hyper.d, ls_p,
                 self.hyper.log_os, self.hyper.kind, self.hyper.noise, ptr(self.pi), ptr(idx), ptr(delta),
                 float(delta_scalar), ptr(skip), k, B, float(hb), ptr(out), ptr(work), nwork, stream())
            return out
        call("algp_score_sets", ptr(self.Wt), self.ldw, self.ncols, ptr(self.X), self.hyper.d, ls_p,
             self.hyper.log_os, self.hyper.kind, self.hyper.noise, ptr(self.pi), ptr(idx), ptr(delta),
             float(delta_scalar), ptr(skip), k, B, float(hb), ptr(out), stream())
        return out

    def argmax(self, x, idx_offset=0, out=None):
        """Device {value, index} pair with np.argmax first-max semantics."""
        if out is None:
            out = torch.empty(2, dtype=torch.int64, device=x.device)     # 16 bytes: double + int64
        call("algp_argmax", ptr(x), x.shape[0], int(idx_offset), ptr(out), ptr(self._argwork), stream())
        return out

    def check_indices(self, idx, out=None):
        """Range-check OUR device copy idx [B,k] (int32) of a caller's slot array before it is scored: returns the
        device int64[1] count of slots outside [-1, n) and blanks them (-1), so the scoring kernels never read out of
        bounds.  The reference indexes NumPy arrays with the path lists (agent.py:377) and raises IndexError there; the
        caller of this method raises it once the count has come back with the winner."""
        if out is None:
            out = torch.empty(1, dtype=torch.int64, device=idx.device)
        call("algp_check_indices", ptr(idx), idx.numel(), self.n, ptr(out), stream())
        return out

    def greedy_utilities(self, d_static, out=None):
        if out is None:
            out = torch.empty(self.n, dtype=torch.float64, device=self.X.device)
        call("algp_greedy_utilities", ptr(self.diagP), ptr(self.pi), ptr(self.is_static), float(d_static), self.n,
             ptr(out), stream())
        return out

    def append(self, j_dev, delta, mark_static=True):
        """Commit location *j_dev (device int64) with precision increment delta."""
        if self.ncols >= self.ldw:
            raise RuntimeError("PosteriorState capacity exhausted (%d columns)" % self.ldw)
        ls, ls_p = _lib.host_f64(self.hyper.log_ls)
        call("algp_append", ptr(self.Wt), self.ldw, self.ncols, ptr(self.X), self.n, self.hyper.d, ls_p,
             self.hyper.log_os, self.hyper.kind, self.hyper.noise, ptr(self.diagP), ptr(self.pi), ptr(self.is_static),
             C.c_void_p(j_dev.data_ptr()), float(delta), int(mark_static), ptr(self._appwork), stream())
        self.ncols += 1

    def append_block(self, idx, delta, mark_static=False):
        """Commit the DISTINCT locations idx (host ints or a device int64 tensor) with precision increment delta each:
        the same columns as successive append() calls, with Wt read once per 16 locations."""
        if not torch.is_tensor(idx):
            idx = to_dev(np.asarray(idx, dtype=np.int64), dtype=torch.int64, device=self.X.device)
        k_all = int(idx.shape[0])
        if self.ncols + k_all > self.ldw:
            raise RuntimeError("PosteriorState capacity exhausted (%d columns)" % self.ldw)
        if getattr(self, "_blkwork", None) is None:
            self._blkwork = torch.empty(_lib.lib.algp_append_block_work_doubles(), dtype=torch.float64, device=self.X.device)
        ls, ls_p = _lib.host_f64(self.hyper.log_ls)
        dl_dev, dl_scalar = None, 0.0
        if np.ndim(delta) == 0 and not torch.is_tensor(delta):
            dl_scalar = float(delta)
        else:                                            # one precision increment per location
            dl_dev = delta if torch.is_tensor(delta) else to_dev(np.asarray(delta, dtype=np.float64), device=self.X.device)
        for lo in range(0, k_all, 16):
            k = min(16, k_all - lo)
            call("algp_append_block", ptr(self.Wt), self.ldw, self.ncols, ptr(self.X), self.n, self.hyper.d, ls_p,
                 self.hyper.log_os, self.hyper.kind, self.hyper.noise, ptr(self.diagP), ptr(self.pi), ptr(self.is_static),
                 C.c_void_p(idx.data_ptr() + 8 * lo), k,
                 None if dl_dev is None else C.c_void_p(dl_dev.data_ptr() + 8 * lo), dl_scalar, int(mark_static),
                 ptr(self._blkwork), stream())
            self.ncols += k

    def greedy(self, num_samples, d_static, return_utilities=False):
        """Agent.greedy's selection loop (agent.py:313-354) entirely on the device:
        utilities -> argmax -> rank-1 append, one host read at the end."""
        dev = self.X.device
        pairs = torch.empty((num_samples, 2), dtype=torch.int64, device=dev)
        uts = []
        for s in range(num_samples):
            ut = self.greedy_utilities(d_static)
            self.argmax(ut, 0, out=pairs[s])
            self.append(pairs[s, 1:2], d_static, mark_static=True)
            # H(B + j) = H(B) + ut_j  (agent.py:315: cond = ent_v + sum(cumm_utilities))
            self.H_base_dev += pairs[s, 0:1].view(torch.float64)
            self._H_base = None
            if return_utilities:
                uts.append(ut)
        picks = [int(v) for v in pairs[:, 1].cpu().tolist()]
        if self.factor is not None and not getattr(self, "_factor_checked", False):
            self.factor.check()                      # the factorisation's status word: one 4-byte read, once per state
            self._factor_checked = True
        if return_utilities:
            return picks, torch.stack(uts).cpu().numpy()
        return picks


class MIContext(object):
    """The two extra factorizations the mutual-information criterion needs for one sampled set
    (reference agent.py:330-339): A2 = Sigma_AbarAbar (unsampled locations) and A3 = Sigma + D
    (D = per-location noise variance on the sampled locations, 0 elsewhere)."""

    def __init__(self, hyper, X, pi, full_inverse=False, precision="fp64"):
        """precision="i8": the two n-scale factorizations run as INT8 digit GEMMs where n is large enough to pay
        (GPFactor factor="auto"), as for the posterior state itself."""
        dev = X.device
        fkind = "auto" if precision == "i8" else "dmma"
        self.n = X.shape[0]
        pi = np.asarray(pi, dtype=np.float64)
        sampled = pi > 0
        abar = np.nonzero(~sampled)[0]
        self.n_abar = len(abar)
        pos2 = -np.ones(self.n, dtype=np.int32)
        pos2[abar] = np.arange(self.n_abar, dtype=np.int32)
        self.pos2 = to_dev(pos2, dtype=torch.int32, device=dev)
        self.inv2 = self.inv3 = None
        if self.n_abar:
            xa = X.index_select(0, to_dev(abar, dtype=torch.int64, device=dev)).contiguous()
            self.f2 = GPFactor(hyper, xa, diag_add=None, diag_scalar=hyper.noise, factor=fkind)        # cov_matrix[~S][:, ~S]
            self.ld2 = self.f2.logdet_quad()[0:1].clone()
            self.diag2 = self._inv_diag(self.f2)[:self.n_abar]
            if full_inverse:
                self.inv2 = self._inverse(self.f2)
        else:
            self.f2 = None
            self.ld2 = torch.zeros(1, dtype=torch.float64, device=dev)
            self.diag2 = torch.ones(1, dtype=torch.float64, device=dev)
        var_all = np.where(sampled, 1.0 / np.where(sampled, pi, 1.0), 0.0)               # agent.py:334-337
        self.f3 = GPFactor(hyper, X, diag_add=to_dev(var_all, device=dev), diag_scalar=hyper.noise, factor=fkind)
        self.ld3 = self.f3.logdet_quad()[0:1].clone()
        self.diag3 = self._inv_diag(self.f3)[:self.n]
        if full_inverse:
            self.inv3 = self._inverse(self.f3)

    @staticmethod
    def _inv_diag(f):
        out = torch.empty(f.Npad, dtype=torch.float64, device=f.L.device)
        work = torch.empty(_lib.lib.algp_colsumsq_work_doubles(f.Npad), dtype=torch.float64, device=f.L.device)
        call("algp_colsumsq_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(out), ptr(work), stream())
        return out

    @staticmethod
    def _inverse(f):
        inv = torch.empty((f.Npad, f.Npad), dtype=torch.float64, device=f.L.device)
        call("algp_potri_lower", ptr(f.Linv), f.Npad, f.Npad, ptr(inv), f.Npad, stream())
        return inv

    def check(self):
        if self.f2 is not None:
            self.f2.check()
        self.f3.check()

    MAX_COMMITS = 48          # rank-1 corrections kept per context (algp_inv_rank1_update accepts t < 64)

    def _column(self, f, j):
        """Column j of the ORIGINAL inverse of a factored matrix: A0^-1 e_j (two triangular gemv passes)."""
        e = torch.zeros(f.N, dtype=torch.float64, device=f.L.device)
        e[j] = 1.0
        return f.solve(e)[0]

    def commit(self, j, static_std, mobile_std):
        """A static reading is taken at location j (a greedy pick, agent.py:349-354): maintain diag(A2^-1), logdet A2
        (row / column j leaves Sigma_AbarAbar if j was unsampled) and diag(A3^-1), logdet A3 (the noise variance of j
        changes) by rank-1 updates instead of re-factorising.  Returns False when the context is full (rebuild it)."""
        if getattr(self, "_t3", 0) >= self.MAX_COMMITS or getattr(self, "_t2", 0) >= self.MAX_COMMITS:
            return False
        dev = self.f3.L.device
        ss2, ms2 = static_std ** 2, mobile_std ** 2
        vboth = 1.0 / (1.0 / ss2 + 1.0 / ms2)
        if getattr(self, "_U3", None) is None:
            self._t3 = self._t2 = 0
            self._U3 = torch.empty((self.MAX_COMMITS, self.f3.Npad), dtype=torch.float64, device=dev)
            self._c3 = torch.zeros(64, dtype=torch.float64, device=dev)
            if self.f2 is not None:
                self._U2 = torch.empty((self.MAX_COMMITS, self.f2.Npad), dtype=torch.float64, device=dev)
                self._c2 = torch.zeros(64, dtype=torch.float64, device=dev)
            self._pos2_host = self.pos2.cpu().numpy()
        p = int(self._pos2_host[j])
        was_new = p >= 0
        col3 = self._column(self.f3, j)
        call("algp_inv_rank1_update", ptr(col3), self.n, ptr(self._U3), self._U3.stride(0), ptr(self._c3), self._t3, int(j), 1,
             float(ss2 if was_new else vboth - ms2), ptr(self.diag3), ptr(self.ld3), stream())
        self._t3 += 1
        if was_new:
            col2 = self._column(self.f2, p)
            call("algp_inv_rank1_update", ptr(col2), self.f2.N, ptr(self._U2), self._U2.stride(0), ptr(self._c2), self._t2, p, 0,
                 0.0, ptr(self.diag2), ptr(self.ld2), stream())
            self._t2 += 1
            self._pos2_host[j] = -1
            self.pos2[j] = -1
            self.n_abar -= 1
        self.inv2 = self.inv3 = None          # full inverses (path scoring) are not maintained
        return True

    def greedy_utilities(self, ent_a, static_std, mobile_std):
        """ut_i = ent_a_i + H(Sigma_{Abar \\ i}) - H(Sigma + D + Delta_i e_i e_i^T) for every location
        (agent.py:330-339); ent_a is -inf where the location is already static."""
        ss2, ms2 = static_std ** 2, mobile_std ** 2
        vboth = 1.0 / (1.0 / ss2 + 1.0 / ms2)
        is_new = self.pos2 >= 0
        d2 = self.diag2[self.pos2.clamp_min(0).long()]
        ent_abar = torch.where(is_new, (self.n_abar - 1) * CONST + 0.5 * (self.ld2 + torch.log(d2)),
                               self.n_abar * CONST + 0.5 * self.ld2)
        delta = torch.where(is_new, torch.full_like(d2, ss2), torch.full_like(d2, vboth - ms2))
        ent_all = self.n * CONST + 0.5 * (self.ld3 + torch.log1p(delta * self.diag3))
        return ent_a + ent_abar - ent_all

    def path_utilities(self, ent_a, idx, skip, static_std, mobile_std):
        """ut_p = ent_a_p + H(Sigma_{Abar \\ C_p}) - H(Sigma + D + E_C Delta E_C^T) (agent.py:388-397)."""
        ss2, ms2 = static_std ** 2, mobile_std ** 2
        vboth = 1.0 / (1.0 / ss2 + 1.0 / ms2)
        B, k = idx.shape
        out = torch.empty((B, 3), dtype=torch.float64, device=idx.device)
        nwork, work = scratch_for(self, "_largework", _lib.lib.algp_mi_terms_large_work_doubles(k, B), k * (k + 1), idx.device)
        call("algp_mi_terms_large", ptr(self.inv2), self.inv2.stride(0) if self.inv2 is not None else 0, ptr(self.pos2),
             ptr(self.inv3), self.inv3.stride(0), ptr(idx), k, B, ptr(skip), float(ms2), float(vboth - ss2), ptr(out),
             ptr(work), nwork, stream())
        ent_abar = (self.n_abar - out[:, 1]) * CONST + 0.5 * (self.ld2 + out[:, 0])
        ent_all = self.n * CONST + 0.5 * (self.ld3 + out[:, 2])
        return ent_a + ent_abar - ent_all
