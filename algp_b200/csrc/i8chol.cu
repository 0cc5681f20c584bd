// Cholesky factor AND its inverse with the O(n^3) work on the INT8 tensor cores.
//
// algp_potrf + algp_trtri spend their time in fp64 DMMA GEMMs (35 TFLOP/s).  The same factor and inverse follow
// from a recursive 2 x 2 splitting whose four products per level are large-K GEMMs, which the exact INT8 digit
// GEMM of i8.cu runs at ~100 TFLOP/s fp64-equivalent:
//     (L11, Linv11) = rec(A11)
//     L21   = A21 Linv11^T                         k <= j     (Linv11 lower)
//     A22  -= L21 L21^T                            lower tiles
//     (L22, Linv22) = rec(A22)
//     W^T   = (L21 T1^T)^T,  T1 = Linv11^T         k >= j     (stored transposed by the epilogue)
//     Linv21 = -Linv22 (W^T)^T                     k <= i     (Linv22 lower)
// Blocks of `base` rows or fewer use the DMMA path (potrf_block + algp_trtri: latency-bound there anyway).
// Replaces the float32 `inv` of the reference (utils.py:300) like algp_potrf / algp_trtri do.
#include "i8.cuh"

int potrf_block(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int* info_dev, int col0, cudaStream_t st);
extern "C" int algp_trtri(const double* L, int64_t npad, int64_t ld, double* Linv, int64_t ldi, double* work, int zero_upper,
                          void* stream);
extern "C" int64_t algp_trtri_work_doubles(int64_t npad);

namespace {

__global__ void transpose_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = src[(int64_t)(by + r) * lds + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) dst[(int64_t)(bx + r) * ldd + by + threadIdx.x] = tile[threadIdx.x][r];
}

__global__ void zero_upper_kernel(double* __restrict__ M, int64_t ld, int nb) {
  // zero the strictly-upper 128 x 128 blocks so Linv is a clean lower-triangular matrix
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj <= bi) return;
  double* blk = M + (int64_t)bi * ALGP_BLK * ld + (int64_t)bj * ALGP_BLK;
  for (int e = threadIdx.x; e < ALGP_BLK * ALGP_BLK / 2; e += blockDim.x) {
    const int r = e / (ALGP_BLK / 2), c2 = (e % (ALGP_BLK / 2)) * 2;
    *reinterpret_cast<double2*>(blk + (int64_t)r * ld + c2) = make_double2(0.0, 0.0);
  }
}

#define I8C_MAXD 12
struct Work {
  int S;
  int64_t base;
  double* t1[I8C_MAXD];  // per recursion depth, [h x h] fp64: Linv11^T, then W^T
  double* trtri_work;    // base-case trtri scratch
  int8_t* da; double* sa;   // left-operand digits + scales
  int8_t* db; double* sb;   // right-operand digits + scales
  uint8_t* ma; uint8_t* mb; // their plane-occupancy masks (spatially ordered points: far blocks are empty)
  int8_t* da2; double* sa2; int8_t* db2; double* sb2; uint8_t* ma2; uint8_t* mb2;   // the side stream's own set
  int* info;
  cudaStream_t st, side;
  cudaEvent_t ev_l21[I8C_MAXD], ev_w[I8C_MAXD];
};

// W^T = (L21 Linv11)^T only needs L21 and Linv11, not the factor of A22: it runs on a low-priority side stream and
// fills the SMs that the latency-bound base cases of the A22 recursion leave idle.
struct SideAux {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_l21[I8C_MAXD], ev_w[I8C_MAXD];
};
SideAux g_side[64];

int side_get(SideAux** out) {
  int dev = 0;
  ALGP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return ALGP_ERR_UNSUPPORTED;
  SideAux& a = g_side[dev];
  if (!a.side) {
    int prio_lo = 0, prio_hi = 0;
    ALGP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    ALGP_CUDA(cudaStreamCreateWithPriority(&a.side, cudaStreamNonBlocking, prio_lo));
    for (int d = 0; d < I8C_MAXD; ++d) {
      ALGP_CUDA(cudaEventCreateWithFlags(&a.ev_l21[d], cudaEventDisableTiming));
      ALGP_CUDA(cudaEventCreateWithFlags(&a.ev_w[d], cudaEventDisableTiming));
    }
  }
  *out = &a;
  return ALGP_OK;
}

int rec(double* A, int64_t n, int64_t ld, double* Li, int64_t ldi, int col0, const Work& w, int depth) {
  int rc;
  if (n <= w.base) {
    if ((rc = potrf_block(A, n, ld, Li, ldi, w.info, col0, w.st))) return rc;
    return algp_trtri(A, n, ld, Li, ldi, w.trtri_work, 0, (void*)w.st);
  }
  // split at a multiple of 128 (the left half is the larger one when n/128 is odd)
  const int64_t h = ((n / ALGP_BLK + 1) / 2) * ALGP_BLK, h2 = n - h;
  double* A21 = A + h * ld;
  double* A22 = A + h * (ld + 1);
  double* Li21 = Li + h * ldi;
  double* Li22 = Li + h * (ldi + 1);
  if (depth >= I8C_MAXD) return ALGP_ERR_UNSUPPORTED;
  if ((rc = rec(A, h, ld, Li, ldi, col0, w, depth + 1))) return rc;

  // L21 = A21 Linv11^T   [h2 x h], k <= j
  if ((rc = i8_split(A21, h2, h, ld, w.S, I8_TM, w.da, w.sa, w.ma, w.st))) return rc;
  if ((rc = i8_split(Li, h, h, ldi, w.S, I8_TN, w.db, w.sb, w.mb, w.st))) return rc;
  I8Gemm g = i8_gemm_default();
  g.MT = (int)(h2 / I8_TM); g.NT = (int)(h / I8_TN); g.kchunks = (int)(h / I8_KC);
  g.a_tiles = w.da; g.scale_a = w.sa; g.b_tiles = w.db; g.scale_b = w.sb;
  g.a_mask = w.ma; g.b_mask = w.mb; g.mask_ld = i8_mask_ld((int64_t)g.kchunks * I8_KC);
  g.kend_rule = I8_KE_NT; g.nt_desc = 1;
  g.C = A21; g.ldc = ld;
  if ((rc = i8_gemm(g, w.S, w.st))) return rc;

  // side stream: W^T = (L21 T1^T)^T, T1 = Linv11^T   [W is h2 x h], k >= j -- concurrent with the A22 branch below
  {
    double* t1s = w.t1[depth];
    ALGP_CUDA(cudaEventRecord(w.ev_l21[depth], w.st));
    ALGP_CUDA(cudaStreamWaitEvent(w.side, w.ev_l21[depth], 0));
    if ((rc = i8_split(A21, h2, h, ld, w.S, I8_TM, w.da2, w.sa2, w.ma2, w.side))) return rc;
    transpose_kernel<<<dim3((unsigned)(h / 32), (unsigned)(h / 32)), dim3(32, 8), 0, w.side>>>(Li, ldi, t1s, h);
    ALGP_LAUNCH_CHECK();
    if ((rc = i8_split(t1s, h, h, h, w.S, I8_TN, w.db2, w.sb2, w.mb2, w.side))) return rc;
    I8Gemm gs = i8_gemm_default();
    gs.MT = (int)(h2 / I8_TM); gs.NT = (int)(h / I8_TN); gs.kchunks = (int)(h / I8_KC);
    gs.a_tiles = w.da2; gs.scale_a = w.sa2; gs.b_tiles = w.db2; gs.scale_b = w.sb2;
    gs.a_mask = w.ma2; gs.b_mask = w.mb2; gs.mask_ld = i8_mask_ld((int64_t)gs.kchunks * I8_KC);
    gs.kbeg_rule = I8_KB_NT;
    gs.C = t1s; gs.ldc = h2; gs.transposed = 1;        // t1 <- W^T [h x h2] (its digits were taken above)
    if ((rc = i8_gemm(gs, w.S, w.side))) return rc;
    ALGP_CUDA(cudaEventRecord(w.ev_w[depth], w.side));
  }

  // A22 -= L21 L21^T     [h2 x h2], K = h, lower tiles
  if ((rc = i8_split(A21, h2, h, ld, w.S, I8_TM, w.da, w.sa, w.ma, w.st))) return rc;
  if ((rc = i8_split(A21, h2, h, ld, w.S, I8_TN, w.db, w.sb, w.mb, w.st))) return rc;
  g = i8_gemm_default();
  g.MT = (int)(h2 / I8_TM); g.NT = (int)(h2 / I8_TN); g.kchunks = (int)(h / I8_KC);
  g.a_tiles = w.da; g.scale_a = w.sa; g.b_tiles = w.db; g.scale_b = w.sb;
  g.a_mask = w.ma; g.b_mask = w.mb; g.mask_ld = i8_mask_ld((int64_t)g.kchunks * I8_KC);
  g.lower_only = 1;
  g.C = A22; g.ldc = ld; g.alpha = -1.0; g.beta = 1.0;
  if ((rc = i8_gemm(g, w.S, w.st))) return rc;

  if ((rc = rec(A22, h2, ld, Li22, ldi, col0 + (int)h, w, depth + 1))) return rc;

  ALGP_CUDA(cudaStreamWaitEvent(w.st, w.ev_w[depth], 0));      // W^T from the side stream
  double* t1 = w.t1[depth];

  // Linv21 = -Linv22 (W^T)^T   [h2 x h], K = h2, k <= i
  if ((rc = i8_split(Li22, h2, h2, ldi, w.S, I8_TM, w.da, w.sa, w.ma, w.st))) return rc;
  if ((rc = i8_split(t1, h, h2, h2, w.S, I8_TN, w.db, w.sb, w.mb, w.st))) return rc;
  g = i8_gemm_default();
  g.MT = (int)(h2 / I8_TM); g.NT = (int)(h / I8_TN); g.kchunks = (int)(h2 / I8_KC);
  g.a_tiles = w.da; g.scale_a = w.sa; g.b_tiles = w.db; g.scale_b = w.sb;
  g.a_mask = w.ma; g.b_mask = w.mb; g.mask_ld = i8_mask_ld((int64_t)g.kchunks * I8_KC);
  g.kend_rule = I8_KE_MT; g.mt_desc = 1;
  g.C = Li21; g.ldc = ldi; g.alpha = -1.0;
  return i8_gemm(g, w.S, w.st);
}

int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }


// sizes of the per-depth W^T buffers: the left child is the larger one, so its half is the bound of every node of a depth
int depth_halves(int64_t npad, int64_t base, int64_t* halves) {
  int d = 0;
  int64_t n = npad;
  while (n > base && d < I8C_MAXD) {
    const int64_t h = ((n / ALGP_BLK + 1) / 2) * ALGP_BLK;
    halves[d++] = h;
    n = h;
  }
  return d;
}

}  // namespace

// scratch for algp_potrf_inv_i8, in bytes
extern "C" int64_t algp_potrf_inv_i8_work_bytes(int64_t npad, int nslices, int64_t base) {
  if (npad <= 0) return 256;
  const int64_t h = ((npad / ALGP_BLK + 1) / 2) * ALGP_BLK;
  int64_t halves[I8C_MAXD];
  const int nd = depth_halves(npad, base, halves);
  int64_t b = 0;
  for (int d = 0; d < nd; ++d) b += align256(halves[d] * halves[d] * 8);      // t1 per depth
  b += align256(algp_trtri_work_doubles(base < npad ? base : npad) * 8 + 16);
  b += 4 * align256(h * h * (int64_t)nslices);                // da, db, da2, db2
  b += 4 * align256(h * 8);                                   // sa, sb, sa2, sb2
  b += 4 * align256((h / I8_TN + 1) * i8_mask_ld(h) + 8);     // occupancy masks
  return b + 256;
}

// L (lower triangle of A) and Linv = L^-1 (lower triangular, strictly-upper 128-blocks zeroed) of the SPD matrix A
// [npad x npad], npad % 128 == 0.  Blocks larger than `base` rows are split recursively and their products run as
// exact INT8 digit GEMMs with `nslices` planes (8: fp64-grade); *info_dev as algp_potrf.
extern "C" int algp_potrf_inv_i8(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int nslices, int64_t base,
                                 void* work, int64_t work_bytes, int* info_dev, void* stream) {
  if (!A || !Linv || !info_dev || npad < 0 || npad % ALGP_BLK || ld < npad || ldi < npad || (ld & 1) || (ldi & 1) ||
      nslices < 2 || nslices > I8_MAX_S || base < ALGP_BLK || base % ALGP_BLK || npad > 65536)
    return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  ALGP_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (npad == 0) return ALGP_OK;
  if (!work || work_bytes < algp_potrf_inv_i8_work_bytes(npad, nslices, base) || ((uintptr_t)work & 15)) return ALGP_ERR_INVALID;
  const int nb = (int)(npad / ALGP_BLK);
  if (nb > 1) {
    zero_upper_kernel<<<dim3(nb, nb), 256, 0, st>>>(Linv, ldi, nb);
    ALGP_LAUNCH_CHECK();
  }
  const int64_t h = ((npad / ALGP_BLK + 1) / 2) * ALGP_BLK;
  char* p = (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  SideAux* sx = nullptr;
  int rc = side_get(&sx);
  if (rc) return rc;
  Work w;
  w.S = nslices; w.base = base; w.info = info_dev; w.st = st; w.side = sx->side;
  for (int d = 0; d < I8C_MAXD; ++d) {
    w.ev_l21[d] = sx->ev_l21[d];
    w.ev_w[d] = sx->ev_w[d];
    w.t1[d] = nullptr;
  }
  int64_t halves[I8C_MAXD];
  const int nd = depth_halves(npad, base, halves);
  for (int d = 0; d < nd; ++d) { w.t1[d] = (double*)p; p += align256(halves[d] * halves[d] * 8); }
  w.trtri_work = (double*)p; p += align256(algp_trtri_work_doubles(base < npad ? base : npad) * 8 + 16);
  const int64_t dig = align256(h * h * (int64_t)nslices), scl = align256(h * 8), msk = align256((h / I8_TN + 1) * i8_mask_ld(h) + 8);
  w.da = (int8_t*)p; p += dig;  w.db = (int8_t*)p; p += dig;  w.da2 = (int8_t*)p; p += dig;  w.db2 = (int8_t*)p; p += dig;
  w.sa = (double*)p; p += scl;  w.sb = (double*)p; p += scl;  w.sa2 = (double*)p; p += scl;  w.sb2 = (double*)p; p += scl;
  w.ma = (uint8_t*)p; p += msk; w.mb = (uint8_t*)p; p += msk; w.ma2 = (uint8_t*)p; p += msk; w.mb2 = (uint8_t*)p; p += msk;
  // the side stream must not start before the caller's stream has produced A (and everything queued before this call)
  return rec(A, npad, ld, Linv, ldi, 0, w, 0);
}
