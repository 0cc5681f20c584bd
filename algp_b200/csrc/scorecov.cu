// Candidate scoring from a resident posterior covariance ("If the build precomputes P (allowed)").
//
// score.cu streams k rows of Wt (N doubles each) per candidate to form the Gram matrix G = Wt_C Wt_C^T: 262 KB per
// candidate at N = 4096, k = 8, and the kernel sits on the L2 -> SM delivery roof.  When the same factored base set
// is scored again and again (many candidate batches per planning step, agent.py:373-400), the n x n posterior
// covariance
//     P = Sigma + sigma_n^2 I - Wt Wt^T                      (lower triangle; one SYRK per base set)
// is cheaper to keep: a candidate then needs the k(k+1)/2 entries P[c_i][c_j] and nothing else --
//     logdet(I + D P_CC D),  D = diag(sqrt(delta))          (same matrix as score.cu, same elimination order)
// i.e. 36 eight-byte gathers instead of 32 768 doubles for k = 8.
//
//   score_cov_k8      one THREAD per candidate, k <= 8: the 36 entries and the elimination live in registers
//   score_cov_generic one CTA per candidate, k <= 128: the k x k matrix in shared memory (as score_sets_generic)
//
// Slot semantics are those of score.cu (reference agent.py:377-400): idx < 0 or delta <= 0 or skip[idx] = empty slot,
// duplicates inside a set count once (the first), score = H_base + n_new CONST + (logdet - sum log((pi0+delta)/pi0)) / 2.
#include "common.cuh"
#include <math.h>

#define ALGP_CONST 1.4189385332046727   // 0.5*log(2*pi*e), utils.py:10

namespace {

struct CovArgs {
  const double* P;         // [n x ldp] lower triangle of the posterior covariance of the base set
  int64_t ldp;
  const double* pi0;       // [n] base precisions (0 = unsampled)
  const int32_t* idx;      // [B x k]   (-1 = empty slot)
  const double* delta;     // [B x k] or null
  double delta_scalar;
  const uint8_t* skip;     // [n] or null
  int k;
  int64_t B;
  double H_base;
  double* scores;          // [B]
};

__device__ __forceinline__ double ldg_nc(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(128) score_cov_k8_kernel(const CovArgs a) {
  const int64_t cand = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= a.B) return;
  int ix[8];
  double sq[8];
  double term = 0.0, nnew = 0.0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    int id = -1;
    double dl = 0.0;
    if (s < a.k) {
      id = a.idx[cand * a.k + s];
      dl = a.delta ? a.delta[cand * a.k + s] : a.delta_scalar;
    }
    bool act = id >= 0 && dl > 0.0;
    if (act && a.skip && a.skip[id]) act = false;
#pragma unroll
    for (int q = 0; q < s; ++q)
      if (ix[q] == id) act = false;          // ix[q] >= 0 only for active slots: duplicates count once
    ix[s] = act ? id : -1;
    sq[s] = act ? sqrt(dl) : 0.0;
    if (act) {
      const double p0 = a.pi0[id];
      term += log(p0 + dl) - (p0 > 0.0 ? log(p0) : 0.0);
      nnew += (p0 > 0.0) ? 0.0 : 1.0;
    }
  }
  // all gathers first (36 independent loads in flight), then the arithmetic
  double m[36];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      m[e] = 0.0;
      if (ix[i] >= 0 && ix[j] >= 0) {
        const int hi = ix[i] > ix[j] ? ix[i] : ix[j], lo = ix[i] > ix[j] ? ix[j] : ix[i];
        m[e] = ldg_nc(a.P + (int64_t)hi * a.ldp + lo);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      m[e] = m[e] * sq[i] * sq[j] + ((i == j) ? 1.0 : 0.0);
    }
  }
  // un-normalised elimination on the lower triangle (the order of score_sets_k8_kernel)
  double logdet = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double piv = m[c * (c + 1) / 2 + c];
#pragma unroll
    for (int r = c + 1; r < 8; ++r) {
      const double f = m[r * (r + 1) / 2 + c] / piv;
#pragma unroll
      for (int cc = c + 1; cc <= r; ++cc) m[r * (r + 1) / 2 + cc] = fma(-f, m[cc * (cc + 1) / 2 + c], m[r * (r + 1) / 2 + cc]);
    }
    logdet += log(piv);
  }
  a.scores[cand] = a.H_base + nnew * ALGP_CONST + 0.5 * (logdet - term);
}

#define SC_MAXK 128
__global__ void __launch_bounds__(256) score_cov_generic_kernel(const CovArgs a) {
  extern __shared__ __align__(16) double sc_smem[];
  const int k = a.k, pitch = k + 1;
  double* M = sc_smem;                       // [k][k+1]
  double* sqd = M + k * pitch;               // [k]
  int* sidx = (int*)(sqd + k);               // [k]
  __shared__ double s_t[8], s_n[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t cand = blockIdx.x; cand < a.B; cand += gridDim.x) {
    __syncthreads();
    for (int s = tid; s < k; s += 256) {
      const int id = a.idx[cand * k + s];
      const double dl = a.delta ? a.delta[cand * k + s] : a.delta_scalar;
      bool act = id >= 0 && dl > 0.0;
      if (act && a.skip && a.skip[id]) act = false;
      for (int q = 0; q < s && act; ++q) {
        const int oid = a.idx[cand * k + q];
        const double odl = a.delta ? a.delta[cand * k + q] : a.delta_scalar;
        if (oid == id && odl > 0.0) act = false;
      }
      sidx[s] = act ? id : -1;
      sqd[s] = act ? sqrt(dl) : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += 256) {
      const int r = e / k, c = e % k;
      if (c > r) continue;
      double m = (r == c) ? 1.0 : 0.0;
      const int ir = sidx[r], ic = sidx[c];
      if (ir >= 0 && ic >= 0) {
        const int hi = ir > ic ? ir : ic, lo = ir > ic ? ic : ir;
        m = fma(ldg_nc(a.P + (int64_t)hi * a.ldp + lo), sqd[r] * sqd[c], m);
      }
      M[r * pitch + c] = m;
    }
    __syncthreads();
    double logdet = 0.0;
    for (int c = 0; c < k; ++c) {
      const double piv = M[c * pitch + c];
      const double inv = 1.0 / piv;
      if (tid == 0) logdet += log(piv);
      const int rem = k - 1 - c;
      for (int e = tid; e < rem * rem; e += 256) {
        const int r = c + 1 + e / rem, cc = c + 1 + e % rem;
        if (cc <= r) M[r * pitch + cc] = fma(-M[r * pitch + c] * inv, M[cc * pitch + c], M[r * pitch + cc]);
      }
      __syncthreads();
    }
    double term = 0.0, nnew = 0.0;
    for (int s = tid; s < k; s += 256)
      if (sidx[s] >= 0) {
        const double p0 = a.pi0[sidx[s]];
        const double dl = sqd[s] * sqd[s];
        term += log(p0 + dl) - (p0 > 0.0 ? log(p0) : 0.0);
        nnew += (p0 > 0.0) ? 0.0 : 1.0;
      }
    term = warp_sum(term);
    nnew = warp_sum(nnew);
    if (lane == 0) { s_t[warp] = term; s_n[warp] = nnew; }
    __syncthreads();
    if (tid == 0) {
      double tt = 0.0, nn = 0.0;
      for (int w = 0; w < 8; ++w) { tt += s_t[w]; nn += s_n[w]; }
      a.scores[cand] = a.H_base + nn * ALGP_CONST + 0.5 * (logdet - tt);
    }
  }
}

// ---------------------------------------------------------------------------
// Keeping P alive across commits.  An acquisition is a rank-1 downdate of the posterior covariance, stored as one
// new column of Wt (score.cu: append / append_block).  The resident P follows by
//     P <- P - sum_c w_c w_c^T        over the k columns appended since P was last brought up to date
// on the lower triangle: one pass over P (8 n^2 / 2 bytes read and written, HBM bound) for up to 16 columns.
// 64 x 64 tiles of the lower triangle (diagonal tiles whole: the upper triangle of P is never read), 256 threads,
// 4 x 4 elements per thread, the two 64 x k slabs of Wt staged in shared memory.
// ---------------------------------------------------------------------------
#define CD_T 64
#define CD_MAXK 32
#define CD_PITCH (CD_MAXK + 4)
__global__ void __launch_bounds__(256) cov_downdate_kernel(double* __restrict__ P, int64_t ldp, int64_t n, const double* __restrict__ Wt,
                                                           int64_t ldw, int col0, int k, int tiles) {
  // k-major slabs of the two row ranges; pitch = 4 (mod 16) doubles: the 8-byte fragment reads of a half-warp
  // (4 rows x 4 k) hit 32 distinct banks
  __shared__ double wi[CD_T][CD_PITCH], wj[CD_T][CD_PITCH];
  // blockIdx -> (ti, tj), tj <= ti
  const int64_t b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((int64_t)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((int64_t)ti * (ti + 1) / 2 > b) --ti;
  const int tj = (int)(b - (int64_t)ti * (ti + 1) / 2);
  (void)tiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;               // warp sub-tile: rows 16 wm .. +15, columns 32 wn .. +31
  // The tile of P is requested FIRST, straight into the DMMA accumulator fragments (lane (g, t) holds columns 2t, 2t+1
  // of row g of every 8x8 block: 8 independent 16-byte loads): its DRAM latency overlaps the slab loads and the
  // barrier, and the rank-k product then runs as P - Wi Wj^T on the FP64 tensor cores with the accumulators starting
  // at P (B fragments negated).  The first version did the product with DFMAs on 4x4 register patches: 8 shared
  // operand loads per 16 DFMAs per thread and column, bound by shared-memory delivery (0.016 ms per column at
  // n = 16 384); a fragment is shared by the 8 lanes of a row / column inside the tensor core instead.
  double2 c[2][4];
  int64_t grow[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    grow[i] = (int64_t)ti * CD_T + wm * 16 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gc = (int64_t)tj * CD_T + wn * 32 + j * 8 + 2 * t;
      c[i][j] = make_double2(0.0, 0.0);
      if (grow[i] < n && gc + 1 < n) c[i][j] = *reinterpret_cast<const double2*>(P + grow[i] * ldp + gc);
      else if (grow[i] < n && gc < n) c[i][j].x = P[grow[i] * ldp + gc];
    }
  }
  const int k4 = (k + 3) & ~3;
  for (int e = tid; e < CD_T * k4; e += 256) {
    const int r = e / k4, q = e % k4;
    const int64_t gi = (int64_t)ti * CD_T + r, gjr = (int64_t)tj * CD_T + r;
    wi[r][q] = (gi < n && q < k) ? Wt[gi * ldw + col0 + q] : 0.0;
    wj[r][q] = (gjr < n && q < k) ? Wt[gjr * ldw + col0 + q] : 0.0;
  }
  __syncthreads();
  for (int q0 = 0; q0 < k4; q0 += 4) {
    double af[2], bf[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) af[i] = wi[wm * 16 + i * 8 + g][q0 + t];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = -wj[wn * 32 + j * 8 + g][q0 + t];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(c[i][j].x, c[i][j].y, af[i], bf[j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (grow[i] >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gc = (int64_t)tj * CD_T + wn * 32 + j * 8 + 2 * t;
      if (gc + 1 < n) *reinterpret_cast<double2*>(P + grow[i] * ldp + gc) = c[i][j];
      else if (gc < n) P[grow[i] * ldp + gc] = c[i][j].x;
    }
  }
}

}  // namespace

// P[i][j] -= sum_{c < k} Wt[i][col0 + c] Wt[j][col0 + c] on the lower triangle (j <= i) of the resident posterior
// covariance P [n x ldp]: the k columns Wt gained through algp_append / algp_append_block since P was current.
// k <= algp_cov_downdate_max_cols() = 32 per call (one pass over the lower triangle each); ldp even and P 16-byte aligned.
extern "C" int algp_cov_downdate_max_cols(void) { return CD_MAXK; }

extern "C" int algp_cov_downdate(double* P, int64_t ldp, int64_t n, const double* Wt, int64_t ldw, int64_t col0, int k,
                                 void* stream) {
  if (!P || !Wt || n < 1 || ldp < n || (ldp & 1) || ((uintptr_t)P & 15) || col0 < 0 || k < 1 || k > CD_MAXK || col0 + k > ldw)
    return ALGP_ERR_INVALID;
  const int64_t t = (n + CD_T - 1) / CD_T;
  const int64_t blocks = t * (t + 1) / 2;
  cov_downdate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, ldp, n, Wt, ldw, (int)col0, k, (int)t);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}


// scores[c] = H(base set + candidate set c) from the lower triangle of the posterior covariance P [n x ldp] of the
// base set (P = Sigma + sigma_n^2 I - Wt Wt^T): the restructured form of the slogdet loops of agent.py:373-400 /
// utils.py:188-194 with the Schur complement read instead of recomputed.  k <= 128.
extern "C" int algp_score_sets_cov(const double* P, int64_t ldp, const double* pi0, const int32_t* idx, const double* delta,
                                   double delta_scalar, const uint8_t* skip, int k, int64_t B, double H_base, double* scores,
                                   void* stream) {
  if (!P || !pi0 || !idx || !scores || k < 1 || B < 0 || ldp < 1) return ALGP_ERR_INVALID;
  if (k > SC_MAXK) return ALGP_ERR_UNSUPPORTED;
  if (B == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CovArgs a;
  a.P = P; a.ldp = ldp; a.pi0 = pi0; a.idx = idx; a.delta = delta; a.delta_scalar = delta_scalar; a.skip = skip;
  a.k = k; a.B = B; a.H_base = H_base; a.scores = scores;
  if (k <= 8) {
    score_cov_k8_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(a);
  } else {
    const size_t smem = (size_t)k * (k + 1) * sizeof(double) + (size_t)k * sizeof(double) + (size_t)k * sizeof(int) + 16;
    static AlgpPerDevice configured;
    if (configured.raise(1)) {
      ALGP_CUDA(cudaFuncSetAttribute(score_cov_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)SC_MAXK * (SC_MAXK + 1) * 8 + SC_MAXK * 12 + 16)));
    }
    int dev = 0, sms = 148;
    ALGP_CUDA(cudaGetDevice(&dev));
    ALGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t cap = (int64_t)sms * 4;
    score_cov_generic_kernel<<<(unsigned)(B < cap ? B : cap), 256, smem, st>>>(a);
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
