// Mutual-information criterion (reference agent.py:330-339 greedy, 388-397 best_path):
//     ut = H(Sigma_SS + D_S) + H(Sigma_AbarAbar) - H(Sigma + D),   Abar = unsampled locations,
// i.e. I(y_S ; f_Abar).  The reference pays two more n x n slogdets PER CANDIDATE.  With
// A2 = Sigma_AbarAbar and A3 = Sigma + D factored once per base set, a candidate that touches the
// locations C only needs k x k blocks of the two inverses:
//     logdet Sigma_{Abar\C}      = logdet A2 + logdet [A2^-1]_CC                       (C new in Abar)
//     logdet(A3 + E_C Delta E_C^T) = logdet A3 + sum log|Delta| + logdet|Delta^-1 + [A3^-1]_CC|
// (Delta = change of the per-location noise variance; it is negative where a static-only location
// gains a mobile reading, which makes Delta^-1 + [A3^-1]_CC symmetric quasi-definite: a pivot-free
// LDL^T exists and log|det| is the sum of log|pivot|.)
#include "common.cuh"

// out[c] = sum_{r>=c} M[r][c]^2 : diag(A^-1) from the inverse factor; chunked partials, fixed order
#define MI_CHUNK 256
__global__ void colsumsq_lower_kernel(const double* __restrict__ M, int64_t n, int64_t ld, double* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * MI_CHUNK;
  if (c >= n) return;
  int64_t rb = r0 > c ? r0 : c;
  int64_t re = r0 + MI_CHUNK < n ? r0 + MI_CHUNK : n;
  double s0 = 0.0, s1 = 0.0;
  int64_t r = rb;
  for (; r + 1 < re; r += 2) {
    const double a = M[r * ld + c], b = M[(r + 1) * ld + c];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (r < re) { const double a = M[r * ld + c]; s0 = fma(a, a, s0); }
  partial[(int64_t)blockIdx.y * n + c] = s0 + s1;
}
__global__ void mi_colsum_kernel(const double* __restrict__ partial, int64_t n, int chunks, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += partial[(int64_t)k * n + c];
  out[c] = s;
}

extern "C" int64_t algp_colsumsq_work_doubles(int64_t n) { return ((n + MI_CHUNK - 1) / MI_CHUNK) * n; }

extern "C" int algp_colsumsq_lower(const double* M, int64_t n, int64_t ld, double* out, double* work, void* stream) {
  if (!M || !out || !work || n < 0 || ld < n) return ALGP_ERR_INVALID;
  if (n == 0) return ALGP_OK;
  const int chunks = (int)((n + MI_CHUNK - 1) / MI_CHUNK);
  cudaStream_t st = (cudaStream_t)stream;
  colsumsq_lower_kernel<<<dim3((unsigned)((n + 127) / 128), chunks), 128, 0, st>>>(M, n, ld, work);
  ALGP_LAUNCH_CHECK();
  mi_colsum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(work, n, chunks, out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// Rank-1 maintenance of diag(A^-1) and logdet A across greedy picks: the mutual-information
// criterion needs diag((Sigma_AbarAbar)^-1) and diag((Sigma + D)^-1); a pick changes Sigma_AbarAbar by deleting one
// row / column and Sigma + D by one diagonal entry, so both inverses change by a rank-1 term instead of being
// re-factorised (reference agent.py:330-339 recomputes two n x n slogdets per CANDIDATE; round 1 re-factorised per pick).
//   col          column j of the ORIGINAL inverse (A0^-1 e_j, two triangular gemv passes over the inverse factor)
//   U, coef, t   the t earlier corrections: current inverse = A0^-1 - sum_s coef[s] U[s] U[s]^T
//   mode 0       delete row / column j:       inv' = inv - c c^T / c_j,            logdet += log c_j
//   mode 1       A += delta e_j e_j^T:        inv' = inv - g c c^T, g = delta / (1 + delta c_j), logdet += log1p(delta c_j)
// with c = column j of the CURRENT inverse.  Updates diag in place, appends U[t] = c, coef[t] = g.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inv_rank1_update_kernel(const double* __restrict__ col, int64_t n, double* __restrict__ U,
                                                               int64_t ldu, double* __restrict__ coef, int t, int64_t j, int mode,
                                                               double delta, double* __restrict__ diag, double* __restrict__ logdet) {
  __shared__ double w[64];            // coef[s] * U[s][j]
  __shared__ double cj_s;
  if (threadIdx.x < t) w[threadIdx.x] = coef[threadIdx.x] * U[(int64_t)threadIdx.x * ldu + j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double cj = col[j];
    for (int s = 0; s < t; ++s) cj -= w[s] * U[(int64_t)s * ldu + j];
    cj_s = cj;
  }
  __syncthreads();
  const double cj = cj_s;
  const double g = mode == 0 ? 1.0 / cj : delta / (1.0 + delta * cj);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double c = col[i];
    for (int s = 0; s < t; ++s) c -= w[s] * U[(int64_t)s * ldu + i];
    U[(int64_t)t * ldu + i] = c;
    diag[i] -= g * c * c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // coef[t] is only read by LATER launches (slots s < t here), so writing it now does not race
    coef[t] = g;
    logdet[0] += mode == 0 ? log(cj) : log1p(delta * cj);
  }
}

extern "C" int algp_inv_rank1_update(const double* col, int64_t n, double* U, int64_t ldu, double* coef, int t, int64_t j,
                                     int mode, double delta, double* diag, double* logdet, void* stream) {
  if (!col || !U || !coef || !diag || !logdet || n < 1 || ldu < n || t < 0 || t >= 64 || j < 0 || j >= n || (mode != 0 && mode != 1))
    return ALGP_ERR_INVALID;
  const unsigned grid = (unsigned)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  inv_rank1_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(col, n, U, ldu, coef, t, j, mode, delta, diag, logdet);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

struct MiArgs {
  const double* inv2; int64_t ld2;      // A2^-1 (lower triangle valid), indexed by position in Abar
  const int32_t* pos2;                  // [n]: position of a location in Abar, -1 if sampled
  const double* inv3; int64_t ld3;      // A3^-1 (lower triangle valid), indexed by location
  const int32_t* idx; int k; int64_t B; // [B x k] candidate locations (-1 = empty)
  const uint8_t* skip;                  // [n] or null: already-mobile locations add nothing
  double delta_new;                     // variance change of a brand-new location
  double delta_old;                     // variance change of an already-sampled (static-only) location
  double* out;                          // [B x 3]: logdet [A2^-1]_CC, number of new locations, sum log|Delta| + logdet|...|
  double* work;                         // k > 128: [grid][k][k+1] elimination scratch in global memory
};

// un-normalised elimination of the lower triangle of a kk x kk matrix in shared memory; returns sum log|pivot|
__device__ double mi_logabsdet(double* M, int kk, int pitch, int tid) {
  double acc = 0.0;
  for (int c = 0; c < kk; ++c) {
    const double piv = M[c * pitch + c];
    const double inv = 1.0 / piv;
    if (tid == 0) acc += log(fabs(piv));
    const int rem = kk - 1 - c;
    for (int e = tid; e < rem * rem; e += blockDim.x) {
      const int r = c + 1 + e / rem, cc = c + 1 + e % rem;
      if (cc <= r) M[r * pitch + cc] = fma(-M[r * pitch + c] * inv, M[cc * pitch + c], M[r * pitch + cc]);
    }
    __syncthreads();
  }
  return acc;   // valid in thread 0
}

template <bool GLOBAL_M>
__global__ void __launch_bounds__(256) mi_terms_kernel(const MiArgs a) {
  extern __shared__ __align__(16) double mi_smem[];
  const int k = a.k, pitch = k + 1;
  double* M = GLOBAL_M ? a.work + (size_t)blockIdx.x * k * pitch : mi_smem;   // [k][k+1]
  double* dl = GLOBAL_M ? mi_smem : M + k * pitch;                            // [k] Delta of the slot (0 = inactive)
  int* loc = (int*)(dl + k);              // [k] location, -1 inactive
  int* p2 = loc + k;                      // [k] position in Abar, -1 if not new
  const int tid = threadIdx.x;
  for (int64_t cand = blockIdx.x; cand < a.B; cand += gridDim.x) {
    __syncthreads();
    for (int s = tid; s < k; s += blockDim.x) {
      int ix = a.idx[cand * k + s];
      bool act = ix >= 0 && !(a.skip && a.skip[ix]);
      for (int q = 0; q < s && act; ++q)
        if (a.idx[cand * k + q] == ix) act = false;          // duplicates are idempotent (agent.py:377)
      loc[s] = act ? ix : -1;
      const int pp = act ? a.pos2[ix] : -1;
      p2[s] = pp;
      dl[s] = act ? (pp >= 0 ? a.delta_new : a.delta_old) : 0.0;
    }
    __syncthreads();
    // ---- term 2: logdet [A2^-1]_CC over the brand-new locations --------------------------------
    for (int e = tid; e < k * k; e += blockDim.x) {
      const int r = e / k, c = e % k;
      if (c > r) continue;
      double v = (r == c) ? 1.0 : 0.0;                        // inactive slots: identity
      if (p2[r] >= 0 && p2[c] >= 0) {
        const int hi = p2[r] > p2[c] ? p2[r] : p2[c], lo = p2[r] > p2[c] ? p2[c] : p2[r];
        v = a.inv2[(int64_t)hi * a.ld2 + lo];
      } else if (r != c) {
        v = 0.0;
      }
      M[r * pitch + c] = v;
    }
    __syncthreads();
    const double t2 = mi_logabsdet(M, k, pitch, tid);
    // ---- term 3: sum log|Delta| + logdet|Delta^-1 + [A3^-1]_CC| -------------------------------
    for (int e = tid; e < k * k; e += blockDim.x) {
      const int r = e / k, c = e % k;
      if (c > r) continue;
      double v = (r == c) ? 1.0 : 0.0;
      if (loc[r] >= 0 && loc[c] >= 0) {
        const int hi = loc[r] > loc[c] ? loc[r] : loc[c], lo = loc[r] > loc[c] ? loc[c] : loc[r];
        v = a.inv3[(int64_t)hi * a.ld3 + lo] + ((r == c) ? 1.0 / dl[r] : 0.0);
      } else if (r != c) {
        v = 0.0;
      }
      M[r * pitch + c] = v;
    }
    __syncthreads();
    const double t3 = mi_logabsdet(M, k, pitch, tid);
    if (tid == 0) {
      double nnew = 0.0, sl = 0.0;
      for (int s = 0; s < k; ++s) {
        if (p2[s] >= 0) nnew += 1.0;
        if (loc[s] >= 0) sl += log(fabs(dl[s]));
      }
      a.out[cand * 3 + 0] = t2;
      a.out[cand * 3 + 1] = nnew;
      a.out[cand * 3 + 2] = sl + t3;
    }
  }
}

#define MI_MAXK 128
#define MI_MAXK_LARGE 2048

extern "C" int64_t algp_mi_terms_large_work_doubles(int k, int64_t B) {
  if (k <= MI_MAXK || B <= 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t grid = B < (int64_t)sms * 2 ? B : (int64_t)sms * 2;
  return grid * k * ((int64_t)k + 1);
}

extern "C" int algp_mi_terms_large(const double* inv2, int64_t ld2, const int32_t* pos2, const double* inv3, int64_t ld3,
                                   const int32_t* idx, int k, int64_t B, const uint8_t* skip, double delta_new,
                                   double delta_old, double* out3, double* work, int64_t work_doubles, void* stream) {
  if (!inv3 || !pos2 || !idx || !out3 || k < 1 || k > MI_MAXK_LARGE || B < 0 || delta_new == 0.0 || delta_old == 0.0)
    return ALGP_ERR_INVALID;
  if (B == 0) return ALGP_OK;
  const bool large = k > MI_MAXK;
  // the scratch may be smaller than the preferred size: fewer CTAs run (at least one candidate's k x (k+1) matrix)
  const int64_t per_cta = (int64_t)k * ((int64_t)k + 1);
  if (large && (!work || work_doubles < per_cta)) return ALGP_ERR_INVALID;
  MiArgs a;
  a.inv2 = inv2; a.ld2 = ld2; a.pos2 = pos2; a.inv3 = inv3; a.ld3 = ld3; a.idx = idx; a.k = k; a.B = B; a.skip = skip;
  a.delta_new = delta_new; a.delta_old = delta_old; a.out = out3; a.work = work;
  size_t smem = ((large ? 0 : (size_t)k * (k + 1)) + k) * 8 + (size_t)2 * k * 4 + 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (int)(B < (int64_t)sms * 2 ? B : (int64_t)sms * 2);
  if (large && work_doubles / per_cta < grid) grid = (int)(work_doubles / per_cta);
  if (large) {
    static AlgpPerDevice configured;
    if (configured.raise(smem)) {
      ALGP_CUDA(cudaFuncSetAttribute(mi_terms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    mi_terms_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  } else {
    static AlgpPerDevice configured;
    if (configured.raise(smem)) {
      ALGP_CUDA(cudaFuncSetAttribute(mi_terms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    mi_terms_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

extern "C" int algp_mi_terms(const double* inv2, int64_t ld2, const int32_t* pos2, const double* inv3, int64_t ld3,
                             const int32_t* idx, int k, int64_t B, const uint8_t* skip, double delta_new, double delta_old,
                             double* out3, void* stream) {
  if (k > MI_MAXK) return ALGP_ERR_INVALID;
  return algp_mi_terms_large(inv2, ld2, pos2, inv3, ld3, idx, k, B, skip, delta_new, delta_old, out3, nullptr, 0, stream);
}


// ---------------------------------------------------------------------------
// Growing-prefix posteriors (reference agent.py:497-518 re-solves the GP for every prefix of the
// collected samples).  The factor of a leading principal sub-matrix is the leading block of the full
// factor, and so is its inverse, so with V = K(X*,X) Linv^T of the FULL ordered set
//     var_p[m]  = k** - sum_{j<p} V[m][j]^2
//     mean_p[m] = sum_{j<p} V[m][j] (beta[j] - ybar_p gamma[j]) + ybar_p,  beta = Linv y, gamma = Linv 1
// for every prefix length p: one pass over V with running sums.  out[i][m] = {sum V beta, sum V gamma, sum V^2}
// over j < prefix[i]; one warp per test row, prefixes ascending.
// ---------------------------------------------------------------------------
__global__ void prefix_reduce_kernel(const double* __restrict__ V, int64_t ldv, int64_t rows, const double* __restrict__ beta,
                                     const double* __restrict__ gamma, const int32_t* __restrict__ prefix, int nprefix,
                                     double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= rows) return;
  const double* row = V + m * ldv;
  double t1 = 0.0, t2 = 0.0, t3 = 0.0;
  int lo = 0;
  for (int i = 0; i < nprefix; ++i) {
    const int hi = prefix[i];
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int j = lo + lane; j < hi; j += 32) {
      const double v = row[j];
      s1 = fma(v, beta[j], s1);
      s2 = fma(v, gamma[j], s2);
      s3 = fma(v, v, s3);
    }
    t1 += warp_sum(s1);
    t2 += warp_sum(s2);
    t3 += warp_sum(s3);
    if (lane == 0) {
      double* o = out + ((int64_t)i * rows + m) * 3;
      o[0] = t1; o[1] = t2; o[2] = t3;
    }
    lo = hi;
  }
}

extern "C" int algp_prefix_reduce(const double* V, int64_t ldv, int64_t rows, const double* beta, const double* gamma,
                                  const int32_t* prefix_dev, int nprefix, double* out, void* stream) {
  if (!V || !beta || !gamma || !prefix_dev || !out || rows < 0 || nprefix < 0) return ALGP_ERR_INVALID;
  if (rows == 0 || nprefix == 0) return ALGP_OK;
  prefix_reduce_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(V, ldv, rows, beta, gamma, prefix_dev, nprefix, out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
