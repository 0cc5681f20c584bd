// Exact marginal-likelihood gradient for GPR.fit (reference models.py:145-158: 200 Adam steps on
// -mll/N through gpytorch autograd).  With A = K + diag(var) + sigma_n^2 I = L L^T, alpha = A^-1 y0
// and G = alpha alpha^T - A^-1,
//     d(-ll)/d theta_p = -1/2 sum_ij G_ij dA_ij/d theta_p ,
// for theta = (log lengthscale[d], log outputscale, log noise).  A^-1 = Linv^T Linv comes from the
// DMMA GEMM core (lower tiles); this file fuses "rebuild K_ij and its derivatives, multiply by G_ij,
// reduce" into one pass over the lower triangle, so neither K nor dK/dtheta is ever stored.
#include "gemm.cuh"

#define MG_TILE 64
#define MG_NOUT (ALGP_MAX_D + 2)

struct MllArgs {
  KernelParams kp;
  const double* x;       // [n x d]
  int64_t n;
  const double* alpha;   // [n]
  const double* Ainv;    // [npad x lda], lower triangle valid
  int64_t lda;
  double* partial;       // [tiles x MG_NOUT]
};

// one CTA per lower 64x64 tile; thread (ty,tx) of a 16x16 grid owns the 4x4 patch rows ty+16a, cols tx+16b
__global__ void __launch_bounds__(256) mll_grad_kernel(const MllArgs a) {
  __shared__ double sxi[MG_TILE][ALGP_MAX_D], sxj[MG_TILE][ALGP_MAX_D], sai[MG_TILE], saj[MG_TILE];
  __shared__ double sred[8][MG_NOUT];
  const int L = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= L) ++ti;
  while (ti * (ti + 1) / 2 > L) --ti;
  const int tj = L - ti * (ti + 1) / 2;
  const int d = a.kp.d;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t i0 = (int64_t)ti * MG_TILE, j0 = (int64_t)tj * MG_TILE;
  for (int e = tid; e < MG_TILE * d; e += 256) {
    int r = e / d, q = e - r * d;
    sxi[r][q] = (i0 + r < a.n) ? a.x[(i0 + r) * d + q] * a.kp.inv_ls[q] : 0.0;
    sxj[r][q] = (j0 + r < a.n) ? a.x[(j0 + r) * d + q] * a.kp.inv_ls[q] : 0.0;
  }
  if (tid < MG_TILE) {
    sai[tid] = (i0 + tid < a.n) ? a.alpha[i0 + tid] : 0.0;
    saj[tid] = (j0 + tid < a.n) ? a.alpha[j0 + tid] : 0.0;
  }
  __syncthreads();

  double acc[MG_NOUT];
#pragma unroll
  for (int q = 0; q < MG_NOUT; ++q) acc[q] = 0.0;
  const double os = a.kp.outputscale;
  const double s3 = 1.7320508075688772;
#pragma unroll
  for (int aa = 0; aa < 4; ++aa)
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int r = ty + 16 * aa, c = tx + 16 * bb;
      const int64_t gi = i0 + r, gj = j0 + c;
      if (gi >= a.n || gj > gi) continue;                 // lower triangle of the valid part
      const double w = (gi == gj) ? 1.0 : 2.0;
      const double G = sai[r] * saj[c] - a.Ainv[gi * a.lda + gj];
      double dp[ALGP_MAX_D];
      double r2 = 0.0;
#pragma unroll
      for (int q = 0; q < ALGP_MAX_D; ++q) {
        dp[q] = 0.0;
        if (q < d) {
          const double df = sxi[r][q] - sxj[c][q];
          dp[q] = df * df;
          r2 += dp[q];
        }
      }
      double kij, dk;                                     // dK/dlog l_q = dk * dp[q]
      if (a.kp.kind == 0) {
        kij = os * exp_nonpos(-0.5 * r2);
        dk = kij;
      } else {
        const double rr = sqrt(r2);
        const double e = exp_nonpos(-s3 * rr);
        kij = os * fma(s3, rr, 1.0) * e;
        dk = 3.0 * os * e;
      }
      const double wg = w * G;
#pragma unroll
      for (int q = 0; q < ALGP_MAX_D; ++q) acc[q] = fma(wg * dk, dp[q], acc[q]);
      acc[ALGP_MAX_D] = fma(wg, kij, acc[ALGP_MAX_D]);
      if (gi == gj) acc[ALGP_MAX_D + 1] += G;
    }
  // fixed-order block reduction
#pragma unroll
  for (int q = 0; q < MG_NOUT; ++q) {
    double v = warp_sum(acc[q]);
    if ((tid & 31) == 0) sred[tid >> 5][q] = v;
  }
  __syncthreads();
  if (tid < MG_NOUT) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += sred[w8][tid];
    a.partial[(int64_t)L * MG_NOUT + tid] = v;
  }
}

// out[q] = 0.5 * sum_tiles partial[tile][q]  (q < d: lengthscales; d: outputscale; d+1: noise * sigma_n^2)
__global__ void mll_grad_reduce_kernel(const double* __restrict__ partial, int64_t tiles, int d, double noise,
                                       double* __restrict__ out) {
  __shared__ double s[256];
  for (int q = 0; q < d + 2; ++q) {
    const int src = (q < d) ? q : (q == d ? ALGP_MAX_D : ALGP_MAX_D + 1);
    double v = 0.0;
    for (int64_t t = threadIdx.x; t < tiles; t += 256) v += partial[t * MG_NOUT + src];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = 0.5 * s[0] * (q == d + 1 ? noise : 1.0);
    __syncthreads();
  }
}

int make_kernel_params(KernelParams* kp, int d, const double* log_ls_host, double log_os, int kind);

extern "C" int64_t algp_mll_grad_work_doubles(int64_t n) {
  int64_t t = (n + MG_TILE - 1) / MG_TILE;
  return t * (t + 1) / 2 * MG_NOUT;
}

// grad_out[d+2] (device) = 0.5 * tr(G dA/dtheta): multiply by -1/N for the gradient of the reference's loss
extern "C" int algp_mll_grad(const double* x, int64_t n, int d, const double* log_ls_host, double log_os, int kind,
                             double noise, const double* alpha, const double* Ainv, int64_t lda, double* work,
                             double* grad_out, void* stream) {
  if (!x || !alpha || !Ainv || !work || !grad_out || n <= 0 || lda < n) return ALGP_ERR_INVALID;
  MllArgs a;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.x = x; a.n = n; a.alpha = alpha; a.Ainv = Ainv; a.lda = lda; a.partial = work;
  const int64_t t = (n + MG_TILE - 1) / MG_TILE;
  const int64_t tiles = t * (t + 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  mll_grad_kernel<<<(unsigned)tiles, 256, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  mll_grad_reduce_kernel<<<1, 256, 0, st>>>(work, tiles, d, noise, grad_out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// Ainv (lower tiles) = Linv^T Linv: A^-1 from the inverse factor (potri)
extern "C" int algp_potri_lower(const double* Linv, int64_t npad, int64_t ldi, double* Ainv, int64_t lda, void* stream) {
  if (!Linv || !Ainv || npad % ALGP_BLK || ldi < npad || lda < npad || (ldi & 1) || (lda & 1)) return ALGP_ERR_INVALID;
  GemmArgs g = gemm_args_default();
  g.A = Linv; g.lda = ldi;            // mn-major: A[m][k] = Linv[k][m]
  g.B = Linv; g.ldb = ldi;            // mn-major: B[k][n] = Linv[k][n]
  g.C = Ainv; g.ldc = lda;
  g.MT = g.NT = (int)(npad / ALGP_BLK); g.K = (int)npad;
  g.tmap = TM_LOWER;
  g.kbeg_rule = KB_MT;                // Linv[k][m] = 0 for k < m, and m >= n on the lower tiles
  return gemm_f64_launch(g, LAY_MNMAJ, LAY_MNMAJ, 1, (cudaStream_t)stream);
}
