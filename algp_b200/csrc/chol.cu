// K2: blocked right-looking Cholesky A = L L^T (fp64), the explicit blocked
// inverse L^-1, and the O(N^2) solves built on it.
//
// Replaces np.linalg.inv(cov_aa) (reference utils.py:300, getrf+getri) and the
// per-call np.linalg.slogdet (utils.py:193).  Everything is row-major with
// N padded to a multiple of 128 (identity on the padded diagonal).
//
//   potrf : for each 128-wide panel j
//             potf2inv   one CTA: factor the diagonal block and invert it
//             panel      P <- P * inv(L_jj)^T            (DMMA GEMM, in place)
//             syrk       A22 <- A22 - P P^T, lower tiles (DMMA GEMM)
//   trtri : recursive doubling over block sizes b = 128, 256, ...:
//             T = L21 * Linv11 ;  Linv21 = -Linv22 * T    (two batched DMMA GEMMs / level)
//   solves: beta = L^-1 y, alpha = L^-T beta as triangular GEMVs over Linv (HBM bound),
//           log det A = 2 sum log L_ii.
#include "gemm.cuh"
#include <stdlib.h>

// ---------------------------------------------------------------------------
// GEMM launcher
// ---------------------------------------------------------------------------
template <int ALAY, int BLAY, int TM, int TN, bool SUBC, int NSTAGE = G_STAGES>
static int gemm_launch_t(GemmArgs a, int batch, cudaStream_t st) {
  static AlgpPerDevice configured;
  if (configured.raise(1)) {
    ALGP_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<ALAY, BLAY, TM, TN, SUBC, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   G_SMEM_BYTES(TM, TN, NSTAGE)));
  }
  a.MT *= 128 / TM;                                  // callers count 128-tiles
  a.NT *= 128 / TN;
  int64_t tiles = (a.tmap == TM_LOWER) ? ((TN * 2 == TM) ? (int64_t)a.MT * (a.MT + 1) : (int64_t)a.MT * (a.MT + 1) / 2)
                                       : (int64_t)a.MT * a.NT;
  if (tiles <= 0 || batch <= 0) return ALGP_OK;
  dim3 grid((unsigned)tiles, 1, (unsigned)batch);
  gemm_f64_kernel<ALAY, BLAY, TM, TN, SUBC, NSTAGE><<<grid, 256, G_SMEM_BYTES(TM, TN, NSTAGE), st>>>(a);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

template <int ALAY, int BLAY>
static int gemm_launch_l(const GemmArgs& a, int batch, cudaStream_t st) {
  // Latency shapes when the 128-tile grid cannot fill the SMs.  Row-norm partials are per 64-column
  // tile (both shapes used with them have TN = 64); an in-place launch (NT == 1) may only shrink its
  // row extent.
  int64_t tiles = (a.tmap == TM_LOWER) ? (int64_t)a.MT * (a.MT + 1) / 2 : (int64_t)a.MT * a.NT;
  const bool small = tiles * batch < 120;
  const bool subc = a.alpha == -1.0 && a.beta == 1.0 && a.store_c && !a.rn_partial && !a.inplace_rows && a.C;
  // Default for large launches: 128x64 tiles, 3 stages, so that two CTAs share an SM and overlap each
  // other's pipeline fill and epilogue (measured 8% faster than 128x128x4 on the N=16384 factorisation;
  // ALGP_GEMM_SHAPE=0 selects the one-CTA 128x128 shape for A/B runs).
  static int shape = -1;
  if (shape < 0) {
    const char* e = getenv("ALGP_GEMM_SHAPE");
    shape = e ? atoi(e) : 1;
  }
  if (a.rn_partial) {                                   // partial columns are 64 wide: TN = 64 shapes only
    if (small) return gemm_launch_t<ALAY, BLAY, 64, 64, false>(a, batch, st);
    return gemm_launch_t<ALAY, BLAY, 128, 64, false, 3>(a, batch, st);
  }
  if (shape == 1 && !small) {
    if (subc) return gemm_launch_t<ALAY, BLAY, 128, 64, true, 3>(a, batch, st);
    return gemm_launch_t<ALAY, BLAY, 128, 64, false, 3>(a, batch, st);
  }
  if (subc) {
    if (small) return gemm_launch_t<ALAY, BLAY, 64, 64, true>(a, batch, st);
    return gemm_launch_t<ALAY, BLAY, 128, 128, true>(a, batch, st);
  }
  if (small && a.inplace_rows) return gemm_launch_t<ALAY, BLAY, 64, 128, false>(a, batch, st);
  if (small) return gemm_launch_t<ALAY, BLAY, 64, 64, false>(a, batch, st);
  return gemm_launch_t<ALAY, BLAY, 128, 128, false>(a, batch, st);
}

int gemm_f64_launch(const GemmArgs& a, int alay, int blay, int batch, cudaStream_t st) {
  if (a.K % GK) return ALGP_ERR_INVALID;
  if (alay == LAY_KMAJ && blay == LAY_KMAJ) return gemm_launch_l<LAY_KMAJ, LAY_KMAJ>(a, batch, st);
  if (alay == LAY_KMAJ && blay == LAY_MNMAJ) return gemm_launch_l<LAY_KMAJ, LAY_MNMAJ>(a, batch, st);
  if (alay == LAY_MNMAJ && blay == LAY_MNMAJ) return gemm_launch_l<LAY_MNMAJ, LAY_MNMAJ>(a, batch, st);
  return gemm_launch_l<LAY_MNMAJ, LAY_KMAJ>(a, batch, st);
}

// ---------------------------------------------------------------------------
// potf2inv: factor one 128x128 diagonal block and invert the factor, one CTA.
//
// 256 threads in a 16x16 grid; thread (ty,tx) owns the cyclic patch rows ty+16i,
// cols tx+16j in registers, so every thread stays busy as the active window
// shrinks.  The factorisation is an un-normalised elimination (LDL^T style: no
// sqrt on the column-to-column critical path, one __syncthreads per column
// through double-buffered broadcasts); L = Ltilde D^1/2 is formed at the end.
// The same multipliers are applied, in the same loop, to the identity, which
// yields Y = Ltilde^-1 and hence L^-1 = D^-1/2 Y.
// ---------------------------------------------------------------------------
#define P2_SMEM_BYTES ((2 * 128 + 2 * 128 + 2 * 128) * 8)

// Columns 16*IC .. 16*IC+15.  One barrier per column: before it the owners publish column c of
// the Schur complement (for the elimination) and row c of Y (for the inverse); after it every
// thread updates its live patches of both.  Only lower-triangular patches (j <= i) are live, so
// the two register patches cost 2 x 36 doubles.  IC is a template parameter: patch indices are
// static and only the boundary block needs a mask.
template <int IC>
__device__ __forceinline__ void p2_block(double (&acc)[8][8], double (&yac)[8][8], double* colbuf, double* rowbuf,
                                         double* pivbuf, int ty, int tx, int tid, int j0, int* info) {
#pragma unroll 1
  for (int cc = 0; cc < 16; ++cc) {
    const int c = IC * 16 + cc;
    double* cb = colbuf + (c & 1) * 128;
    double* rb = rowbuf + (c & 1) * 128;
    if (tx == cc) {
#pragma unroll
      for (int i = IC; i < 8; ++i) cb[ty + 16 * i] = acc[i][IC];
    }
    if (ty == cc) {
#pragma unroll
      for (int j = 0; j <= IC; ++j) rb[tx + 16 * j] = yac[IC][j];
    }
    __syncthreads();
    const double piv = cb[c];
    const bool ok = piv > 0.0;
    const double rcp = ok ? 1.0 / piv : 0.0;
    if (tid == 0) {
      pivbuf[c] = piv;
      if (!ok) atomicCAS(info, 0, j0 + c + 1);
    }
    double ri[8], cj[8], yj[8];
#pragma unroll
    for (int i = IC; i < 8; ++i) ri[i] = cb[ty + 16 * i] * rcp;      // multipliers l~_{r,c}
    if (ty <= cc) ri[IC] = 0.0;                                      // rows at or above the pivot
#pragma unroll
    for (int j = IC; j < 8; ++j) cj[j] = cb[tx + 16 * j];
    if (tx <= cc) cj[IC] = 0.0;                                      // columns at or left of the pivot
#pragma unroll
    for (int j = 0; j <= IC; ++j) yj[j] = rb[tx + 16 * j];           // Y[c][.] (zero right of c by construction)
#pragma unroll
    for (int i = IC; i < 8; ++i) {
#pragma unroll
      for (int j = IC; j <= i; ++j) acc[i][j] = fma(-ri[i], cj[j], acc[i][j]);
#pragma unroll
      for (int j = 0; j <= IC; ++j) yac[i][j] = fma(-ri[i], yj[j], yac[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256, 1) potf2inv_kernel(double* __restrict__ A, int64_t ld, double* __restrict__ Linv,
                                                          int64_t ldi, int j0, int* __restrict__ info) {
  extern __shared__ __align__(16) double p2_smem[];
  double* colbuf = p2_smem;                   // [2][128]
  double* rowbuf = colbuf + 256;              // [2][128]
  double* pivbuf = rowbuf + 256;              // [128]
  double* rsbuf = pivbuf + 128;               // [128] 1/sqrt(piv)

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double acc[8][8], yac[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      acc[i][j] = A[(int64_t)(ty + 16 * i) * ld + tx + 16 * j];
      yac[i][j] = (i == j && ty == tx) ? 1.0 : 0.0;
    }

  p2_block<0>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<1>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<2>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<3>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<4>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<5>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<6>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  p2_block<7>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  __syncthreads();
  if (tid < 128) {
    const double piv = pivbuf[tid];
    rsbuf[tid] = piv > 0.0 ? 1.0 / sqrt(piv) : 0.0;
  }
  __syncthreads();

  // L = Ltilde D^1/2 (acc holds the un-normalised columns), Linv = D^-1/2 Y; zeros above the diagonal
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      double l = 0.0, li = 0.0;
      if (j <= i) {
        if (r > c) {
          l = acc[i][j] * rsbuf[c];
          li = yac[i][j] * rsbuf[r];
        } else if (r == c) {
          l = pivbuf[c] * rsbuf[c];
          li = rsbuf[r];
        }
      }
      A[(int64_t)r * ld + c] = l;
      Linv[(int64_t)r * ldi + c] = li;
    }
}

// ---------------------------------------------------------------------------
// Rank-R variant: R columns per barrier.  The owners publish R raw columns of the Schur complement and R raw
// rows of Y; every thread then eliminates the R x R pivot block REDUNDANTLY in registers (R dependent
// reciprocals instead of R barrier + shared-memory round trips), brings its own rows / columns of the R
// published vectors up to date with the same multipliers, and applies one rank-R update to its patches.
// The column-to-column chain (publish -> barrier -> read -> reciprocal -> update) is paid once per R columns.
// ---------------------------------------------------------------------------
#define PR_SMEM_DOUBLES(R) (2 * (R) * 128 + 2 * (R) * 128 + 128 + 128)

template <int IC, int R>
__device__ __forceinline__ void pr_block(double (&acc)[8][8], double (&yac)[8][8], double* colbuf, double* rowbuf,
                                         double* pivbuf, int ty, int tx, int tid, int j0, int* info) {
#pragma unroll 1
  for (int cc = 0; cc < 16; cc += R) {
    const int c = IC * 16 + cc;
    double* cb = colbuf + ((c / R) & 1) * (R * 128);
    double* rb = rowbuf + ((c / R) & 1) * (R * 128);
    {
      const int q = tx - cc;
      if (q >= 0 && q < R) {
#pragma unroll
        for (int i = IC; i < 8; ++i) cb[q * 128 + ty + 16 * i] = acc[i][IC];
      }
      const int qy = ty - cc;
      if (qy >= 0 && qy < R) {
#pragma unroll
        for (int j = 0; j <= IC; ++j) rb[qy * 128 + tx + 16 * j] = yac[IC][j];
      }
    }
    __syncthreads();
    // ---- R x R pivot block, eliminated redundantly by every thread ----
    double P[R][R], M[R][R], rcp[R];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) P[a][b] = cb[b * 128 + c + a];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const double piv = P[b][b];
      const bool ok = piv > 0.0;
      rcp[b] = ok ? 1.0 / piv : 0.0;
      if (tid == 0) {
        pivbuf[c + b] = piv;
        if (!ok) atomicCAS(info, 0, j0 + c + b + 1);
      }
#pragma unroll
      for (int a = b + 1; a < R; ++a) M[a][b] = P[a][b] * rcp[b];
#pragma unroll
      for (int a = b + 1; a < R; ++a)
#pragma unroll
        for (int b2 = b + 1; b2 <= a; ++b2) P[a][b2] = fma(-M[a][b], P[b2][b], P[a][b2]);
    }
    // ---- multipliers of this thread's rows for the R columns (normalised, masked at or above each pivot) ----
    double ri[R][8];
#pragma unroll
    for (int i = IC; i < 8; ++i) {
#pragma unroll
      for (int b = 0; b < R; ++b) {
        double v = cb[b * 128 + ty + 16 * i];
#pragma unroll
        for (int b1 = 0; b1 < b; ++b1) v = fma(-ri[b1][i], P[b][b1], v);
        ri[b][i] = v * rcp[b];
        if (i == IC && ty <= cc + b) ri[b][i] = 0.0;
      }
    }
    // ---- per column patch: the R updated (un-normalised) column values, then the rank-R update ----
#pragma unroll
    for (int j = IC; j < 8; ++j) {
      double cj[R];
#pragma unroll
      for (int b = 0; b < R; ++b) {
        double v = cb[b * 128 + tx + 16 * j];
#pragma unroll
        for (int b1 = 0; b1 < b; ++b1) v = fma(-cj[b1], M[b][b1], v);
        cj[b] = v;
        if (j == IC && tx <= cc + b) cj[b] = 0.0;
      }
#pragma unroll
      for (int i = j; i < 8; ++i) {
        double t = acc[i][j];
#pragma unroll
        for (int b = 0; b < R; ++b) t = fma(-ri[b][i], cj[b], t);
        acc[i][j] = t;
      }
    }
#pragma unroll
    for (int j = 0; j <= IC; ++j) {
      double yj[R];
#pragma unroll
      for (int b = 0; b < R; ++b) {
        double v = rb[b * 128 + tx + 16 * j];                       // Y[c+b][.], zero right of its diagonal
#pragma unroll
        for (int b1 = 0; b1 < b; ++b1) v = fma(-M[b][b1], yj[b1], v);
        yj[b] = v;
      }
#pragma unroll
      for (int i = IC; i < 8; ++i) {
        double t = yac[i][j];
#pragma unroll
        for (int b = 0; b < R; ++b) t = fma(-ri[b][i], yj[b], t);
        yac[i][j] = t;
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(256, 1) potf2inv_rank_kernel(double* __restrict__ A, int64_t ld, double* __restrict__ Linv,
                                                               int64_t ldi, int j0, int* __restrict__ info) {
  extern __shared__ __align__(16) double pr_smem[];
  double* colbuf = pr_smem;                   // [2][R][128]
  double* rowbuf = colbuf + 2 * R * 128;      // [2][R][128]
  double* pivbuf = rowbuf + 2 * R * 128;      // [128]
  double* rsbuf = pivbuf + 128;               // [128] 1/sqrt(piv)

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double acc[8][8], yac[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      acc[i][j] = A[(int64_t)(ty + 16 * i) * ld + tx + 16 * j];
      yac[i][j] = (i == j && ty == tx) ? 1.0 : 0.0;
    }

  pr_block<0, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<1, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<2, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<3, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<4, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<5, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<6, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  pr_block<7, R>(acc, yac, colbuf, rowbuf, pivbuf, ty, tx, tid, j0, info);
  __syncthreads();
  if (tid < 128) {
    const double piv = pivbuf[tid];
    rsbuf[tid] = piv > 0.0 ? 1.0 / sqrt(piv) : 0.0;
  }
  __syncthreads();

  // L = Ltilde D^1/2 (acc holds the un-normalised columns), Linv = D^-1/2 Y; zeros above the diagonal
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      double l = 0.0, li = 0.0;
      if (j <= i) {
        if (r > c) {
          l = acc[i][j] * rsbuf[c];
          li = yac[i][j] * rsbuf[r];
        } else if (r == c) {
          l = pivbuf[c] * rsbuf[c];
          li = rsbuf[r];
        }
      }
      A[(int64_t)r * ld + c] = l;
      Linv[(int64_t)r * ldi + c] = li;
    }
}

// 1 = one column per barrier (potf2inv_kernel), 2 / 4 = rank-R steps.  Measured on B200 (scripts/prof_potf2.py):
// 46.4 / 41.1 / 43.5 us per 128-block: the kernel is bound by instruction issue at two warps per scheduler
// (~115 issued instructions per column per warp, a third of them DFMA), not by the barrier chain.  A DMMA-fragment
// variant (8x8 tiles dealt to the warps, one DMMA per tile per 4-column step) was correct but 2x slower: per-tile
// addressing and activity tests cost more issue slots than the eight DFMAs a DMMA replaces.
// Round 2: a BLOCKED variant (16-column panels eliminated inside the warps by shuffles, rank-16 register updates, two
// barriers per panel; scripts/micro/potf2_blocked.cuh) is correct and no faster, 43.5 us: shared memory delivers
// 128 B per clock to the register files however few distinct addresses a load has, so both the pivot-row broadcasts
// of the in-warp elimination and the operand loads of the rank-16 update (16 loads per 36 DFMAs per thread) sit on
// the LSU; with one pivot warp and trailing consumer warps the per-column publish / release costs as much as the
// barrier it replaced (47.6 us).  Measurements and the phase breakdown: profiles/r02_potf2_blocked.log.
static int g_potf2_rank = 2;
extern "C" int algp_set_potf2_rank(int r) {
  if (r != 1 && r != 2 && r != 4) return ALGP_ERR_INVALID;
  g_potf2_rank = r;
  return ALGP_OK;
}

static int potf2inv_launch(double* A, int64_t ld, double* Linv, int64_t ldi, int j0, int* info, cudaStream_t st) {
  static AlgpPerDevice configured;
  if (configured.raise(1)) {
    ALGP_CUDA(cudaFuncSetAttribute(potf2inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
  }
  if (g_potf2_rank == 4) {
    potf2inv_rank_kernel<4><<<1, 256, PR_SMEM_DOUBLES(4) * 8, st>>>(A, ld, Linv, ldi, j0, info);
  } else if (g_potf2_rank == 2) {
    potf2inv_rank_kernel<2><<<1, 256, PR_SMEM_DOUBLES(2) * 8, st>>>(A, ld, Linv, ldi, j0, info);
  } else {
    potf2inv_kernel<<<1, 256, P2_SMEM_BYTES, st>>>(A, ld, Linv, ldi, j0, info);
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// potrf (right-looking, with look-ahead)
//
// The diagonal-block kernel runs on ONE SM; serialised with the trailing update it
// would idle the other 147.  So each step first updates only block column j+1,
// hands "factor block j+1 + solve panel j+1" to an auxiliary stream, and runs the
// rest of step j's trailing update (block columns >= j+2, disjoint from the panel
// being solved) on the caller's stream meanwhile.
// ---------------------------------------------------------------------------
struct PotrfAux {
  cudaStream_t aux = nullptr;
  cudaEvent_t col_ready = nullptr, panel_ready = nullptr;
};
static PotrfAux g_potrf_aux[64];

static int potrf_aux_get(PotrfAux** out) {
  int dev = 0;
  ALGP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return ALGP_ERR_UNSUPPORTED;
  PotrfAux& a = g_potrf_aux[dev];
  if (!a.aux) {
    // highest priority: the diagonal-block kernel and the panel solve are the critical path and must win
    // SM slots against the thousands of queued CTAs of the concurrent trailing update
    int prio_lo = 0, prio_hi = 0;
    ALGP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    ALGP_CUDA(cudaStreamCreateWithPriority(&a.aux, cudaStreamNonBlocking, prio_hi));
    ALGP_CUDA(cudaEventCreateWithFlags(&a.col_ready, cudaEventDisableTiming));
    ALGP_CUDA(cudaEventCreateWithFlags(&a.panel_ready, cudaEventDisableTiming));
  }
  *out = &a;
  return ALGP_OK;
}

static int potrf_panel(double* A, int64_t ld, double* Linv, int64_t ldi, int j, int nb, cudaStream_t st) {
  // P <- P * inv(L_jj)^T for the row blocks below the diagonal (in place: each CTA reads only the rows it writes)
  const int rem = nb - 1 - j;
  if (rem <= 0) return ALGP_OK;
  double* panel = A + (int64_t)(j + 1) * ALGP_BLK * ld + (int64_t)j * ALGP_BLK;
  GemmArgs g = gemm_args_default();
  g.A = panel; g.lda = ld;
  g.B = Linv + (int64_t)j * ALGP_BLK * (ldi + 1); g.ldb = ldi;
  g.C = panel; g.ldc = ld;
  g.MT = rem; g.NT = 1; g.K = ALGP_BLK;
  g.inplace_rows = 1;
  return gemm_f64_launch(g, LAY_KMAJ, LAY_KMAJ, 1, st);
}

// Factor the npad x npad block at A (diagonal 128-blocks of Linv receive inv(L_jj)).  *info_dev is NOT reset:
// a failing pivot records col0 + column + 1 unless an earlier failure is already recorded.
int potrf_block(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int* info_dev, int col0, cudaStream_t st) {
  PotrfAux* ax = nullptr;
  int rc = potrf_aux_get(&ax);
  if (rc) return rc;
  const int nb = (int)(npad / ALGP_BLK);
  if (nb == 0) return ALGP_OK;
  rc = potf2inv_launch(A, ld, Linv, ldi, col0, info_dev, st);
  if (rc) return rc;
  rc = potrf_panel(A, ld, Linv, ldi, 0, nb, st);
  if (rc) return rc;
  for (int j = 0; j + 1 < nb; ++j) {
    const int rem = nb - 1 - j;
    double* panel = A + (int64_t)(j + 1) * ALGP_BLK * ld + (int64_t)j * ALGP_BLK;
    {  // block column j+1 of the trailing update: A[i, j+1] -= P_i P_{j+1}^T, i >= j+1
      GemmArgs g = gemm_args_default();
      g.A = panel; g.lda = ld;
      g.B = panel; g.ldb = ld;
      g.C = A + (int64_t)(j + 1) * ALGP_BLK * (ld + 1); g.ldc = ld;
      g.MT = rem; g.NT = 1; g.K = ALGP_BLK;
      g.alpha = -1.0; g.beta = 1.0;
      rc = gemm_f64_launch(g, LAY_KMAJ, LAY_KMAJ, 1, st);
      if (rc) return rc;
    }
    ALGP_CUDA(cudaEventRecord(ax->col_ready, st));
    ALGP_CUDA(cudaStreamWaitEvent(ax->aux, ax->col_ready, 0));
    rc = potf2inv_launch(A + (int64_t)(j + 1) * ALGP_BLK * (ld + 1), ld, Linv + (int64_t)(j + 1) * ALGP_BLK * (ldi + 1), ldi,
                         col0 + (j + 1) * ALGP_BLK, info_dev, ax->aux);
    if (rc) return rc;
    rc = potrf_panel(A, ld, Linv, ldi, j + 1, nb, ax->aux);
    if (rc) return rc;
    ALGP_CUDA(cudaEventRecord(ax->panel_ready, ax->aux));
    if (rem > 1) {  // the rest of step j: lower tiles of the block columns >= j+2
      GemmArgs g = gemm_args_default();
      g.A = panel + (int64_t)ALGP_BLK * ld; g.lda = ld;
      g.B = g.A; g.ldb = ld;
      g.C = A + (int64_t)(j + 2) * ALGP_BLK * (ld + 1); g.ldc = ld;
      g.MT = rem - 1; g.NT = rem - 1; g.K = ALGP_BLK;
      g.tmap = TM_LOWER; g.alpha = -1.0; g.beta = 1.0;
      rc = gemm_f64_launch(g, LAY_KMAJ, LAY_KMAJ, 1, st);
      if (rc) return rc;
    }
    ALGP_CUDA(cudaStreamWaitEvent(st, ax->panel_ready, 0));
  }
  return ALGP_OK;
}

extern "C" int algp_potrf(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int* info_dev, void* stream) {
  if (!A || !Linv || !info_dev || npad < 0 || npad % ALGP_BLK || ld < npad || ldi < npad || (ld & 1) || (ldi & 1))
    return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  ALGP_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  return potrf_block(A, npad, ld, Linv, ldi, info_dev, 0, st);
}

// ---------------------------------------------------------------------------
// trtri: complete Linv (diagonal 128-blocks already hold inv(L_jj)) by
// recursive doubling.  work: >= npad*npad/4 doubles.
// ---------------------------------------------------------------------------
__global__ void zero_upper_blocks_kernel(double* __restrict__ M, int64_t ld, int nb) {
  // zero the strictly-upper 128x128 blocks so Linv is a clean lower-triangular matrix
  int bi = blockIdx.y, bj = blockIdx.x;
  if (bj <= bi) return;
  double* blk = M + (int64_t)bi * ALGP_BLK * ld + (int64_t)bj * ALGP_BLK;
  for (int e = threadIdx.x; e < ALGP_BLK * ALGP_BLK / 2; e += blockDim.x) {
    int r = e / (ALGP_BLK / 2), c2 = (e % (ALGP_BLK / 2)) * 2;
    *reinterpret_cast<double2*>(blk + (int64_t)r * ld + c2) = make_double2(0.0, 0.0);
  }
}

extern "C" int64_t algp_trtri_work_doubles(int64_t npad) {
  int64_t need = 2;
  for (int64_t b = ALGP_BLK; b < npad; b *= 2) {
    int64_t pairs = (npad - b + 2 * b - 1) / (2 * b);
    if (pairs * b * b > need) need = pairs * b * b;
  }
  return need;
}

extern "C" int algp_trtri(const double* L, int64_t npad, int64_t ld, double* Linv, int64_t ldi, double* work,
                          int zero_upper, void* stream) {
  if (!L || !Linv || npad < 0 || npad % ALGP_BLK || ld < npad || ldi < npad || (ld & 1) || (ldi & 1)) return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)(npad / ALGP_BLK);
  if (nb > 1 && !work) return ALGP_ERR_INVALID;
  if (zero_upper && nb > 1) {
    zero_upper_blocks_kernel<<<dim3(nb, nb), 256, 0, st>>>(Linv, ldi, nb);
    ALGP_LAUNCH_CHECK();
  }
  for (int64_t b = ALGP_BLK; b < npad; b *= 2) {
    const int bt = (int)(b / ALGP_BLK);
    const int pairs = (int)((npad - b + 2 * b - 1) / (2 * b));   // pairs whose second block is non-empty
    if (pairs <= 0) break;
    // T_p = L21_p * Linv11_p        (Linv11 lower: k >= n)
    GemmArgs g = gemm_args_default();
    g.A = L + b * ld; g.lda = ld; g.a_bs = 2 * b * (ld + 1);
    g.B = Linv; g.ldb = ldi; g.b_bs = 2 * b * (ldi + 1);
    g.C = work; g.ldc = b; g.c_bs = b * b;
    g.MT = bt; g.NT = bt; g.K = (int)b;
    g.kbeg_rule = KB_NT;
    g.rows_left0 = npad - b; g.rows_left_bs = 2 * b;
    int rc = gemm_f64_launch(g, LAY_KMAJ, LAY_MNMAJ, pairs, st);
    if (rc) return rc;
    // Linv21_p = -Linv22_p * T_p    (Linv22 lower: k <= m)
    GemmArgs h = gemm_args_default();
    h.A = Linv + b * (ldi + 1); h.lda = ldi; h.a_bs = 2 * b * (ldi + 1);
    h.B = work; h.ldb = b; h.b_bs = b * b;
    h.C = Linv + b * ldi; h.ldc = ldi; h.c_bs = 2 * b * (ldi + 1);
    h.MT = bt; h.NT = bt; h.K = (int)b;
    h.kend_rule = KE_MT1;
    h.alpha = -1.0;
    h.rows_left0 = npad - b; h.rows_left_bs = 2 * b;
    rc = gemm_f64_launch(h, LAY_KMAJ, LAY_MNMAJ, pairs, st);
    if (rc) return rc;
  }
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// triangular GEMVs over a lower-triangular row-major matrix, log-det
// ---------------------------------------------------------------------------
// out[r] = sum_{k<=r} M[r][k] v[k] : one warp per row, coalesced along k
__global__ void gemv_lower_n_kernel(const double* __restrict__ M, int64_t n, int64_t ld, const double* __restrict__ v,
                                    double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const double* row = M + r * ld;
  double s0 = 0.0, s1 = 0.0;
  const int64_t kend = r + 1;
  int64_t k = 2 * lane;
  for (; k + 1 < kend; k += 64) {
    double2 m = *reinterpret_cast<const double2*>(row + k);
    s0 = fma(m.x, v[k], s0);
    s1 = fma(m.y, v[k + 1], s1);
  }
  if (k < kend) s0 = fma(row[k], v[k], s0);
  double s = warp_sum(s0 + s1);
  if (lane == 0) out[r] = s;
}

// partial[chunk][c] = sum_{r in chunk, r>=c} M[r][c] v[r] : thread per column, rows streamed coalesced
#define GT_CHUNK 256
__global__ void gemv_lower_t_kernel(const double* __restrict__ M, int64_t n, int64_t ld, const double* __restrict__ v,
                                    double* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * GT_CHUNK;
  if (c >= n) return;
  int64_t rb = r0 > c ? r0 : c;
  int64_t re = r0 + GT_CHUNK < n ? r0 + GT_CHUNK : n;
  double s0 = 0.0, s1 = 0.0;
  int64_t r = rb;
  for (; r + 1 < re; r += 2) {
    s0 = fma(M[r * ld + c], v[r], s0);
    s1 = fma(M[(r + 1) * ld + c], v[r + 1], s1);
  }
  if (r < re) s0 = fma(M[r * ld + c], v[r], s0);
  partial[(int64_t)blockIdx.y * n + c] = s0 + s1;
}

__global__ void colsum_kernel(const double* __restrict__ partial, int64_t n, int chunks, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += partial[(int64_t)k * n + c];
  out[c] = s;
}

extern "C" int64_t algp_gemv_work_doubles(int64_t n) { return ((n + GT_CHUNK - 1) / GT_CHUNK) * n; }

// beta = Linv y
extern "C" int algp_gemv_lower(const double* M, int64_t n, int64_t ld, const double* v, double* out, void* stream) {
  if (!M || !v || !out || n < 0 || ld < n || (ld & 1)) return ALGP_ERR_INVALID;
  if (n == 0) return ALGP_OK;
  gemv_lower_n_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(M, n, ld, v, out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// alpha = Linv^T beta ; work >= algp_gemv_work_doubles(n)
extern "C" int algp_gemv_lower_t(const double* M, int64_t n, int64_t ld, const double* v, double* out, double* work,
                                 void* stream) {
  if (!M || !v || !out || !work || n < 0 || ld < n) return ALGP_ERR_INVALID;
  if (n == 0) return ALGP_OK;
  const int chunks = (int)((n + GT_CHUNK - 1) / GT_CHUNK);
  cudaStream_t st = (cudaStream_t)stream;
  gemv_lower_t_kernel<<<dim3((unsigned)((n + 127) / 128), chunks), 128, 0, st>>>(M, n, ld, v, work);
  ALGP_LAUNCH_CHECK();
  colsum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(work, n, chunks, out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// out[0] = 2 * sum_i log L_ii ; out[1] = sum_i v_i^2 (optional v): fixed-order single-CTA reduction
__global__ void logdet_sumsq_kernel(const double* __restrict__ L, int64_t n, int64_t ld, const double* __restrict__ v,
                                    double* __restrict__ out) {
  __shared__ double s_ld[32], s_sq[32];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    a += log(L[i * (ld + 1)]);
    if (v) b = fma(v[i], v[i], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s_ld[threadIdx.x >> 5] = a; s_sq[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += s_ld[w]; tb += s_sq[w]; }
    out[0] = 2.0 * ta;
    out[1] = tb;
  }
}

extern "C" int algp_logdet_sumsq(const double* L, int64_t n, int64_t ld, const double* v, double* out2, void* stream) {
  if (!L || !out2 || n < 0) return ALGP_ERR_INVALID;
  logdet_sumsq_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L, n, ld, v, out2);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// C-ABI views of the GEMM core used outside this file
// ---------------------------------------------------------------------------
// V = Ks * Linv^T restricted to the lower-triangular structure of Linv
// (C[m][j] = sum_{k <= j} Ks[m][k] Linv[j][k], k-range cut per column tile); optional store of V and
// optional fused row partials  rn_partial[m][jt] = sum_{j in 64-column tile jt} V[m][j]^2.
extern "C" int algp_trmm_rt(const double* Ks, int64_t mpad, int64_t ldk, const double* Linv, int64_t npad, int64_t ldi,
                            double* V, int64_t ldv, double* rn_partial, void* stream) {
  if (!Ks || !Linv || mpad % ALGP_BLK || npad % ALGP_BLK || ldk < npad || ldi < npad || (ldk & 1) || (ldi & 1)) return ALGP_ERR_INVALID;
  if (!V && !rn_partial) return ALGP_ERR_INVALID;
  if (V && (ldv < npad || (ldv & 1))) return ALGP_ERR_INVALID;
  GemmArgs g = gemm_args_default();
  g.A = Ks; g.lda = ldk;
  g.B = Linv; g.ldb = ldi;
  g.C = V; g.ldc = ldv; g.store_c = V ? 1 : 0;
  g.MT = (int)(mpad / ALGP_BLK); g.NT = (int)(npad / ALGP_BLK); g.K = (int)npad;
  g.kend_rule = KE_NT1;
  g.rn_partial = rn_partial; g.rn_nt = 2 * g.NT;       // one partial per 64-column tile
  return gemm_f64_launch(g, LAY_KMAJ, LAY_KMAJ, 1, (cudaStream_t)stream);
}

// C = beta*C + alpha * A * B^T, A [mpad x k], B [npad x k] row-major; lower_only: MT == NT, lower tiles only
extern "C" int algp_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                            int64_t mpad, int64_t npad, int64_t kpad, double alpha, double beta, int lower_only, void* stream) {
  if (!A || !B || !C || mpad % ALGP_BLK || npad % ALGP_BLK || kpad % GK || (lda & 1) || (ldb & 1) || (ldc & 1)) return ALGP_ERR_INVALID;
  if (lower_only && mpad != npad) return ALGP_ERR_INVALID;
  GemmArgs g = gemm_args_default();
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.MT = (int)(mpad / ALGP_BLK); g.NT = (int)(npad / ALGP_BLK); g.K = (int)kpad;
  g.alpha = alpha; g.beta = beta;
  g.tmap = lower_only ? TM_LOWER : TM_FULL;
  return gemm_f64_launch(g, LAY_KMAJ, LAY_KMAJ, 1, (cudaStream_t)stream);
}
