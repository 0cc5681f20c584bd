// FP64 tensor-core GEMM core: C = alpha * A * B (+ beta * C) on 128x128 tiles.
//
// One kernel serves every O(N^3) / O(N^2 M) step of the path: the Cholesky
// panel solve and trailing SYRK, the blocked triangular inverse, the variance
// TRMM V = K(X*,X) L^-T with the row-norm fused into the epilogue, and the
// W = L^-1 Sigma_{B,:} build for candidate scoring.  Those replace the
// reference's np.linalg.inv + two np.dot (utils.py:300-305).
//
// * 256 threads = 8 warps in a 2 x 4 grid; a warp owns a 64 x 32 sub-tile as
//   8 x 4 DMMA.8x8x4 accumulator blocks (64 doubles / lane).
// * Operands stream through a 4-stage cp.async ring, 16 k-values per stage.
//   Either operand may be "k-major" (element (row,k) at row*ld + k: A row-major,
//   or B given as rows of B^T) or "mn-major" (element at k*ld + row).  Shared
//   tiles are padded (pitch 20 / 132 doubles) so the 8-byte fragment reads of a
//   half-warp hit 32 distinct banks.
// * All dimensions are multiples of 128 (the host pads with identity), so there
//   is no edge predication; triangular structure is expressed as per-tile
//   k-ranges and a lower-triangle tile map.
// * Callers describe the problem in 128-tiles; the launcher switches to the 64x64 CTA
//   tile (4x the CTAs, a quarter of the work each) when the grid would not fill the GPU.
#pragma once
#include "common.cuh"

#define GT 128
#define GK 16
#define G_STAGES 4
#define G_PITCH_K 20
#define G_PITCH_MN 132
#define G_SMEM_BYTES(TM, TN, ST) ((ST) * ((TM) + (TN)) * G_PITCH_K * 8)   // 163840 for 128x128x4, 92160 for 128x64x3

enum { LAY_KMAJ = 0, LAY_MNMAJ = 1 };
enum { TM_FULL = 0, TM_LOWER = 1 };
enum { KB_ZERO = 0, KB_MT = 1, KB_NT = 2 };             // k begins at 0 | mt*128 | nt*128
enum { KE_FULL = 0, KE_MT1 = 1, KE_NT1 = 2 };           // k ends at K | (mt+1)*128 | (nt+1)*128

struct GemmArgs {
  const double* A; int64_t lda; int64_t a_bs;
  const double* B; int64_t ldb; int64_t b_bs;
  double* C; int64_t ldc; int64_t c_bs;
  int MT, NT, K;                      // tile counts in units of the CTA tile edge chosen by the launcher
  int tmap, kbeg_rule, kend_rule;
  double alpha, beta;
  int64_t rows_left0, rows_left_bs;   // batch z has rows_left0 - z*rows_left_bs valid rows (ragged last batch)
  double* rn_partial; int rn_nt;      // optional: rn_partial[row*rn_nt + nt] = sum over the tile's 128 cols of value^2
  int store_c;
  int inplace_rows;                   // C aliases A: a CTA must cover the full row extent it reads
};

// TM x TN = CTA tile.  128x128 (8 warps x 64x32) is the throughput shape; 64x64 is the latency
// shape for launches with too few 128-tiles to fill the 148 SMs; 64x128 is the latency shape for
// the in-place panel solve, where a CTA must own whole rows (it overwrites its own A operand).
template <int LAY, int TM>
__device__ __forceinline__ void g_load_tile(double* s, const double* __restrict__ g, int64_t ld, int64_t mn0, int k0, int tid) {
  if (LAY == LAY_KMAJ) {
#pragma unroll
    for (int i = 0; i < TM / 32; ++i) {          // TM rows x 8 granules of 16 B
      int q = tid + 256 * i;
      int r = q >> 3, gk = (q & 7) * 2;
      cp_async16(s + r * G_PITCH_K + gk, g + (mn0 + r) * ld + k0 + gk);
    }
  } else {
    constexpr int GPR = TM / 2;                   // granules per k-row
#pragma unroll
    for (int i = 0; i < TM / 32; ++i) {
      int q = tid + 256 * i;
      int kk = q / GPR, gm = (q % GPR) * 2;
      cp_async16(s + kk * (TM + 4) + gm, g + (int64_t)(k0 + kk) * ld + mn0 + gm);
    }
  }
}

template <int LAY, int TM>
__device__ __forceinline__ double g_frag(const double* s, int row, int k) {
  return (LAY == LAY_KMAJ) ? s[row * G_PITCH_K + k] : s[k * (TM + 4) + row];
}

// SUBC: the update C <- C - A*B (alpha = -1, beta = 1: every Cholesky trailing update).  The
// accumulators start as C, loaded while the cp.async prologue is in flight, and B fragments are
// negated, so the epilogue is a plain store and no C read sits exposed after the last MMA.
template <int ALAY, int BLAY, int TM, int TN, bool SUBC, int NSTAGE>
__global__ void __launch_bounds__(256, (TM == 128 && TN == 128) ? 1 : 2) gemm_f64_kernel(const GemmArgs p) {
  constexpr int MI = TM / 16, NI = TN / 32;     // 8x8 accumulator blocks per warp (rows x cols)
  constexpr int WM = TM / 2, WN = TN / 4;       // warp sub-tile
  constexpr int OPER_A = TM * G_PITCH_K;        // doubles per operand per stage (>= 16*(TM+4))
  constexpr int OPER_B = TN * G_PITCH_K;
  constexpr int STAGE = OPER_A + OPER_B;
  extern __shared__ __align__(16) double g_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;

  // ---- tile map -----------------------------------------------------------
  int mt, nt;
  if (p.tmap == TM_LOWER && TN * 2 == TM) {
    // half-width column tiles: row mt owns column tiles 0 .. 2*mt+1, mt*(mt+1) tiles precede it
    int L = blockIdx.x;
    mt = (int)((sqrt(4.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((mt + 1) * (mt + 2) <= L) ++mt;
    while (mt * (mt + 1) > L) --mt;
    nt = L - mt * (mt + 1);
  } else if (p.tmap == TM_LOWER) {
    int L = blockIdx.x;
    mt = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((mt + 1) * (mt + 2) / 2 <= L) ++mt;
    while (mt * (mt + 1) / 2 > L) --mt;
    nt = L - mt * (mt + 1) / 2;
  } else if (p.kbeg_rule == KB_MT || p.kend_rule == KE_MT1) {
    // k-extent depends on mt: issue the long tiles first (KE_MT1: large mt, KB_MT: small mt)
    int q = (int)(blockIdx.x / p.NT);
    mt = (p.kend_rule == KE_MT1) ? p.MT - 1 - q : q;
    nt = (int)(blockIdx.x % p.NT);
  } else {
    // KE_NT1: large nt is long; KB_NT: small nt is long; uniform otherwise
    int q = (int)(blockIdx.x / p.MT);
    nt = (p.kbeg_rule == KB_NT) ? q : p.NT - 1 - q;
    mt = (int)(blockIdx.x % p.MT);
  }
  const int z = blockIdx.z;
  if ((int64_t)mt * TM >= p.rows_left0 - (int64_t)z * p.rows_left_bs) return;

  const double* __restrict__ A = p.A + (int64_t)z * p.a_bs;
  const double* __restrict__ B = p.B + (int64_t)z * p.b_bs;
  double* __restrict__ C = p.C ? p.C + (int64_t)z * p.c_bs : nullptr;

  int kb = (p.kbeg_rule == KB_MT) ? mt * TM : (p.kbeg_rule == KB_NT) ? nt * TN : 0;
  int ke = (p.kend_rule == KE_MT1) ? (mt + 1) * TM : (p.kend_rule == KE_NT1) ? (nt + 1) * TN : p.K;
  if (ke > p.K) ke = p.K;
  const int KT = (ke > kb) ? (ke - kb) / GK : 0;
  const int64_t m0 = (int64_t)mt * TM, n0 = (int64_t)nt * TN;

  // ---- pipeline prologue --------------------------------------------------
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < KT) {
      double* sa = g_smem + (size_t)s * STAGE;
      g_load_tile<ALAY, TM>(sa, A, p.lda, m0, kb + s * GK, tid);
      g_load_tile<BLAY, TN>(sa + OPER_A, B, p.ldb, n0, kb + s * GK, tid);
    }
    cp_async_commit();
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      if (SUBC) {
        const double2 old = *reinterpret_cast<const double2*>(C + (m0 + wm * WM + i * 8 + g) * p.ldc + n0 + wn * WN + j * 8 + 2 * t);
        acc[i][j][0] = old.x;
        acc[i][j][1] = old.y;
      } else {
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      int nk = kt + NSTAGE - 1;
      if (nk < KT) {
        double* sa = g_smem + (size_t)(nk % NSTAGE) * STAGE;
        g_load_tile<ALAY, TM>(sa, A, p.lda, m0, kb + nk * GK, tid);
        g_load_tile<BLAY, TN>(sa + OPER_A, B, p.ldb, n0, kb + nk * GK, tid);
      }
      cp_async_commit();
    }
    const double* sa = g_smem + (size_t)(kt % NSTAGE) * STAGE;
    const double* sb = sa + OPER_A;
#pragma unroll
    for (int ks = 0; ks < GK / 4; ++ks) {
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = g_frag<ALAY, TM>(sa, wm * WM + i * 8 + g, ks * 4 + t);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const double b = g_frag<BLAY, TN>(sb, wn * WN + j * 8 + g, ks * 4 + t);
        bf[j] = SUBC ? -b : b;
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue -----------------------------------------------------------
  double rs[MI];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    rs[i] = 0.0;
    const int64_t r = m0 + wm * WM + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int64_t c = n0 + wn * WN + j * 8 + 2 * t;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (!SUBC) { v0 *= p.alpha; v1 *= p.alpha; }
      if (!SUBC && p.beta != 0.0) {
        double2 old = *reinterpret_cast<const double2*>(C + r * p.ldc + c);
        v0 += p.beta * old.x;
        v1 += p.beta * old.y;
      }
      if (p.store_c) *reinterpret_cast<double2*>(C + r * p.ldc + c) = make_double2(v0, v1);
      rs[i] += v0 * v0 + v1 * v1;
    }
  }
  if (p.rn_partial) {
    __syncthreads();                       // pipeline smem is dead: reuse it
    double* red = g_smem;                  // [4][TM]
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      double s = rs[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (t == 0) red[wn * TM + wm * WM + i * 8 + g] = s;
    }
    __syncthreads();
    if (tid < TM)
      p.rn_partial[(m0 + tid) * p.rn_nt + nt] = (red[tid] + red[TM + tid]) + (red[2 * TM + tid] + red[3 * TM + tid]);
  }
}

// host-side launcher (defined in chol.cu)
int gemm_f64_launch(const GemmArgs& a, int alay, int blay, int batch, cudaStream_t st);

static inline GemmArgs gemm_args_default() {
  GemmArgs a;
  a.A = nullptr; a.lda = 0; a.a_bs = 0;
  a.B = nullptr; a.ldb = 0; a.b_bs = 0;
  a.C = nullptr; a.ldc = 0; a.c_bs = 0;
  a.MT = a.NT = a.K = 0;
  a.tmap = TM_FULL; a.kbeg_rule = KB_ZERO; a.kend_rule = KE_FULL;
  a.alpha = 1.0; a.beta = 0.0;
  a.rows_left0 = ((int64_t)1) << 60; a.rows_left_bs = 0;
  a.rn_partial = nullptr; a.rn_nt = 0;
  a.store_c = 1;
  a.inplace_rows = 0;
  return a;
}
