// Candidate-set scoring, k <= 8, with the working set tiled for the L2 (opt-in variant of the headline kernel of
// BASELINE configs[2]; PosteriorState.score_mode = "tiled").
//
// score.cu's score_sets_k8_kernel streams the k rows of Wt of a candidate end to end: with 65 536 random sets over
// 12 288 distinct rows of N = 4096 doubles the rows (403 MB) do not fit the 126 MB L2 and every row is re-fetched from
// DRAM ~19 times (ncu round 1: 7.7 GB of DRAM reads for 0.4 GB of compulsory traffic).  Here the N columns are cut
// into chunks whose slice of Wt (rows x chunk x 8 B) stays L2-resident, one launch per chunk:
//
//   score_slots_kernel      one thread per candidate: the slot rule (empty / duplicate / zero-increment / skipped)
//                           evaluated once per call into a table of row indices.
//   score_gram_k8_kernel    one warp per candidate and chunk: lane l owns columns {2l, 2l+1} (+64 per step), loads the
//                           8 row segments with 16-byte loads (512 contiguous bytes per row and warp request) and
//                           accumulates the 36 DISTINCT entries of the symmetric Gram G = Wt_C Wt_C^T with plain DFMA
//                           (the DMMA.8x8x4 form computes all 64 entries: 64 -> 36 FMAs per column, and B200's DMMA
//                           and DFMA rates are equal); a transposing butterfly leaves entry e in lane e, which adds
//                           it (red.global.add.f64) to the partial Gram buffer G[B][36] (L2-resident, 19 MB).
//   score_finish_k8_kernel  one THREAD per candidate: Sigma_CC from the coordinates, P_CC = Sigma_CC - G, the 8x8
//                           un-normalised elimination in registers (the order of score_sets_k8_kernel and
//                           score_cov_k8_kernel) and the size / precision bookkeeping of SURVEY.md 9.3.
//
// MEASURED (B200, round 2, profiles/r02_score_tiling.log): 1.67 ms at a 1024-column chunk against 1.66 ms for the
// row-streaming kernel -- the DRAM re-fetch disappears but the time does not move, because both kernels deliver
// 17.2 GB into the SMs at ~10.4 TB/s, which is the L2 -> SM (LTS) throughput cap of the chip; a chunked DMMA variant
// was slower (2.3-2.7 ms).  Streaming k rows per candidate is therefore at its roof; the faster algorithm is the
// resident posterior covariance (scorecov.cu).  The row-streaming kernel stays the default.
//
// Slot semantics are those of score.cu (reference agent.py:373-400): idx < 0, delta <= 0 or skip[idx] = empty slot,
// duplicates inside a set count once (the first).
#include "common.cuh"
#include <math.h>

#define ALGP_CONST 1.4189385332046727   // 0.5*log(2*pi*e), utils.py:10

int make_kernel_params(KernelParams* kp, int d, const double* log_ls_host, double log_os, int kind);

namespace {

struct TileArgs {
  KernelParams kp;
  double noise;
  const double* Wt;
  int64_t ldw;
  int ncols16;             // valid columns rounded up to 16 (tail columns are zero)
  int col0, col1;          // this launch covers columns [col0, col1), col0 a multiple of 64
  int first;               // 1: G = partial, 0: G += partial
  const double* X;
  const double* pi0;
  const int32_t* idx;
  const double* delta;
  double delta_scalar;
  const uint8_t* skip;
  int k;
  int64_t B;
  double H_base;
  double* G;               // [B][36] partial lower-triangle Gram matrices, entry e = i (i + 1) / 2 + j, j <= i
  int32_t* rowoff;         // [B][8] location of every ACTIVE slot, -1 for empty / duplicate / zero-increment / skipped
  double* scores;
};

__device__ __forceinline__ double2 ld128(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// active slots of candidate `cand` as a bit mask (bit s = slot s contributes): the rule of score_sets_k8_kernel
__device__ __forceinline__ unsigned slot_mask(const TileArgs& a, int64_t cand, int (&ix)[8], double (&dl)[8]) {
  unsigned m = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    int id = -1;
    double d = 0.0;
    if (s < a.k) {
      id = a.idx[cand * a.k + s];
      d = a.delta ? a.delta[cand * a.k + s] : a.delta_scalar;
    }
    bool act = id >= 0 && d > 0.0;
    if (act && a.skip && a.skip[id]) act = false;
#pragma unroll
    for (int q = 0; q < s; ++q)
      if ((m >> q & 1u) && ix[q] == id) act = false;
    ix[s] = id;
    dl[s] = d;
    if (act) m |= 1u << s;
  }
  return m;
}

// one thread per candidate: the slot rule evaluated once per call instead of once per chunk
__global__ void __launch_bounds__(128) score_slots_kernel(const TileArgs a) {
  const int64_t cand = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= a.B) return;
  int ix[8];
  double dl[8];
  const unsigned mask = slot_mask(a, cand, ix, dl);
  int4 lo = make_int4((mask & 1u) ? ix[0] : -1, (mask & 2u) ? ix[1] : -1, (mask & 4u) ? ix[2] : -1, (mask & 8u) ? ix[3] : -1);
  int4 hi = make_int4((mask & 16u) ? ix[4] : -1, (mask & 32u) ? ix[5] : -1, (mask & 64u) ? ix[6] : -1, (mask & 128u) ? ix[7] : -1);
  int4* dst = reinterpret_cast<int4*>(a.rowoff + cand * 8);
  dst[0] = lo;
  dst[1] = hi;
}

__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <int STEPS>   // 64-column steps whose loads are issued together
__global__ void __launch_bounds__(128, 3) score_gram_k8_kernel(const TileArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int cend = a.col1 < a.ncols16 ? a.col1 : a.ncols16;

  for (int64_t cand = warp0; cand < a.B; cand += nwarps) {
    const int4 lo = *reinterpret_cast<const int4*>(a.rowoff + cand * 8);
    const int4 hi = *reinterpret_cast<const int4*>(a.rowoff + cand * 8 + 4);
    const int ix[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned mask = 0;
    const double* row[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (ix[s] >= 0) mask |= 1u << s;
      row[s] = a.Wt + (int64_t)(ix[s] >= 0 ? ix[s] : 0) * a.ldw + 2 * lane;
    }

    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;

    for (int c = a.col0; c < cend; c += 64 * STEPS) {
      double2 v[STEPS][8];
#pragma unroll
      for (int u = 0; u < STEPS; ++u) {
        const int cc = c + 64 * u;
        const bool in = cc + 2 * lane < cend;          // cend is a multiple of 16, the lane's pair is inside or outside
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          v[u][s] = make_double2(0.0, 0.0);
          if (in && (mask >> s & 1u)) v[u][s] = ld128(row[s] + cc);
        }
      }
#pragma unroll
      for (int u = 0; u < STEPS; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            const int e = i * (i + 1) / 2 + j;
            acc[e] = fma(v[u][i].x, v[u][j].x, acc[e]);
            acc[e] = fma(v[u][i].y, v[u][j].y, acc[e]);
          }
        }
      }
    }

    // transposing butterfly over the first 32 entries: after the step with offset h a lane keeps the half of its
    // values whose index bit matches its own lane bit, so five steps (16 + 8 + 4 + 2 + 1 exchanges) leave the
    // warp-wide sum of entry e in lane e
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
      const bool up = (lane & h) != 0;
#pragma unroll
      for (int e = 0; e < h; ++e) {
        const double keep = up ? acc[e + h] : acc[e];
        const double send = up ? acc[e] : acc[e + h];
        acc[e] = keep + __shfl_xor_sync(0xffffffffu, send, h);
      }
    }
    // entries 32..35: plain butterflies (every lane ends with the sums)
#pragma unroll
    for (int e = 32; e < 36; ++e) acc[e] = warp_sum(acc[e]);

    // lane -> entry it holds after the butterfly: bit b of the entry index equals bit b of the lane
    double* g = a.G + cand * 36;
    double tail = acc[32];
    if (lane == 1) tail = acc[33];
    if (lane == 2) tail = acc[34];
    if (lane == 3) tail = acc[35];
    if (a.first) {
      g[lane] = acc[0];
      if (lane < 4) g[32 + lane] = tail;
    } else {
      red_add(g + lane, acc[0]);
      if (lane < 4) red_add(g + 32 + lane, tail);
    }
  }
}

__global__ void __launch_bounds__(128) score_finish_k8_kernel(const TileArgs a) {
  const int64_t cand = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= a.B) return;
  int ix[8];
  double dl[8];
  const unsigned mask = slot_mask(a, cand, ix, dl);
  const int d = a.kp.d;
  double sq[8];
  double term = 0.0, nnew = 0.0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const bool act = mask >> s & 1u;
    sq[s] = act ? sqrt(dl[s]) : 0.0;
    if (act) {
      const double p0 = a.pi0[ix[s]];
      term += log(p0 + dl[s]) - (p0 > 0.0 ? log(p0) : 0.0);
      nnew += (p0 > 0.0) ? 0.0 : 1.0;
    }
  }
  double m[36];
  const double* g = a.G + cand * 36;
#pragma unroll
  for (int e = 0; e < 36; ++e) m[e] = g[e];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      double val = (i == j) ? 1.0 : 0.0;
      if ((mask >> i & 1u) && (mask >> j & 1u)) {
        double r2 = 0.0;
        for (int q = 0; q < d; ++q) {
          const double df = (a.X[(int64_t)ix[i] * d + q] - a.X[(int64_t)ix[j] * d + q]) * a.kp.inv_ls[q];
          r2 = fma(df, df, r2);
        }
        const double sig = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((i == j) ? a.noise : 0.0);
        val = fma(sig - m[e], sq[i] * sq[j], val);
      }
      m[e] = val;
    }
  }
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

int g_tile_cols = 0;       // 0 = choose from the L2 size; set by algp_set_score_tile_cols (tuning / tests)

}  // namespace

extern "C" int algp_set_score_tile_cols(int cols) {
  if (cols < 0 || (cols & 63)) return ALGP_ERR_INVALID;
  g_tile_cols = cols;
  return ALGP_OK;
}

// partial Gram matrices [B][36] + slot table [B][8] int32
extern "C" int64_t algp_score_sets_tiled_work_doubles(int64_t B) { return B > 0 ? B * 40 : 0; }

// scores[c] = H(base set + candidate set c) for sets of k <= 8 slots, as algp_score_sets, with the columns of Wt
// processed in L2-sized chunks.  n_rows: number of rows of Wt the candidates can reference (sizes the chunk so that
// n_rows x chunk x 8 bytes stays L2-resident).  work: algp_score_sets_tiled_work_doubles(B) doubles.
extern "C" int algp_score_sets_tiled(const double* Wt, int64_t ldw, int64_t ncols, int64_t n_rows, const double* X, int d,
                                     const double* log_ls_host, double log_os, int kind, double noise, const double* pi0,
                                     const int32_t* idx, const double* delta, double delta_scalar, const uint8_t* skip,
                                     int k, int64_t B, double H_base, double* scores, double* work, int64_t work_doubles,
                                     void* stream) {
  if (!Wt || !X || !pi0 || !idx || !scores || k < 1 || k > 8 || B < 0 || ncols < 0 || n_rows < 1) return ALGP_ERR_INVALID;
  if ((ldw & 1) || ((uintptr_t)Wt & 15)) return ALGP_ERR_INVALID;       // 128-bit row loads
  if (B > 0 && (!work || work_doubles < B * 40 || ((uintptr_t)work & 15))) return ALGP_ERR_INVALID;
  TileArgs a;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.noise = noise; a.Wt = Wt; a.ldw = ldw;
  a.ncols16 = (int)((ncols + 15) / 16 * 16);
  if (a.ncols16 > ldw) return ALGP_ERR_INVALID;
  a.X = X; a.pi0 = pi0; a.idx = idx; a.delta = delta; a.delta_scalar = delta_scalar; a.skip = skip;
  a.k = k; a.B = B; a.H_base = H_base; a.G = work; a.rowoff = (int32_t*)(work + B * 36); a.scores = scores;
  if (B == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148, l2 = 64 << 20;
  ALGP_CUDA(cudaGetDevice(&dev));
  ALGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ALGP_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
  // chunk: the slice of Wt (n_rows x chunk doubles) takes at most ~40 % of the L2, the rest is left to the partial
  // Gram buffer, the candidate arrays and the next slice's first touches
  int chunk = g_tile_cols;
  if (chunk == 0) {
    int64_t c = (int64_t)(0.4 * (double)l2) / (8 * n_rows);
    c = c / 64 * 64;
    if (c < 128) c = 128;
    chunk = (int)(c > a.ncols16 ? (a.ncols16 + 63) / 64 * 64 : c);
    if (chunk < 64) chunk = 64;
  }
  score_slots_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  const int64_t want = (B + 3) / 4;
  const int64_t cap = (int64_t)sms * 3;
  const int grid = (int)(want < cap ? want : cap);
  for (int c0 = 0, first = 1; c0 < a.ncols16 || first; c0 += chunk, first = 0) {
    a.col0 = c0;
    a.col1 = c0 + chunk;
    a.first = first;
    score_gram_k8_kernel<2><<<grid, 128, 0, st>>>(a);
    ALGP_LAUNCH_CHECK();
  }
  score_finish_k8_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
