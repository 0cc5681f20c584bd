// K1: fused kernel-matrix builder.
//
// Replaces GPR.cov_mat (reference models.py:161-181): s^2 k(x1,x2) for
// ScaleKernel(RBF | Matern-1.5) with ARD lengthscales, the optional
// diag(white_noise_var) and sigma_n^2 I adds fused in (the reference builds
// them as two extra N x N temporaries, models.py:175-180), the zero /
// identity padding the blocked factorisation wants, and optionally the
// row-dot with a vector (mean = K(X*,X) alpha, utils.py:301) so that the
// matrix is not re-read for the mean.
//
// Layout: out is row-major [n1_pad x ld].  A CTA owns a 64 x (128*VEC) tile;
// each warp a 32-row x (32*VEC)-col strip; a lane keeps the (scaled)
// coordinates of its VEC consecutive columns in registers, reads the row's
// coordinates as a shared-memory broadcast, and stores 16 bytes per row, so a
// warp writes 512 contiguous bytes per row.  HBM-store bound (8 B/element,
// 4 B in float32 mode) unless the fp64 exp saturates the FP64 pipe first.
#include "common.cuh"

struct KbuildArgs {
  KernelParams kp;
  const double* x1;
  const double* x2;
  int64_t n1, n2;           // valid rows / cols
  int64_t n1_pad, n2_pad;   // written extent (>= n1, n2)
  const double* diag_add;   // [n1] or null: added where r == c
  double diag_scalar;       // added where r == c (r < n1)
  int pad_identity;         // 1: out[r][r] = 1 for r >= n1 (keeps padded factor SPD)
  void* out;
  int64_t ld;
  const double* dot_vec;    // [n2] or null
  double* dot_partial;      // [n1_pad x n_col_tiles]
  int n_col_tiles;
};

template <typename T> struct Vec;
template <> struct Vec<double> { static constexpr int N = 2; typedef double2 type; };
template <> struct Vec<float> { static constexpr int N = 4; typedef float4 type; };

// Coordinates are pre-scaled by inv_ls * KSCALE so that q = -sum df^2 is directly the exponent:
//   RBF     KSCALE = sqrt(1/2):  k = s^2 exp(q)
//   Matern  KSCALE = sqrt(3):    k = s^2 (1 + r') exp(-r'),  r' = sqrt(-q)
// and s^2 is folded into the exp table, so an RBF element costs 11 fp64 ops after the distance.
template <int KIND> __host__ __device__ constexpr double kscale() { return KIND == 0 ? 0.70710678118654752 : 1.7320508075688772; }
template <typename T, int KIND> __device__ __forceinline__ T kern_eval(T q, const double* tab, T os);
template <> __device__ __forceinline__ double kern_eval<double, 0>(double q, const double* tab, double) { return exp_nonpos_tab(q, tab); }
template <> __device__ __forceinline__ double kern_eval<double, 1>(double q, const double* tab, double) {
  const double r = sqrt(-q);
  return (1.0 + r) * exp_nonpos_tab(-r, tab);
}
template <> __device__ __forceinline__ float kern_eval<float, 0>(float q, const double*, float os) { return os * expf(q); }
template <> __device__ __forceinline__ float kern_eval<float, 1>(float q, const double*, float os) {
  const float r = sqrtf(-q);
  return os * (1.0f + r) * expf(-r);
}

#define KB_ROWS 64

// FAST: the tile lies strictly inside [0,n1) x [0,n2) and carries no diagonal work, so the
// row loop is distance -> kernel -> 16-byte store with no masking.
template <typename T, int D, int KIND, bool DOT, bool FAST>
__device__ __forceinline__ void kbuild_rows(const KbuildArgs& a, const T (*sx1)[ALGP_MAX_D], const T (*sx2)[128 * Vec<T>::N],
                                            const T (*x2r)[D == 0 ? 1 : D], const T* dv, double (*sred)[4], const double* tab,
                                            int d, int wr, int wc, int lane, int64_t row0, int64_t c0, int lc) {
  constexpr int VEC = Vec<T>::N;
  const T os = (T)a.kp.outputscale;
  T* outp0 = (T*)a.out + (row0 + (int64_t)wr * 32) * a.ld + c0;
  const bool col_ok = FAST || (c0 < a.n2_pad);
  // rows in groups of 8: with the fused row-dot the 8 per-lane partials are reduced together by a
  // butterfly reduce-scatter (9 double shuffles per 8 rows instead of 40)
#pragma unroll 1
  for (int r8 = 0; r8 < 32; r8 += 8) {
    T* outp = outp0;
    double part[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = wr * 32 + r8 + u;
      const int64_t gr = row0 + r;
      part[u] = 0.0;
      if (!FAST && gr >= a.n1_pad) continue;       // warp-uniform
      T val[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        T q = (T)0;                                  // -sum df^2 over pre-scaled coordinates
        if (D == 0) {
          for (int j = 0; j < d; ++j) {
            T df = sx1[r][j] - sx2[j][lc + v];
            q = fma(-df, df, q);
          }
        } else {
#pragma unroll
          for (int j = 0; j < (D == 0 ? 1 : D); ++j) {
            T df = sx1[r][j] - x2r[v][j];
            q = fma(-df, df, q);
          }
        }
        T k = kern_eval<T, KIND>(q, tab, os);
        if (!FAST) {
          const int64_t gc = c0 + v;
          if (!((gr < a.n1) && (gc < a.n2))) k = (T)0;
          if (gr == gc) {
            if (gr < a.n1) {
              double add = a.diag_scalar + (a.diag_add ? a.diag_add[gr] : 0.0);
              k = (T)((double)k + add);
            } else if (a.pad_identity) {
              k = (T)1;
            }
          }
        }
        val[v] = k;
      }
      if (col_ok) {
        typename Vec<T>::type pk;
#pragma unroll
        for (int v = 0; v < VEC; ++v) ((T*)&pk)[v] = val[v];
        *reinterpret_cast<typename Vec<T>::type*>(outp + (int64_t)(r8 + u) * a.ld) = pk;
      }
      if (DOT) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) part[u] = fma((double)val[v], (double)dv[v], part[u]);
      }
    }
    if (DOT) {
      // reduce-scatter: after the three exchange steps lane l holds the sum over its 8-lane group of row (l & 7)...
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int h = 4 >> s;                        // 4, 2, 1 rows kept per lane after this step
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (u < h) {
            const double keep = upper ? part[u + h] : part[u];
            const double send = upper ? part[u] : part[u + h];
            part[u] = keep + __shfl_xor_sync(0xffffffffu, send, h);
          }
        }
      }
      // ... then two plain steps across the four 8-lane groups
      double sum = part[0];
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      // lane l (< 8) holds row index: bit 2 of l selected rows {4..7}, bit 1 then {2,3}/{6,7}, bit 0 the odd one
      if (lane < 8) sred[wr * 32 + r8 + lane][wc] = sum;
    }
  }
}

template <typename T, int D, int KIND, bool DOT>
__global__ void __launch_bounds__(256) kbuild_kernel(const KbuildArgs a) {
  constexpr int VEC = Vec<T>::N;
  constexpr int TILE_C = 128 * VEC;
  __shared__ __align__(16) T sx1[KB_ROWS][ALGP_MAX_D];
  __shared__ T sx2[D == 0 ? ALGP_MAX_D : 1][TILE_C];
  __shared__ double sred[KB_ROWS][4];
  __shared__ double stab[16];

  const int d = (D == 0) ? a.kp.d : D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 2, wc = warp & 3;
  if (tid < 16) stab[tid] = a.kp.outputscale * c_exp2_16th[tid];
  const int64_t row0 = (int64_t)blockIdx.y * KB_ROWS;
  const int64_t col0 = (int64_t)blockIdx.x * TILE_C;
  const int64_t c0 = col0 + (int64_t)wc * 32 * VEC + lane * VEC;

  // stage the scaled row coordinates (rows past n1 read as 0; masked later)
  for (int i = tid; i < KB_ROWS * d; i += 256) {
    int r = i / d, j = i - r * d;
    int64_t gr = row0 + r;
    double v = (gr < a.n1) ? a.x1[gr * d + j] : 0.0;
    sx1[r][j] = (T)((T)v * (T)(a.kp.inv_ls[j] * kscale<KIND>()));
  }
  T x2r[VEC][D == 0 ? 1 : D];
  if (D == 0) {
    for (int i = tid; i < TILE_C * d; i += 256) {
      int c = i / d, j = i - c * d;
      int64_t gc = col0 + c;
      double v = (gc < a.n2) ? a.x2[gc * d + j] : 0.0;
      sx2[j][c] = (T)((T)v * (T)(a.kp.inv_ls[j] * kscale<KIND>()));
    }
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
      for (int j = 0; j < (D == 0 ? 1 : D); ++j) {
        int64_t gc = c0 + v;
        double xv = (gc < a.n2) ? a.x2[gc * d + j] : 0.0;
        x2r[v][j] = (T)((T)xv * (T)(a.kp.inv_ls[j] * kscale<KIND>()));
      }
  }
  T dv[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dv[v] = (DOT && c0 + v < a.n2) ? (T)a.dot_vec[c0 + v] : (T)0;
  __syncthreads();

  const int lc = wc * 32 * VEC + lane * VEC;   // column inside the tile
  const bool diag_work = (a.diag_add != nullptr) || (a.diag_scalar != 0.0) || (a.pad_identity != 0);
  const bool touches_diag = (row0 < col0 + TILE_C) && (col0 < row0 + KB_ROWS);
  const bool fast = (row0 + KB_ROWS <= a.n1) && (col0 + TILE_C <= a.n2) && !(diag_work && touches_diag);
  if (fast)
    kbuild_rows<T, D, KIND, DOT, true>(a, sx1, sx2, x2r, dv, sred, stab, d, wr, wc, lane, row0, c0, lc);
  else
    kbuild_rows<T, D, KIND, DOT, false>(a, sx1, sx2, x2r, dv, sred, stab, d, wr, wc, lane, row0, c0, lc);

  if (DOT) {
    __syncthreads();
    if (tid < KB_ROWS) {
      int64_t gr = row0 + tid;
      if (gr < a.n1_pad)
        a.dot_partial[gr * a.n_col_tiles + blockIdx.x] = (sred[tid][0] + sred[tid][1]) + (sred[tid][2] + sred[tid][3]);
    }
  }
}

// ---------------------------------------------------------------------------
// Symmetric build K(X,X): one CTA per lower 128x128 tile (I >= J) evaluates the tile once and
// stores it twice, at (I,J) and transposed at (J,I), halving the fp64 work.  A lane owns 2x2
// micro-blocks (rows 2mr..2mr+1, cols 2mc..2mc+1 of an 8x16 warp patch), so both stores are
// 16-byte vectors: 8 lanes cover 128 contiguous bytes of a direct row, 4 lanes cover 64
// contiguous bytes of a mirrored row -- no shared-memory transpose.
// ---------------------------------------------------------------------------
#define KS_T 128
template <int D, int KIND>
__global__ void __launch_bounds__(256) kbuild_sym_kernel(const KbuildArgs a) {
  __shared__ __align__(16) double sxr[KS_T][D], sxc[KS_T][D];
  __shared__ double stab[16];
  const int L = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= L) ++ti;
  while (ti * (ti + 1) / 2 > L) --ti;
  const int tj = L - ti * (ti + 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)ti * KS_T, c0 = (int64_t)tj * KS_T;
  if (tid < 16) stab[tid] = a.kp.outputscale * c_exp2_16th[tid];
  {
    const int64_t gx = (tid < KS_T) ? r0 + tid : c0 + (tid - KS_T);
    double* dst = (tid < KS_T) ? sxr[tid] : sxc[tid - KS_T];
#pragma unroll
    for (int j = 0; j < D; ++j) dst[j] = (gx < a.n1) ? a.x1[gx * D + j] * (a.kp.inv_ls[j] * kscale<KIND>()) : 0.0;
  }
  __syncthreads();

  const int mr = lane >> 3, mc = lane & 7;
  const bool mirror = (ti != tj);
  const bool fast = mirror && (r0 + KS_T <= a.n1);          // off-diagonal and fully inside: no masks
  const double os = a.kp.outputscale;
  double* outp = (double*)a.out;
#pragma unroll 1
  for (int pr = 2 * warp; pr < 2 * warp + 2; ++pr) {
    const int rl = pr * 8 + 2 * mr;                          // local row of the micro-block
    double xr[2][D];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int j = 0; j < D; ++j) xr[e][j] = sxr[rl + e][j];
#pragma unroll 2
    for (int pc = 0; pc < 8; ++pc) {
      const int cl = pc * 16 + 2 * mc;
      double v[2][2];
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        double xc[D];
#pragma unroll
        for (int j = 0; j < D; ++j) xc[j] = sxc[cl + f][j];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double q = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double df = xr[e][j] - xc[j];
            q = fma(-df, df, q);
          }
          double k = kern_eval<double, KIND>(q, stab, os);
          if (!fast) {
            const int64_t gr = r0 + rl + e, gc = c0 + cl + f;
            if (!((gr < a.n1) && (gc < a.n1))) k = 0.0;
            if (gr == gc) {
              if (gr < a.n1) k += a.diag_scalar + (a.diag_add ? a.diag_add[gr] : 0.0);
              else if (a.pad_identity) k = 1.0;
            }
          }
          v[e][f] = k;
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
        *reinterpret_cast<double2*>(outp + (r0 + rl + e) * a.ld + c0 + cl) = make_double2(v[e][0], v[e][1]);
      if (mirror) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
          *reinterpret_cast<double2*>(outp + (c0 + cl + f) * a.ld + r0 + rl) = make_double2(v[0][f], v[1][f]);
      }
    }
  }
}

template <int KIND>
static int launch_kbuild_sym(const KbuildArgs& a, cudaStream_t st) {
  const int64_t nt = a.n1_pad / KS_T;
  const int64_t tiles = nt * (nt + 1) / 2;
  if (tiles > 0x7fffffffLL) return ALGP_ERR_INVALID;
  switch (a.kp.d) {
    case 2: kbuild_sym_kernel<2, KIND><<<(unsigned)tiles, 256, 0, st>>>(a); break;
    case 3: kbuild_sym_kernel<3, KIND><<<(unsigned)tiles, 256, 0, st>>>(a); break;
    case 4: kbuild_sym_kernel<4, KIND><<<(unsigned)tiles, 256, 0, st>>>(a); break;
    case 6: kbuild_sym_kernel<6, KIND><<<(unsigned)tiles, 256, 0, st>>>(a); break;
    default: return ALGP_ERR_UNSUPPORTED;       // register-resident coordinates: small d only (caller falls back)
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// out[r] = bias + scale * sum_t partial[r][t]   (fixed order: deterministic)
__global__ void rowsum_kernel(const double* __restrict__ partial, int64_t rows, int nt, double scale, double bias,
                              const double* __restrict__ addvec, double* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s = 0.0;
  for (int t = 0; t < nt; ++t) s += partial[r * nt + t];
  out[r] = bias + scale * s + (addvec ? addvec[r] : 0.0);
}

// Sigma[loc,B_k] gains sigma_n^2 where field location loc IS base point k
// (agent.py:90 builds cov_matrix with add_likelihood_var=True, so the noise
// sits on the n x n diagonal and travels with the (loc, base) gather).
__global__ void scatter_add_kernel(double* __restrict__ M, int64_t ld, const int32_t* __restrict__ row_of_col, int64_t ncols,
                                   double v) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncols) return;
  int32_t r = row_of_col[k];
  if (r >= 0) M[(int64_t)r * ld + k] += v;
}

template <typename T, int KIND, bool DOT>
static int launch_kbuild_k(const KbuildArgs& a, cudaStream_t st) {
  constexpr int TILE_C = 128 * Vec<T>::N;
  dim3 grid((unsigned)((a.n2_pad + TILE_C - 1) / TILE_C), (unsigned)((a.n1_pad + KB_ROWS - 1) / KB_ROWS));
  if (grid.y > 65535) return ALGP_ERR_INVALID;
  switch (a.kp.d) {
    case 2: kbuild_kernel<T, 2, KIND, DOT><<<grid, 256, 0, st>>>(a); break;
    case 6: kbuild_kernel<T, 6, KIND, DOT><<<grid, 256, 0, st>>>(a); break;
    default: kbuild_kernel<T, 0, KIND, DOT><<<grid, 256, 0, st>>>(a); break;
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

template <typename T, bool DOT>
static int launch_kbuild(const KbuildArgs& a, cudaStream_t st) {
  return a.kp.kind == 0 ? launch_kbuild_k<T, 0, DOT>(a, st) : launch_kbuild_k<T, 1, DOT>(a, st);
}

int make_kernel_params(KernelParams* kp, int d, const double* log_ls_host, double log_os, int kind);

extern "C" int algp_kbuild_col_tiles(int64_t n2_pad, int out_dtype) {
  int tile = 128 * (out_dtype == 0 ? 2 : 4);
  return (int)((n2_pad + tile - 1) / tile);
}

extern "C" int algp_kbuild(const double* x1, int64_t n1, const double* x2, int64_t n2, int d,
                           const double* log_ls_host, double log_os, int kind,
                           const double* diag_add, double diag_scalar, int pad_identity,
                           void* out, int64_t n1_pad, int64_t n2_pad, int64_t ld, int out_dtype,
                           const double* dot_vec, double* dot_partial, void* stream) {
  if (!x1 || !out || n1 < 0 || n2 < 0 || n1_pad < n1 || n2_pad < n2 || ld < n2_pad) return ALGP_ERR_INVALID;
  if (out_dtype != 0 && out_dtype != 1) return ALGP_ERR_INVALID;
  if ((dot_vec == nullptr) != (dot_partial == nullptr)) return ALGP_ERR_INVALID;
  // 16-byte row stores: rows must start 16-byte aligned
  int vec = out_dtype == 0 ? 2 : 4;
  if (ld % vec || n2_pad % vec || ((uintptr_t)out & 15)) return ALGP_ERR_INVALID;
  KbuildArgs a;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.x1 = x1; a.x2 = x2 ? x2 : x1; a.n1 = n1; a.n2 = n2; a.n1_pad = n1_pad; a.n2_pad = n2_pad;
  a.diag_add = diag_add; a.diag_scalar = diag_scalar; a.pad_identity = pad_identity;
  a.out = out; a.ld = ld; a.dot_vec = dot_vec; a.dot_partial = dot_partial;
  a.n_col_tiles = algp_kbuild_col_tiles(n2_pad, out_dtype);
  if (n1_pad == 0 || n2_pad == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // K(X,X) in fp64 on a square padded extent: lower tiles only, mirrored on the way out
  if (!x2 && out_dtype == 0 && !dot_vec && n1_pad == n2_pad && n1_pad % KS_T == 0 && n1 == n2 &&
      (d == 2 || d == 3 || d == 4 || d == 6))
    return a.kp.kind == 0 ? launch_kbuild_sym<0>(a, st) : launch_kbuild_sym<1>(a, st);
  if (out_dtype == 0) return dot_vec ? launch_kbuild<double, true>(a, st) : launch_kbuild<double, false>(a, st);
  return dot_vec ? launch_kbuild<float, true>(a, st) : launch_kbuild<float, false>(a, st);
}

extern "C" int algp_rowsum(const double* partial, int64_t rows, int nt, double scale, double bias,
                           const double* addvec, double* out, void* stream) {
  if (!partial || !out || rows < 0 || nt < 0) return ALGP_ERR_INVALID;
  if (rows == 0) return ALGP_OK;
  rowsum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, rows, nt, scale, bias, addvec, out);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

extern "C" int algp_scatter_add(double* M, int64_t ld, const int32_t* row_of_col, int64_t ncols, double v, void* stream) {
  if (!M || !row_of_col || ncols < 0) return ALGP_ERR_INVALID;
  if (ncols == 0) return ALGP_OK;
  scatter_add_kernel<<<(unsigned)((ncols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, ld, row_of_col, ncols, v);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
