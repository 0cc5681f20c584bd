// TF32 mode of the variance path on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// var_m = k** - |V_m|^2, V = K(X*,X) L^-T (reference utils.py:300-308).  In fp64 this N^2 M step is
// DMMA-bound (35 TFLOP/s); tcgen05 has no f64 kind, so the fast mode runs it as a split-TF32 product
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (a_hi = a truncated to TF32, a_lo = a - a_hi)
// with fp32 accumulation in TMEM: ~1e-6 relative per product, which the k** - |V|^2 cancellation
// needs (plain TF32, 1e-3, would wipe the variance out).  The factor itself stays fp64.
//
//   split_tf32_kernel   fp64 matrix -> fp32 hi / lo planes (one HBM pass; hi has its low 13 mantissa
//                       bits cleared so the tensor core's own truncation is exact)
//   trmm_tf32x3_kernel  one CTA per 128 x 256 output tile: 4 epilogue warps (one lane of warp 0 is
//                       the TMA producer) + 1 MMA warp.  The producer issues four 2-D TMA tile
//                       loads per 16-wide k-stage (A_hi, A_lo: 128 x 64 B; B_hi, B_lo: 256 x 64 B,
//                       SWIZZLE_64B) that complete on the stage's `full` mbarrier; one thread issues
//                       6 tcgen05.mma.kind::tf32 (M128 N256 K8) per stage into a 256-column TMEM
//                       accumulator and tcgen05.commit releases the stage (`empty` mbarrier):
//                       a 4-stage ring with no thread ever touching operand bytes.  Epilogue:
//                       tcgen05.ld 32x32b, a thread owns a whole accumulator row, squares and sums
//                       it -- V is never written.  Triangular k-range (Linv is lower triangular).
#include "common.cuh"
#include <cuda.h>

#define T3_TM 128
#define T3_TN 256
#define T3_KC 16                       // fp32 per row per stage = 4 chunks of 16 B
#define T3_STAGES 4
#define T3_A_BYTES (T3_TM * T3_KC * 4) // 8 KB
#define T3_B_BYTES (T3_TN * T3_KC * 4) // 16 KB
#define T3_STAGE_BYTES (2 * T3_A_BYTES + 2 * T3_B_BYTES)   // A_hi, A_lo, B_hi, B_lo = 48 KB
#define T3_SMEM_BYTES (T3_STAGES * T3_STAGE_BYTES + 1024)   // + slack to align the ring to 1024 B
#define T3_THREADS 160

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, single CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_64B operand tile: rows of 64 B, 8-row groups 512 B apart (SBO); the 16-byte chunks of
// a row are XOR-swizzled by the hardware on both the TMA write and the MMA read
// (cute::UMMA::SmemDescriptor: version 1, layout_type 4, LBO field 1 for swizzled K-major layouts)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c_inner, int c_outer, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
      : "memory");
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format TF32 (2) [7,10) / [10,13), K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void split_tf32_kernel(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ldo) {
  // 4 consecutive columns per thread: 2 x 16-byte loads, 2 x 16-byte stores
  const int64_t c4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t r = blockIdx.y;
  if (c4 >= cols || r >= rows) return;
  const double2 a = *reinterpret_cast<const double2*>(src + r * ld + c4);
  const double2 b = *reinterpret_cast<const double2*>(src + r * ld + c4 + 2);
  const double x[4] = {a.x, a.y, b.x, b.y};
  float h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float f = (float)x[i];
    h[i] = __uint_as_float(__float_as_uint(f) & 0xFFFFE000u);
    l[i] = (float)(x[i] - (double)h[i]);
  }
  *reinterpret_cast<float4*>(hi + r * ldo + c4) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(lo + r * ldo + c4) = make_float4(l[0], l[1], l[2], l[3]);
}

struct T3Args {
  int MT, NT;                                        // 128-row tiles, 256-col tiles (last may be half)
  int64_t npad;
  double* rn_partial; int rn_nt;                     // [mpad x npad/128]
};

__global__ void __launch_bounds__(T3_THREADS, 1)
trmm_tf32x3_kernel(const T3Args p, const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                   const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo) {
  extern __shared__ unsigned char t3_smem_raw[];
  __shared__ uint64_t full_bar[T3_STAGES], empty_bar[T3_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ring = (smem_u32(t3_smem_raw) + 1023u) & ~1023u;      // swizzle atoms need aligned tiles

  // tile map: groups of 16 m-tiles sweep the n-tiles together, long (large nt) tiles first, so that
  // concurrent CTAs stream the same k-window of A and the same rows of B through L2
  constexpr int GM = 16;
  const int per_group = GM * p.NT;
  const int mg = blockIdx.x / per_group, rem = blockIdx.x % per_group;
  const int nt = p.NT - 1 - rem / GM;
  const int mt = mg * GM + rem % GM;
  const bool valid_tile = mt < p.MT;                 // block-uniform
  const int64_t m0 = (int64_t)mt * T3_TM, n0 = (int64_t)nt * T3_TN;
  const int ncols = (int)((p.npad - n0) < T3_TN ? (p.npad - n0) : T3_TN);   // 128 or 256
  const int64_t kend = n0 + ncols;                   // Linv[j][k] = 0 for k > j
  const int KT = valid_tile ? (int)(kend / T3_KC) : 0;

  if (tid == 0) {
    for (int s = 0; s < T3_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(T3_TN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    if (tid == 0) {
      // ===================== TMA producer (one thread) =====================
      for (int it = 0; it < KT; ++it) {
        const int s = it % T3_STAGES, u = it / T3_STAGES;
        if (u > 0) mbar_wait(&empty_bar[s], (u - 1) & 1);       // the MMAs that read this slot are done
        const uint32_t st = ring + (uint32_t)s * T3_STAGE_BYTES;
        const int k0 = it * T3_KC;
        mbar_expect_tx(&full_bar[s], T3_STAGE_BYTES);
        tma_load_2d(st, &tm_ahi, k0, (int)m0, &full_bar[s]);
        tma_load_2d(st + T3_A_BYTES, &tm_alo, k0, (int)m0, &full_bar[s]);
        tma_load_2d(st + 2 * T3_A_BYTES, &tm_bhi, k0, (int)n0, &full_bar[s]);
        tma_load_2d(st + 2 * T3_A_BYTES + T3_B_BYTES, &tm_blo, k0, (int)n0, &full_bar[s]);
      }
    }
    __syncwarp();
    // ===================== epilogue =====================
    if (valid_tile) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const int row = warp * 32 + lane;                         // TMEM lane = accumulator row
      for (int half = 0; half < ncols / 128; ++half) {
        float ss = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 128 + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
              "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float f = __uint_as_float(v[i]);
            ss = fmaf(f, f, ss);
          }
        }
        p.rn_partial[(m0 + row) * p.rn_nt + (n0 / 128 + half)] = (double)ss;
      }
    }
  } else if (warp == 4 && lane == 0) {
    // ===================== MMA issuer (one thread) =====================
    const uint32_t idesc = umma_idesc_tf32(T3_TM, ncols);
    for (int it = 0; it < KT; ++it) {
      const int s = it % T3_STAGES, u = it / T3_STAGES;
      mbar_wait(&full_bar[s], u & 1);
      tc_fence_after();
      const uint32_t a_hi = ring + (uint32_t)s * T3_STAGE_BYTES;
      const uint32_t a_lo = a_hi + T3_A_BYTES, b_hi = a_hi + 2 * T3_A_BYTES, b_lo = b_hi + T3_B_BYTES;
#pragma unroll
      for (int ks = 0; ks < T3_KC / 8; ++ks) {                  // one MMA consumes K = 8 fp32 = 32 B of every row
        const uint32_t off = ks * 32;
        const uint64_t dah = umma_desc_sw64(a_hi + off), dal = umma_desc_sw64(a_lo + off);
        const uint64_t dbh = umma_desc_sw64(b_hi + off), dbl = umma_desc_sw64(b_lo + off);
        tc_mma_tf32(tmem, dah, dbh, idesc, (it | ks) != 0);
        tc_mma_tf32(tmem, dah, dbl, idesc, 1);
        tc_mma_tf32(tmem, dal, dbh, idesc, 1);
      }
      tc_commit(&empty_bar[s]);                                 // arrives when these MMAs have read the stage
    }
    if (valid_tile) tc_commit(&done_bar);                       // accumulator complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(T3_TN));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// fp32 [rows x cols] row-major (ld floats): boxes of box_rows x 16 floats (64 B), SWIZZLE_64B
static int make_tmap(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return ALGP_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {T3_KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? ALGP_OK : ALGP_ERR_INVALID;
}

extern "C" int algp_split_tf32(const double* src, int64_t rows, int64_t cols, int64_t ld, float* hi, float* lo,
                               int64_t ldo, void* stream) {
  if (!src || !hi || !lo || rows < 0 || cols < 0 || (cols & 3) || (ld & 1) || (ldo & 3) || ld < cols || ldo < cols)
    return ALGP_ERR_INVALID;
  if (rows == 0 || cols == 0) return ALGP_OK;
  if (rows > 65535 * 16) {
    // grid.y limit: the matrices on this path are far below it
    return ALGP_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)((cols / 4 + 255) / 256), 1);
  // rows go to grid.y in slabs of 65535
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    grid.y = (unsigned)((rows - r0) < 65535 ? (rows - r0) : 65535);
    split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src + r0 * ld, rows - r0, cols, ld, hi + r0 * ldo, lo + r0 * ldo, ldo);
    ALGP_LAUNCH_CHECK();
  }
  return ALGP_OK;
}

// rn_partial[m][t] (t < npad/128) = sum over column tile t of (K Linv^T)[m][.]^2, split-TF32 on tcgen05
extern "C" int algp_trmm_rt_tf32(const float* Khi, const float* Klo, int64_t mpad, int64_t ldk, const float* Lhi,
                                 const float* Llo, int64_t npad, int64_t ldl, double* rn_partial, void* stream) {
  if (!Khi || !Klo || !Lhi || !Llo || !rn_partial || mpad % ALGP_BLK || npad % ALGP_BLK || ldk < npad || ldl < npad ||
      (ldk & 3) || (ldl & 3))
    return ALGP_ERR_INVALID;
  if (mpad == 0 || npad == 0) return ALGP_OK;
  static AlgpPerDevice configured;
  if (configured.raise(1)) {
    ALGP_CUDA(cudaFuncSetAttribute(trmm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
  }
  if (((uintptr_t)Khi | (uintptr_t)Klo | (uintptr_t)Lhi | (uintptr_t)Llo) & 15) return ALGP_ERR_INVALID;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = make_tmap(&ta_hi, Khi, mpad, npad, ldk, T3_TM))) return rc;
  if ((rc = make_tmap(&ta_lo, Klo, mpad, npad, ldk, T3_TM))) return rc;
  // the L planes are allocated with rows rounded up to 256 (see header), so a half-width last tile never
  // depends on out-of-bounds fill
  const int64_t lrows = (npad + T3_TN - 1) / T3_TN * T3_TN;
  if ((rc = make_tmap(&tb_hi, Lhi, lrows, npad, ldl, T3_TN))) return rc;
  if ((rc = make_tmap(&tb_lo, Llo, lrows, npad, ldl, T3_TN))) return rc;
  T3Args a;
  a.MT = (int)(mpad / T3_TM);
  a.NT = (int)((npad + T3_TN - 1) / T3_TN);
  a.npad = npad;
  a.rn_partial = rn_partial; a.rn_nt = (int)(npad / ALGP_BLK);
  const int groups = (a.MT + 15) / 16;
  const int64_t grid = (int64_t)groups * 16 * a.NT;
  trmm_tf32x3_kernel<<<(unsigned)grid, T3_THREADS, T3_SMEM_BYTES, (cudaStream_t)stream>>>(a, ta_hi, ta_lo, tb_hi, tb_lo);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
