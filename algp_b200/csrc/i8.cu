// fp64-grade predictive variance on the INT8 tensor cores (tcgen05.mma.kind::i8, TMEM s32 accumulators).
//
// var_m = k** - |V_m|^2, V = K(X*,X) L^-T (reference utils.py:300-308) is N^2 M flops and, in fp64, is
// pinned to the DMMA roof (35 TFLOP/s on B200).  tcgen05 has no f64 kind, but it multiplies 8-bit integers
// with EXACT 32-bit accumulation at ~100x the DMMA rate, so the product is evaluated as an error-free
// digit expansion (Ozaki scheme): every row r of an operand is scaled by a power of two 2^-e_r to
// |y| < 1 and written as S signed base-128 digits
//        y = sum_p d_p 2^(-6-7p),   d_p in [-64, 64]            (exact for S = 8: 56 bits)
// so  (A B^T)[m][j] = 2^(eA_m + eB_j - 12) * sum_g 2^(-7g) G_g[m][j],   G_g = sum_{p+q=g} A_p B_q^T.
// Each G_g is an integer GEMM: |d d'| <= 2^12, K <= 2^14 terms, <= 8 (p,q) pairs -> < 2^29, no overflow,
// no rounding.  Groups g >= S are dropped (relative 2^(-7S) of the row scales), the rest is summed in fp64
// in the epilogue, low order first.  S = 7 (28 GEMMs) or 8 (36 GEMMs) meets the fp64 tier.
//
//   split_i8_kernel   fp64 matrix -> S int8 digit planes + the row scale 2^e_r (one CTA per 8 rows; the row
//                     maximum fixes e_r, the second read of the rows hits L1/L2).  The planes are written
//                     in the MMA's own operand layout, tile by tile:
//                         [row tile][32-byte k chunk][plane][row group of 8][k half][8 rows][16 B]
//                     (the K-major no-swizzle canonical layout: 8 x 16 B core matrices, LBO 128 B between
//                     the k halves, SBO 256 B between row groups), so one k-stage of one operand -- all
//                     S planes -- is ONE contiguous block of global memory.
//   trmm_i8_kernel    one CTA per 128 x 64 tile of V.  All S group accumulators of the tile live in TMEM
//                     at once (S x 64 columns: the whole 512-column TMEM for S = 8), so the operand
//                     digits stream through shared memory exactly once.  One producer thread issues two
//                     bulk copies (cp.async.bulk, 32 KB + 16 KB contiguous for S = 8) per 32-byte k-stage
//                     -- full-line L2 requests; a tensor-map box of 32-byte rows saturates the SM's
//                     request path to the crossbar at one 32-byte sector per clock (ncu, r01) --; one
//                     thread issues the stage's MMAs (M128 K32, N = 64..256: the B planes of a stage are one
//                     contiguous operand, so one instruction covers up to 4 plane products) and commits the
//                     stage's `empty` mbarrier; 4-stage ring.  Epilogue: tcgen05.ld, Horner in fp64 over the
//                     groups, column scale, square, row sum -- V is never written.  k-range stops at the
//                     tile's last column (Linv is lower triangular).
#include "common.cuh"
#include "i8.cuh"

#define I8_MAX_PLAN 1024                // k chunks per tile: K <= 32768
#define I8_THREADS 224                  // 4 epilogue warps, warps 4 and 6: MMA issuers (even / odd chunks), warp 5: bulk-copy producer

namespace {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit_to(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
// K-major no-swizzle (INTERLEAVE) operand tile: 8 x 16 B core matrices; LBO = 128 B between the two 16-byte
// k halves of an MMA, SBO = 256 B between 8-row groups (cute::UMMA::SmemDescriptor: version 1 at bit 46,
// layout_type 0)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// The low word holds the start address (>> 4, 14 bits) and the LBO, the high word the SBO and the version: operand
// offsets only ever touch the low word (shared memory is < 256 KB, no carry), so descriptors are handled as 32-bit
// values -- 64-bit adds with carry are what kept them out of the uniform datapath.
__device__ __forceinline__ uint32_t desc_kmajor_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((128u >> 4) << 16); }
#define I8_DESC_HI ((256u >> 4) | (1u << 14))
// cute::UMMA::InstrDescriptor: c_format S32 (2) [4,6), a/b_format signed 8-bit (1) [7,10) / [10,13), K-major,
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256)
split_i8_kernel(const double* __restrict__ src, int64_t cols, int64_t ld, int tile_rows, int8_t* __restrict__ planes,
                double* __restrict__ row_scale, unsigned int* __restrict__ mask_words, int64_t mask_ld) {
  __shared__ int exps[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * 8;
  {
    // warp w: maximum of row row0 + w
    const double* x = src + (row0 + warp) * ld;
    double mx = 0.0;
    for (int64_t c = (int64_t)lane * 2; c < cols; c += 64) {
      const double2 v = *reinterpret_cast<const double2*>(x + c);
      mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0) (void)frexp(mx, &e);              // mx = f 2^e, f in [0.5, 1): |x| 2^-e < 1
    if (lane == 0) {
      exps[warp] = e;
      row_scale[row0 + warp] = ldexp(1.0, e);
    }
  }
  __syncthreads();
  const int rr = tid & 7;                           // row inside the 8-row group
  const int64_t row = row0 + rr;
  const double* x = src + row * ld;
  const double sc = ldexp(1.0, 6 - exps[rr]);       // t0 = 64 y
  const int64_t tile = row / tile_rows;
  const int64_t plane_bytes = (int64_t)tile_rows * I8_KC;
  const int64_t kchunks = cols / I8_KC, pieces = cols / 16;
  // byte offset of (row, k = 0, plane 0) inside its tile's first k chunk
  const int64_t row_off = ((row % tile_rows) / 8) * 256 + rr * 16;
  // a warp covers 4 consecutive 16-column pieces (= 2 k chunks) of its 8 rows per iteration
  for (int64_t j0 = 4 * warp; j0 < pieces; j0 += 32) {
    const int64_t j = j0 + (lane >> 3);
    const bool valid = j < pieces;
    unsigned int occ = 0;                            // bit p: this thread wrote a non-zero digit into plane p
    if (valid) {
      double t[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(x + j * 16 + 2 * i);
        t[2 * i] = v.x * sc;
        t[2 * i + 1] = v.y * sc;
      }
      int8_t* out = planes + ((tile * kchunks + (j >> 1)) * S) * plane_bytes + row_off + (j & 1) * 128;
#pragma unroll
      for (int p = 0; p < S; ++p) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pack = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = q * 4 + b;
            const double d = rint(t[i]);              // |t| <= 64: the digit, exact
            t[i] = (t[i] - d) * 128.0;                // exact remainder, next digit's scale
            pack |= ((uint32_t)(__double2int_rn(d)) & 0xFFu) << (8 * b);
          }
          w[q] = pack;
        }
        if ((w[0] | w[1] | w[2] | w[3]) != 0u) occ |= 1u << p;
        *reinterpret_cast<uint4*>(out + (int64_t)p * plane_bytes) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (mask_words) {
      // plane occupancy of (row tile, k chunk): lanes 0-15 hold chunk j0/2, lanes 16-31 chunk j0/2 + 1
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) occ |= __shfl_xor_sync(0xffffffffu, occ, o);
      if ((lane & 15) == 0 && occ != 0u) {
        const int64_t kc = (j0 >> 1) + (lane >> 4);
        if (kc < kchunks) {
          const int64_t byte = tile * mask_ld + kc;
          atomicOr(mask_words + (byte >> 2), occ << (8 * (int)(byte & 3)));
        }
      }
    }
  }
}

template <int S>
struct I8Cfg {
  static constexpr int A_PLANE = I8_TM * I8_KC;                       // 4 KB
  static constexpr int B_PLANE = I8_TN * I8_KC;                       // 2 KB
  static constexpr int A_BYTES = S * A_PLANE;
  static constexpr int B_BYTES = S * B_PLANE;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;               // S x 6 KB
  // EVEN by construction: the two MMA issuer warps take the plan entries alternately, so with an even ring every stage
  // belongs to one issuer.  With an odd ring a stage alternates between them, an issuer can get two phases ahead of
  // the other on the same `full` barrier, and a parity wait then passes on the stale phase (observed: launch failure /
  // hang with 3 and 5 stages).
#ifndef I8_STAGES_OVERRIDE
  // as deep as 200 KB allow, at most 8: the 4-plane variance GEMM takes 27.8 / 18.5 / 17.9 / 16.1 ms with 2 / 4 / 6 / 8 stages
  static constexpr int SMEM_BUDGET = (S <= 4) ? 104 * 1024 : 200 * 1024;    // S <= 4: two CTAs per SM
  static constexpr int STAGES_RAW = (S >= 8) ? 4 : (SMEM_BUDGET / STAGE_BYTES > 8 ? 8 : SMEM_BUDGET / STAGE_BYTES);
#else
  static constexpr int STAGES_RAW = I8_STAGES_OVERRIDE;   // pipeline-depth experiments
#endif
  static constexpr int STAGES = STAGES_RAW & ~1;
  static_assert(STAGES >= 2, "ring too shallow");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (S * I8_TN > 256) ? 512 : ((S * I8_TN > 128) ? 256 : 128);
};

// One lane's view of one k chunk: 0 if the chunk needs nothing, else bit 31 | pmin | qmin << 4 | qmax << 8 with
// pmin / qmin the lowest occupied digit plane of the A / B tile chunk and qmax the highest occupied B plane.
// Plane p of A meets plane q of B only if p + q < S, so a chunk is needed iff pmin + qmin < S; it then needs the A
// planes [pmin, S-1-qmin] and, for A plane p, the B planes [qmin, min(S-p, qmax+1)).  Empty planes INSIDE those
// ranges are multiplied as the zeros they are: the ranges keep the per-chunk schedule to a handful of instructions.
template <int S>
__device__ __forceinline__ uint32_t plan_word(const uint8_t* am, const uint8_t* bm, int kc, int kend) {
  if (kc >= kend) return 0u;
  constexpr uint32_t ALLP = (1u << S) - 1u;
  const uint32_t a = (am ? (uint32_t)am[kc] : 0xffu) & ALLP, b = (bm ? (uint32_t)bm[kc] : 0xffu) & ALLP;
  if (a == 0u || b == 0u) return 0u;
  const uint32_t pmin = (uint32_t)__ffs(a) - 1u, qmin = (uint32_t)__ffs(b) - 1u, qmax = 31u - (uint32_t)__clz(b);
  if (pmin + qmin > (uint32_t)(S - 1)) return 0u;
  return 0x80000000u | pmin | (qmin << 4) | (qmax << 8);
}
// D[tmem] += A[smem] B[smem]^T, signed 8-bit operands, s32 accumulators (zero-initialised: always accumulating)
__device__ __forceinline__ void mma_i8_acc(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 1, 1;\n\t"
      "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(idesc), "r"(I8_DESC_HI)
      : "memory");
}

// The MMAs of a sparse chunk with L A planes and W reachable B planes (W <= L), every offset and every instruction
// descriptor a compile-time constant relative to the chunk's three bases: the i-th A plane meets min(L - i, W) B planes,
// issued as one MMA of up to four planes (N = 64 per plane) plus a second one when more than four remain.
template <int S, int L, int W>
__device__ __forceinline__ void issue_sparse_lw(uint32_t td, uint32_t da, uint32_t db) {
  using C = I8Cfg<S>;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const int n = (L - i < W) ? L - i : W;
    const int c0 = n < 4 ? n : 4;
    mma_i8_acc(td + (uint32_t)(i * I8_TN), da + (uint32_t)(i * (C::A_PLANE >> 4)), db, idesc_i8(I8_TM, c0 * I8_TN));
    if (n > 4)
      mma_i8_acc(td + (uint32_t)((i + 4) * I8_TN), da + (uint32_t)(i * (C::A_PLANE >> 4)), db + (uint32_t)((4 * C::B_PLANE) >> 4),
                 idesc_i8(I8_TM, (n - 4) * I8_TN));
  }
}
// one switch over the S(S+1)/2 (L, W) pairs: a single indexed branch instead of a chain of compares
template <int S>
__device__ __forceinline__ void issue_sparse(int L, int W, uint32_t td, uint32_t da, uint32_t db) {
#define I8_CASE(l, w) case (l) * 8 + (w): if constexpr ((l) <= S && (w) <= (l)) issue_sparse_lw<S, (l), (w)>(td, da, db); break;
#define I8_ROW(l) I8_CASE(l, 1) I8_CASE(l, 2) I8_CASE(l, 3) I8_CASE(l, 4) I8_CASE(l, 5) I8_CASE(l, 6) I8_CASE(l, 7) I8_CASE(l, 8)
  switch (L * 8 + W) {
    I8_ROW(1) I8_ROW(2) I8_ROW(3) I8_ROW(4) I8_ROW(5) I8_ROW(6) I8_ROW(7) I8_ROW(8)
    default: break;
  }
#undef I8_ROW
#undef I8_CASE
}

// MODE 0: row sums of squares per 64-column tile (variance path, nothing else is written)
// MODE 1: C = alpha A B^T + beta C (optionally stored transposed)
// S <= 4: the group accumulators take 256 TMEM columns and a 4-stage ring 96 KB, so TWO CTAs fit an SM and one tile's
// prologue / epilogue (TMEM zero-fill, plan decode, pipeline ramp, tcgen05.ld + Horner: about half of a 4-plane tile's
// time) overlaps the other tile's MMAs.
template <int S, int MODE>
__global__ void __launch_bounds__(I8_THREADS, (S <= 4) ? 2 : 1)
gemm_i8_kernel(const I8Gemm p) {
  using C = I8Cfg<S>;
  extern __shared__ unsigned char i8_smem_raw[];
  __shared__ uint64_t full_bar[C::STAGES], empty_bar[C::STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ double sb_s[I8_TN];
  __shared__ uint32_t plan_s[I8_MAX_PLAN];   // the k chunks of this tile that need work, in order: plan word | chunk << 12
  __shared__ int plan_n;
  // the warp index is broadcast from lane 0 so that the compiler can prove the role branches warp-uniform (the
  // role bodies then keep their addresses and descriptors in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t ring = (s_u32(i8_smem_raw) + 1023u) & ~1023u;

  // tile map: groups of 16 m-tiles sweep the n-tiles together, so concurrent CTAs stream the same k-window of
  // the A digits and the same B rows through L2; the order flags put the tiles with the longest k-range first
  constexpr int GM = 16;
  const int per_group = GM * p.NT, groups = (p.MT + GM - 1) / GM;
  int mg = blockIdx.x / per_group;
  const int rem = blockIdx.x % per_group;
  if (p.mt_desc) mg = groups - 1 - mg;
  const int nt = p.nt_desc ? p.NT - 1 - rem / GM : rem / GM;
  const int mt = mg * GM + rem % GM;
  const int m0 = mt * I8_TM, n0 = nt * I8_TN;
  bool valid_tile = mt < p.MT;                       // block-uniform
  if (p.lower_only && n0 >= m0 + I8_TM) valid_tile = false;
  int kbeg = 0, kend = p.kchunks;
  if (p.kbeg_rule == I8_KB_NT) kbeg = n0 / I8_KC;                                  // B[j][k] = 0 for k < j
  if (p.kend_rule == I8_KE_NT) kend = min(kend, (n0 + I8_TN) / I8_KC);             // B[j][k] = 0 for k > j
  if (p.kend_rule == I8_KE_MT) kend = min(kend, (m0 + I8_TM) / I8_KC);             // A[i][k] = 0 for k > i
  const int KT = valid_tile ? max(kend - kbeg, 0) : 0;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbarrier_init(&full_bar[s], 1);
      mbarrier_init(&empty_bar[s], 1);
    }
    mbarrier_init(&done_bar, 2);                     // one arrival per MMA issuer warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < I8_TN && valid_tile) sb_s[tid] = p.scale_b[n0 + tid];
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  // the accumulators start from zero, so every MMA accumulates: with occupancy masks there is no fixed "first"
  // product per group
  if (warp < 4 && KT > 0) {
#pragma unroll 1
    for (int c0 = 0; c0 < S * I8_TN; c0 += 16) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
          ::"r"(taddr), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else if (warp == 5) {
    // plane-occupancy masks of the two operands' tiles (one byte per 32-byte k chunk, bit p = plane p holds a
    // non-zero digit): chunks and planes that are all zero are neither loaded nor multiplied.  While warps 0-3 zero
    // the accumulators, this warp decodes the whole k-range 32 chunks at a time and compacts the chunks that need
    // work into plan_s: the role loops below are bound by their own instruction count (one warp alone issues one
    // instruction per ~4.4 cycles, profiles/r01_i8_experiments.log), so everything that can be done here is.
    const uint8_t* am = p.a_mask ? p.a_mask + (int64_t)mt * p.mask_ld : nullptr;
    const uint8_t* bm = p.b_mask ? p.b_mask + (int64_t)nt * p.mask_ld : nullptr;
    int cnt = 0;
    if (KT > 0) {
#pragma unroll 4
      for (int kb = kbeg; kb < kend; kb += 32) {
        const uint32_t w = plan_word<S>(am, bm, kb + lane, kend);
        const uint32_t bits = __ballot_sync(0xffffffffu, w != 0u);
        if (w != 0u) plan_s[cnt + __popc(bits & ((1u << lane) - 1u))] = 0x80000000u | (w & 0xfffu) | ((uint32_t)(kb + lane) << 12);
        cnt += __popc(bits);
      }
    }
    if (lane == 0) plan_n = cnt;
  }
  fence_before();
  __syncthreads();
  fence_after();
  // warp-uniform by construction (broadcast from lane 0), so the role loops stay in uniform control flow
  const int n_plan = __shfl_sync(0xffffffffu, plan_n, 0);

  if (warp < 4) {
    // ===================== epilogue =====================
    if (valid_tile) {
      const int row = warp * 32 + lane;                          // TMEM lane = accumulator row
      const double sa = p.scale_a[m0 + row] * (1.0 / 4096.0);   // 2^(eA_m - 12)
      if (KT > 0) {
        // sleep between polls: four spinning warps would take issue slots from the two single-lane roles
        while (!mbarrier_test(&done_bar, 0)) __nanosleep(256);
        fence_after();
      }
      double ss = 0.0;
#pragma unroll 1
      for (int c0 = 0; c0 < I8_TN; c0 += 16) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0;
        if (KT > 0) {
          // all S groups of these 16 columns are requested before the one wait: the loads are latency-bound, and one
          // round trip per 16 columns instead of one per group takes the epilogue from ~6.4k to ~3k cycles per tile
          uint32_t r[S][16];
#pragma unroll
          for (int g = 0; g < S; ++g) {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * I8_TN + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[g][0]), "=r"(r[g][1]), "=r"(r[g][2]), "=r"(r[g][3]), "=r"(r[g][4]), "=r"(r[g][5]), "=r"(r[g][6]),
                  "=r"(r[g][7]), "=r"(r[g][8]), "=r"(r[g][9]), "=r"(r[g][10]), "=r"(r[g][11]), "=r"(r[g][12]), "=r"(r[g][13]),
                  "=r"(r[g][14]), "=r"(r[g][15])
                : "r"(taddr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int g = S - 1; g >= 0; --g) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fma(v[i], 0.0078125, (double)(int)r[g][i]);   // Horner in 2^-7
          }
        }
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const double t = v[i] * sb_s[c0 + i];
            ss = fma(t, t, ss);
          }
        } else {
          const double as = p.alpha * sa;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] *= sb_s[c0 + i];
            ss = fma(v[i], v[i], ss);                 // row norms of A B^T ride along when asked for
          }
          if (!p.transposed) {
            // thread = row: 16 consecutive doubles (one 128-byte line) per chunk
            double* crow = p.C + (int64_t)(m0 + row) * p.ldc + n0 + c0;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              double2 o = make_double2(v[i] * as, v[i + 1] * as);
              if (p.beta != 0.0) {
                const double2 old = *reinterpret_cast<const double2*>(crow + i);
                o.x = fma(p.beta, old.x, o.x);
                o.y = fma(p.beta, old.y, o.y);
              }
              *reinterpret_cast<double2*>(crow + i) = o;
            }
          } else {
            // C^T: the 32 lanes of a warp are 32 consecutive elements of one row of the transposed matrix
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              double* ct = p.C + (int64_t)(n0 + c0 + i) * p.ldc + m0 + row;
              double o = v[i] * as;
              if (p.beta != 0.0) o = fma(p.beta, *ct, o);
              *ct = o;
            }
          }
        }
      }
      if (MODE == 0 || p.rn_partial) p.rn_partial[(int64_t)(m0 + row) * p.rn_nt + nt] = ss * sa * sa;
    }
  } else if (warp == 5 && KT > 0) {
    // ===================== bulk-copy producer (warp 5) =====================
    // Warp-uniform control flow: the plan word comes from shared memory through a lane-0 broadcast (a result the
    // compiler knows to be uniform), fetched one chunk ahead; the loop is unrolled over the ring's stages, so slot
    // addresses and barrier addresses are constants and everything else lives in uniform registers.
    const int8_t* a_src = p.a_tiles + (int64_t)mt * p.kchunks * C::A_BYTES;
    const int8_t* b_src = p.b_tiles + (int64_t)nt * p.kchunks * C::B_BYTES;
    uint32_t ws_next = __shfl_sync(0xffffffffu, plan_s[0], 0);
    uint32_t par = 1;                                              // parity of the empty phase of the previous round
    for (int it0 = 0; it0 < n_plan; it0 += C::STAGES) {
#pragma unroll
      for (int s = 0; s < C::STAGES; ++s) {
        const int it = it0 + s;
        if (it < n_plan) {
          const uint32_t ws = ws_next;
          ws_next = __shfl_sync(0xffffffffu, plan_s[(it + 1) & (I8_MAX_PLAN - 1)], 0);
          // A planes [pmin, S-1-qmin] and B planes [qmin, min(S-pmin, qmax+1)): both ranges are contiguous in the
          // tiled digit layout, one bulk copy each
          const int pmin = ws & 0xf, qmin = (ws >> 4) & 0xf, qmax = (ws >> 8) & 0xf;
          const int kc = (int)((ws >> 12) & (I8_MAX_PLAN - 1));
          const int nA = S - qmin - pmin, nB = min(S - pmin, qmax + 1) - qmin;
          if (it0 > 0) mbarrier_wait(&empty_bar[s], par);          // the MMAs that read this slot are done
          if (elect_one()) {
            const uint32_t st = ring + (uint32_t)s * C::STAGE_BYTES;
            expect_tx(&full_bar[s], (uint32_t)(nA * C::A_PLANE + nB * C::B_PLANE));
            bulk_load(st + pmin * C::A_PLANE, a_src + (int64_t)kc * C::A_BYTES + pmin * C::A_PLANE,
                      (uint32_t)(nA * C::A_PLANE), &full_bar[s]);
            bulk_load(st + C::A_BYTES + qmin * C::B_PLANE, b_src + (int64_t)kc * C::B_BYTES + qmin * C::B_PLANE,
                      (uint32_t)(nB * C::B_PLANE), &full_bar[s]);
          }
        }
      }
      par ^= 1u;
    }
  } else if ((warp == 4 || warp == 6) && KT > 0) {
    // ===================== MMA issuers (warps 4 and 6; one elected lane each issues) =====================
    // TWO issuers take the surviving chunks alternately (warp 4: even plan entries, warp 6: odd).  One issuer spends
    // ~770 cycles per chunk on its control chain (full-barrier wait, plan look-ahead, elect, UTCIMMA dispatch, commit:
    // profiles/r01_i8_phases.log) against ~380 cycles of tensor work, so the tensor pipe idled half the time; the
    // MMAs of both warps enter the SM's one tensor pipe, which executes them in arrival order, and integer
    // accumulation into a group's TMEM columns is exact and commutative, so the interleaving of the two streams does
    // not change a bit of the result.  Each issuer commits its own chunks' `empty` barriers and arrives on done_bar
    // (count 2) after its last MMA.
    const int iss = __shfl_sync(0xffffffffu, (warp == 6) ? 1 : 0, 0);
    // The B digit planes of a stage are contiguous in shared memory ([plane][64 rows][32 B]), i.e. ONE K-major
    // operand of (S - p) x 64 rows, and group g = p + q lives at TMEM columns g x 64: a single MMA of A_p against
    // planes q0..q0+c-1 (N = 64c <= 256) lands every product in its own group.  S(S+1)/2 plane products become
    // ~S(S+1)/8 + S/2 instructions and A_p is read from shared memory once per <= 4 products instead of once each.
    // Same uniform control flow as the producer; the MMAs of a sparse chunk come from a switch over its (L, W) shape
    // with every offset and descriptor an immediate (issue_sparse).
    uint32_t ws_next = __shfl_sync(0xffffffffu, plan_s[iss], 0);
    const uint32_t da_ring = desc_kmajor_lo(ring);
    for (int it = iss; it < n_plan; it += 2) {
      const uint32_t ws = ws_next;
      ws_next = __shfl_sync(0xffffffffu, plan_s[(it + 2) & (I8_MAX_PLAN - 1)], 0);
      const int s = it % C::STAGES, u = it / C::STAGES;
      const uint32_t da0 = da_ring + (uint32_t)s * (uint32_t)(C::STAGE_BYTES >> 4), db0 = da0 + (uint32_t)(C::A_BYTES >> 4);
      const int pmin = ws & 0xf, qmin = (ws >> 4) & 0xf, qmax = (ws >> 8) & 0xf;
      const int L = S - pmin - qmin;
      int W = qmax + 1 - qmin;
      W = W < L ? W : L;
      mbarrier_wait(&full_bar[s], u & 1);
      fence_after();
      if (elect_one()) {
        if ((ws & 0xfffu) == (uint32_t)((S - 1) << 8)) {
          // dense chunk (pmin = qmin = 0, qmax = S-1): the fixed schedule, every operand a compile-time offset
#pragma unroll
          for (int pa = 0; pa < S; ++pa) {
#pragma unroll
            for (int q0 = 0; q0 < S - pa; q0 += 4) {
              const int cnt = (S - pa - q0) < 4 ? (S - pa - q0) : 4;
              mma_i8_acc(tmem + (uint32_t)((pa + q0) * I8_TN), da0 + (uint32_t)((pa * C::A_PLANE) >> 4),
                         db0 + (uint32_t)((q0 * C::B_PLANE) >> 4), idesc_i8(I8_TM, cnt * I8_TN));
            }
          }
        } else {
          // sparse chunk: the L = S - pmin - qmin A planes from pmin on against the W reachable B planes from qmin on
          issue_sparse<S>(L, W, tmem + (uint32_t)((pmin + qmin) * I8_TN), da0 + (uint32_t)(pmin * (C::A_PLANE >> 4)),
                          db0 + (uint32_t)(qmin * (C::B_PLANE >> 4)));
        }
        commit_to(&empty_bar[s]);                               // arrives when these MMAs have read the stage
      }
    }
    if (elect_one()) {
      if (n_plan > iss) commit_to(&done_bar);                    // this issuer's MMAs complete
      else mbarrier_arrive(&done_bar);                           // it had no chunk (all skipped, or a single one)
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS));
  }
}

template <int S>
int launch_split(const double* src, int64_t rows, int64_t cols, int64_t ld, int tile_rows, int8_t* planes, double* row_scale,
                 uint8_t* mask, cudaStream_t st) {
  const int64_t mask_ld = i8_mask_ld(cols);
  if (mask) ALGP_CUDA(cudaMemsetAsync(mask, 0, (size_t)(rows / tile_rows) * mask_ld, st));
  split_i8_kernel<S><<<(unsigned)(rows / 8), 256, 0, st>>>(src, cols, ld, tile_rows, planes, row_scale,
                                                           reinterpret_cast<unsigned int*>(mask), mask_ld);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

template <int S, int MODE>
int launch_gemm(const I8Gemm& a, cudaStream_t st) {
  using C = I8Cfg<S>;
  static AlgpPerDevice configured;
  if (configured.raise(1)) {
    ALGP_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const int groups = (a.MT + 15) / 16;
  const int64_t grid = (int64_t)groups * 16 * a.NT;
  gemm_i8_kernel<S, MODE><<<(unsigned)grid, I8_THREADS, C::SMEM_BYTES, st>>>(a);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

}  // namespace

int i8_split(const double* src, int64_t rows, int64_t cols, int64_t ld, int nslices, int tile_rows, int8_t* planes,
             double* row_scale, uint8_t* mask, cudaStream_t st) {
  if (mask && ((uintptr_t)mask & 7)) return ALGP_ERR_INVALID;
  if (!src || !planes || !row_scale || rows < 0 || cols < 0 || (cols % I8_KC) || (ld & 1) || ld < cols ||
      (tile_rows != I8_TM && tile_rows != I8_TN) || rows % tile_rows || nslices < 2 || nslices > I8_MAX_S)
    return ALGP_ERR_INVALID;
  if (((uintptr_t)src | (uintptr_t)planes) & 15) return ALGP_ERR_INVALID;
  if (rows == 0 || cols == 0) return ALGP_OK;
  switch (nslices) {
    case 2: return launch_split<2>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    case 3: return launch_split<3>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    case 4: return launch_split<4>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    case 5: return launch_split<5>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    case 6: return launch_split<6>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    case 7: return launch_split<7>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
    default: return launch_split<8>(src, rows, cols, ld, tile_rows, planes, row_scale, mask, st);
  }
}

// C (or row-norm partials) from digit tiles; see I8Gemm in i8.cuh
int i8_gemm(const I8Gemm& a, int nslices, cudaStream_t st) {
  if (!a.a_tiles || !a.b_tiles || !a.scale_a || !a.scale_b || a.MT < 0 || a.NT < 0 || a.kchunks < 0 || nslices < 2 ||
      nslices > I8_MAX_S || a.kchunks > 32768 / I8_KC)
    return ALGP_ERR_INVALID;           // K <= 2^15 keeps every group sum below 2^31 (8 pairs x 2^15 x 2^12 = 2^30)
  if (a.MT == 0 || a.NT == 0) return ALGP_OK;
  if (((uintptr_t)a.a_tiles | (uintptr_t)a.b_tiles) & 15) return ALGP_ERR_INVALID;
  if (((uintptr_t)a.a_mask | (uintptr_t)a.b_mask) & 7) return ALGP_ERR_INVALID;
  if ((a.a_mask || a.b_mask) && a.mask_ld < (a.kchunks + 7) / 8 * 8) return ALGP_ERR_INVALID;
  if (a.C) {
    if (a.ldc < 2 || (a.ldc & 1) || ((uintptr_t)a.C & 15)) return ALGP_ERR_INVALID;
    switch (nslices) {
      case 2: return launch_gemm<2, 1>(a, st);
      case 3: return launch_gemm<3, 1>(a, st);
      case 4: return launch_gemm<4, 1>(a, st);
      case 5: return launch_gemm<5, 1>(a, st);
      case 6: return launch_gemm<6, 1>(a, st);
      case 7: return launch_gemm<7, 1>(a, st);
      default: return launch_gemm<8, 1>(a, st);
    }
  }
  if (!a.rn_partial) return ALGP_ERR_INVALID;
  switch (nslices) {
    case 2: return launch_gemm<2, 0>(a, st);
    case 3: return launch_gemm<3, 0>(a, st);
    case 4: return launch_gemm<4, 0>(a, st);
    case 5: return launch_gemm<5, 0>(a, st);
    case 6: return launch_gemm<6, 0>(a, st);
    case 7: return launch_gemm<7, 0>(a, st);
    default: return launch_gemm<8, 0>(a, st);
  }
}

extern "C" int algp_split_i8(const double* src, int64_t rows, int64_t cols, int64_t ld, int nslices, int tile_rows,
                             int8_t* planes, double* row_scale, uint8_t* mask, void* stream) {
  return i8_split(src, rows, cols, ld, nslices, tile_rows, planes, row_scale, mask, (cudaStream_t)stream);
}

extern "C" int64_t algp_i8_mask_bytes(int64_t rows, int64_t cols, int tile_rows) {
  if (rows < 0 || cols < 0 || (tile_rows != I8_TM && tile_rows != I8_TN)) return 0;
  return (rows + tile_rows - 1) / tile_rows * i8_mask_ld(cols) + 8;
}

// rn_partial[m][t] (t < npad/64) = sum over 64-column tile t of (K Linv^T)[m][.]^2 from the digit tiles
extern "C" int algp_trmm_rt_i8(const int8_t* Kt, const double* Kscale, const uint8_t* Kmask, int64_t mpad, const int8_t* Lt,
                               const double* Lscale, const uint8_t* Lmask, int64_t npad, int nslices, double* rn_partial,
                               void* stream) {
  if (!rn_partial || mpad < 0 || npad < 0 || mpad % ALGP_BLK || npad % ALGP_BLK) return ALGP_ERR_INVALID;
  I8Gemm a = i8_gemm_default();
  a.MT = (int)(mpad / I8_TM);
  a.NT = (int)(npad / I8_TN);
  a.a_tiles = Kt;
  a.b_tiles = Lt;
  a.kchunks = (int)(npad / I8_KC);
  a.scale_a = Kscale;
  a.scale_b = Lscale;
  a.kend_rule = I8_KE_NT;              // Linv is lower triangular
  a.nt_desc = 1;
  a.rn_partial = rn_partial;
  a.rn_nt = a.NT;
  a.a_mask = Kmask; a.b_mask = Lmask; a.mask_ld = i8_mask_ld(npad);
  return i8_gemm(a, nslices, (cudaStream_t)stream);
}

// C [mpad x npad] = alpha A B^T + beta C from the digit tiles of A [mpad x kpad] (tile_rows 128) and B [npad x kpad]
// (tile_rows 64); transposed != 0 stores C^T ([npad x mpad], ldc its row stride)
// V = K Linv^T stored [mpad x npad] (row stride ldv) AND its row-norm partials rn_partial[m][t] (t < npad/64): the
// W^T build of the posterior state (Sigma_{:,B} L^-T) on the INT8 tensor cores
extern "C" int algp_trmm_rt_store_i8(const int8_t* Kt, const double* Kscale, const uint8_t* Kmask, int64_t mpad,
                                     const int8_t* Lt, const double* Lscale, const uint8_t* Lmask, int64_t npad, int nslices,
                                     double* V, int64_t ldv, double* rn_partial, void* stream) {
  if (!V || mpad < 0 || npad < 0 || mpad % ALGP_BLK || npad % ALGP_BLK || ldv < npad) return ALGP_ERR_INVALID;
  I8Gemm a = i8_gemm_default();
  a.MT = (int)(mpad / I8_TM);
  a.NT = (int)(npad / I8_TN);
  a.a_tiles = Kt;
  a.b_tiles = Lt;
  a.kchunks = (int)(npad / I8_KC);
  a.scale_a = Kscale;
  a.scale_b = Lscale;
  a.kend_rule = I8_KE_NT;              // Linv is lower triangular
  a.nt_desc = 1;
  a.C = V;
  a.ldc = ldv;
  a.rn_partial = rn_partial;
  a.rn_nt = a.NT;
  a.a_mask = Kmask; a.b_mask = Lmask; a.mask_ld = i8_mask_ld(npad);
  return i8_gemm(a, nslices, (cudaStream_t)stream);
}

extern "C" int algp_gemm_nt_i8(const int8_t* At, const double* Ascale, int64_t mpad, const int8_t* Bt, const double* Bscale,
                               int64_t npad, int64_t kpad, int nslices, double alpha, double beta, double* C, int64_t ldc,
                               int transposed, int lower_only, void* stream) {
  if (!C || mpad < 0 || npad < 0 || kpad < 0 || mpad % I8_TM || npad % I8_TN || kpad % I8_KC ||
      ldc < (transposed ? mpad : npad))
    return ALGP_ERR_INVALID;
  I8Gemm a = i8_gemm_default();
  a.MT = (int)(mpad / I8_TM);
  a.NT = (int)(npad / I8_TN);
  a.a_tiles = At;
  a.b_tiles = Bt;
  a.kchunks = (int)(kpad / I8_KC);
  a.scale_a = Ascale;
  a.scale_b = Bscale;
  a.C = C;
  a.ldc = ldc;
  a.alpha = alpha;
  a.beta = beta;
  a.transposed = transposed ? 1 : 0;
  a.lower_only = lower_only ? 1 : 0;
  return i8_gemm(a, nslices, (cudaStream_t)stream);
}
