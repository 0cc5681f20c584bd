// EXPERIMENT (round 2), not part of libalgp_b200.so: a blocked variant of the diagonal-block kernel of algp_potrf
// (csrc/chol.cu, potf2inv_rank_kernel).  Correct (same L and L^-1 as the shipped kernel to 7e-15 / 4e-13) and not
// faster: 47.6 us per 128 x 128 block in this form, 43.5 us with all eight warps eliminating redundantly, against
// 41.4 us for the shipped rank-2 kernel.  Why, with the phase timings: profiles/r02_potf2_blocked.log.
// Built only by scripts/micro/potf2_phases.cu.
#pragma once
#include "../../algp_b200/csrc/common.cuh"

// ---------------------------------------------------------------------------
// Blocked variant: 8 panels of 16 columns, TWO barriers per panel instead of one per column (or per R columns).
//
// Per panel p (columns 16p .. 16p+15):
//   publish    every thread writes its piece of the panel (patch column p of acc, rows >= 16p; the diagonal 16x16
//              block symmetrically) and of the panel's rows of Y (patch row p of yac) into shared memory;
//   eliminate  a 16-column LDL^T elimination in registers with no CTA barrier and no shared-memory round trip on the
//              column-to-column chain.  The panel is 16 rows of the diagonal block plus 128 "vectors": the 112-16p
//              rows below the block and the 16(p+1) non-zero columns of the panel's rows of Y.  Row r below the block
//              and column y of Y obey the same recurrence (L_d x = b), so one instruction stream serves both, and
//              the panel solve X = B L_d^-T and Y_p <- L_d^-1 Y_p are the same elimination applied to more lanes.
//              Warp 0 (pb_pivot_warp) holds the diagonal block and 16 vectors and owns the chain; it publishes every
//              eliminated column (pivot row, -1/piv, 1/sqrt(piv)) in a shared table and releases a counter; three
//              and a half other warps (pb_trail), one per scheduler, eliminate the remaining 112 vectors from the
//              table as the columns appear.  Results go back to shared memory and straight to global memory (they
//              are final entries of L and L^-1);
//   update     rank-16 update of the live register patches from shared memory (16-byte loads over k, conflict-free
//              with an 18-double row pitch): acc[i][j] -= X_i X_j^T (i >= j > p), yac[i][j] -= X_i Y_p[:, j] (i > p >= j).
// The shared panels are double-buffered by panel parity, so the update of panel p and the publish of panel p+1 need
// no barrier between them.
// ---------------------------------------------------------------------------
#define PB_LD 18
#ifdef PB_PROF
__device__ long long g_pb_prof[64];
#define PB_MARK(slot) do { if (tid == 0) g_pb_prof[slot] = clock64(); } while (0)
#else
#define PB_MARK(slot) do { } while (0)
#endif
#define PB_SMEM_BYTES ((4 * 128 * PB_LD + 16 * PB_LD + 2) * 8)      // two (X, Y^T) panel pairs + the pivot-row table + the counter

__device__ __forceinline__ double pb_rcp(double x) {
  // MUFU.RCP64H (~20 bits) + one cubic Newton step: relative error ~2^-60 before rounding, i.e. within an ulp or two
  // for the positive normal pivots of this path, and three dependent DFMAs on the pivot chain instead of the six of
  // the IEEE division sequence
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double e2 = fma(e, e, e);
  return fma(r, e2, r);
}

__device__ __forceinline__ double pb_rsqrt(double x) {
  // MUFU.RSQ64H (~20 bits) + one third-order step y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2: ~2^-60 before rounding
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// Publication of the eliminated columns from the pivot warp to the trailing warps: a counter in shared memory,
// written with release / read with acquire semantics (CTA scope).
__device__ __forceinline__ void pb_publish(unsigned flag_saddr, unsigned count) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(flag_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void pb_wait(unsigned flag_saddr, unsigned want) {
  unsigned f;
  do {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(f) : "r"(flag_saddr) : "memory");
  } while (f < want);
}

// One "vector" of the panel: row 16(P+1)+v below the diagonal block for v < NB, else column v - NB of Y_p.
template <int P>
__device__ __forceinline__ double* pb_vector(double* X, double* YT, int v) {
  constexpr int NB = 112 - 16 * P;
  return (v < NB) ? X + (16 * (P + 1) + v) * PB_LD : YT + (v - NB) * PB_LD;
}

// Final entries of a vector: back to shared memory for the rank-16 update, and straight to L / L^-1 in global memory.
template <int P>
__device__ __forceinline__ void pb_store_vector(const double (&a)[16], double* dst_s, int v, double* __restrict__ A, int64_t ld,
                                                double* __restrict__ Linv, int64_t ldi) {
  constexpr int NB = 112 - 16 * P;
#pragma unroll
  for (int q = 0; q < 8; ++q) *reinterpret_cast<double2*>(dst_s + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
  if (v < NB) {                                   // L[row][16P .. 16P+15]
    double* dst = A + (int64_t)(16 * (P + 1) + v) * ld + 16 * P;
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<double2*>(dst + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
  } else {                                        // Linv[16P + k][col], k = 0..15: coalesced over the lanes
    double* dst = Linv + (int64_t)(16 * P) * ldi + (v - NB);
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[(int64_t)k * ldi] = a[k];
  }
}

// The pivot warp (warp 0, alone on its scheduler while the panel is eliminated).  Lanes 0..15 hold the rows of the
// diagonal 16x16 block, lanes 16..31 the first 16 vectors of the panel; the 16 columns are eliminated in registers.
//  * Pivot row c = lane c's own row.  Its first three entries -- the pivot, and the two that update the next two
//    pivots -- travel by shuffle: they are the column-to-column chain
//        shuffle -> reciprocal -> multiply -> fused multiply-add -> shuffle          (~100 cycles per column).
//  * The whole row is also written to a table in shared memory (tab[c][c+1..15]), followed by (-1/piv, 1/sqrt(piv)).
//    The warp's own updates of the entries >= c+3 read it back from there (a quarter of the instructions of per-entry
//    shuffles), and the trailing warps (pb_trail) eliminate the other 112 vectors from the table alone.  The store
//    also pins the right-looking order: with shuffles only, ptxas evaluates every entry lazily (c dependent
//    shuffle + DFMA steps right before column c needs it, to shorten live ranges), which puts those steps on the
//    chain; lane c cannot store entries it has not brought up to date.
//  * After column c the counter at flag_saddr is released as 16P + c + 1.
// The masks (rows at or above the pivot, failed pivots) are applied to operands beside the chain, not on it.  Entry c
// of every lane is final once column c is eliminated: x_c = (un-normalised entry) / sqrt(piv_c) is the entry of L_d
// for the rows of the diagonal block (sqrt(piv_c) on the diagonal, zero above it) and solves L_d x = b for the rows
// below AND for the columns of Y_p, so it is stored in place right away.
// Why one warp: a first version let all 8 warps eliminate the diagonal block redundantly beside 16 vectors each.
// Shared memory delivers 128 B per clock to the register files however few distinct addresses a load has, so the
// eight copies of every pivot row cost ~220 LSU cycles per column against a chain of ~100 (measured 360 per column).
template <int P>
__device__ __forceinline__ void pb_pivot_warp(double* X, double* YT, double* tab, unsigned flag_saddr, double* __restrict__ A,
                                              int64_t ld, double* __restrict__ Linv, int64_t ldi, int lane, int j0, int* info) {
  // X, YT and tab are shared memory other threads read and write: deliberately NOT __restrict__ (with it the compiler
  // forwards a lane's own stores across __syncwarp and never reads what lane c published)
  constexpr unsigned FULL = 0xffffffffu;
  const int r = lane & 15;
  const bool vec = lane >= 16;
  double* src = vec ? pb_vector<P>(X, YT, r) : X + (16 * P + r) * PB_LD;
  double a[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double2 t = *reinterpret_cast<const double2*>(src + 2 * q);
    a[2 * q] = t.x;
    a[2 * q + 1] = t.y;
  }
  int fail = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const double piv = __shfl_sync(FULL, a[c], c);
    const double rc1 = (c + 1 < 16) ? __shfl_sync(FULL, a[(c + 1) & 15], c) : 0.0;
    const double rc2 = (c + 2 < 16) ? __shfl_sync(FULL, a[(c + 2) & 15], c) : 0.0;
    double* trow = tab + c * PB_LD;
    if (lane == c) {
#pragma unroll
      for (int q = (c + 1) >> 1; q < 8; ++q) *reinterpret_cast<double2*>(trow + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
    }
    __syncwarp();
    double rc[16];
#pragma unroll
    for (int q = (c + 3) >> 1; q < 8; ++q) {
      const double2 t = *reinterpret_cast<const double2*>(trow + 2 * q);
      rc[2 * q] = t.x;
      rc[2 * q + 1] = t.y;
    }
    // positive, normal and finite: one integer compare on the high word (zero, negative, NaN, Inf and denormal
    // pivots all count as "not positive definite").  A failed pivot is replaced by 1 for the reciprocals and the
    // column's entries by zero, so the rest of the block stays finite -- no branch inside the 16 columns.
    const bool ok = (unsigned)(__double2hiint(piv) - 0x00100000) < 0x7fe00000u;
    if (!ok && fail == 0) fail = 16 * P + c + 1;
    const double pivs = ok ? piv : 1.0;
    const double nrcp = -pb_rcp(pivs);
    const double rs = pb_rsqrt(pivs);
    if (lane == c) *reinterpret_cast<double2*>(trow + 16) = make_double2(ok ? nrcp : 0.0, ok ? rs : 0.0);
    __syncwarp();
    if (lane == 0) pb_publish(flag_saddr, 16 * P + c + 1);
    const double am = (ok && (vec || r > c)) ? a[c] : 0.0;      // rows at or above the pivot stay as they are
    const double ax = (ok && (vec || r >= c)) ? a[c] : 0.0;     // ... and are zero right of the diagonal of L_d
    const double nm = am * nrcp;                                // minus the multiplier
    if (c + 1 < 16) a[(c + 1) & 15] = fma(nm, rc1, a[(c + 1) & 15]);
    if (c + 2 < 16) a[(c + 2) & 15] = fma(nm, rc2, a[(c + 2) & 15]);
#pragma unroll
    for (int j = c + 3; j < 16; ++j) a[j] = fma(nm, rc[j], a[j]);
    a[c] = ax * rs;
  }
  if (fail != 0 && lane == 0) atomicCAS(info, 0, j0 + fail);
  if (!vec) {                                     // the diagonal block of L, zeros right of the diagonal
    double* dst = A + (int64_t)(16 * P + r) * ld + 16 * P;
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<double2*>(dst + 2 * q) = make_double2(a[2 * q], a[2 * q + 1]);
    return;
  }
  pb_store_vector<P>(a, src, r, A, ld, Linv, ldi);
}

// A trailing warp: 32 of the panel's other 112 vectors, one per lane, eliminated from the table as the pivot warp
// releases its columns.  No shuffles, no reciprocals: per column one multiply, 15 - c fused multiply-adds against the
// broadcast pivot row, one scaling.
template <int P>
__device__ __forceinline__ void pb_trail(double* X, double* YT, const double* tab, unsigned flag_saddr, double* __restrict__ A,
                                         int64_t ld, double* __restrict__ Linv, int64_t ldi, int v) {
  if (v >= 128) return;
  double* src = pb_vector<P>(X, YT, v);
  double a[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double2 t = *reinterpret_cast<const double2*>(src + 2 * q);
    a[2 * q] = t.x;
    a[2 * q + 1] = t.y;
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    pb_wait(flag_saddr, 16 * P + c + 1);
    const double* trow = tab + c * PB_LD;
    const double2 sc = *reinterpret_cast<const double2*>(trow + 16);      // (-1/piv, 1/sqrt(piv)); zeros for a failed pivot
    double rc[16];
#pragma unroll
    for (int q = (c + 1) >> 1; q < 8; ++q) {
      const double2 t = *reinterpret_cast<const double2*>(trow + 2 * q);
      rc[2 * q] = t.x;
      rc[2 * q + 1] = t.y;
    }
    const double nm = a[c] * sc.x;
#pragma unroll
    for (int j = c + 1; j < 16; ++j) a[j] = fma(nm, rc[j], a[j]);
    a[c] *= sc.y;
  }
  pb_store_vector<P>(a, src, v, A, ld, Linv, ldi);
}

template <int P>
__device__ __forceinline__ void pb_panel(double (&acc)[8][8], double (&yac)[8][8], double* smem, double* __restrict__ A,
                                         int64_t ld, double* __restrict__ Linv, int64_t ldi, int ty, int tx, int tid, int j0,
                                         int* info) {
  double* X = smem + (P & 1) * (2 * 128 * PB_LD);
  double* YT = X + 128 * PB_LD;
  PB_MARK(4 * P);
  // ---- publish: panel column block (diagonal block symmetric) and the panel's rows of Y, column-major ----
  if (ty >= tx) X[(16 * P + ty) * PB_LD + tx] = acc[P][P];
  if (ty > tx) X[(16 * P + tx) * PB_LD + ty] = acc[P][P];
#pragma unroll
  for (int i = P + 1; i < 8; ++i) X[(ty + 16 * i) * PB_LD + tx] = acc[i][P];
#pragma unroll
  for (int j = 0; j <= P; ++j) YT[(tx + 16 * j) * PB_LD + ty] = yac[P][j];
  __syncthreads();
  PB_MARK(4 * P + 1);
  {
    // warp 0 runs the column chain, warps 1, 2, 3 and half of warp 5 trail it with the other vectors (one warp per
    // scheduler; warp 4 would share the pivot warp's), warps 4, 6, 7 go straight to the barrier
    double* tab = smem + 4 * 128 * PB_LD;
    const unsigned flag_saddr = (unsigned)__cvta_generic_to_shared(tab + 16 * PB_LD);
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) pb_pivot_warp<P>(X, YT, tab, flag_saddr, A, ld, Linv, ldi, lane, j0, info);
    else if (warp <= 3 || warp == 5) pb_trail<P>(X, YT, tab, flag_saddr, A, ld, Linv, ldi, 16 + 32 * (warp <= 3 ? warp - 1 : 3) + lane);
  }
  PB_MARK(4 * P + 2);
  if (P == 7) return;
  __syncthreads();
  PB_MARK(4 * P + 3);
  // ---- rank-16 update of the live patches ----
#pragma unroll 2
  for (int k2 = 0; k2 < 8; ++k2) {
    double2 ri[8];
#pragma unroll
    for (int i = P + 1; i < 8; ++i) ri[i] = *reinterpret_cast<const double2*>(X + (ty + 16 * i) * PB_LD + 2 * k2);
#pragma unroll
    for (int j = P + 1; j < 8; ++j) {
      const double2 cj = *reinterpret_cast<const double2*>(X + (tx + 16 * j) * PB_LD + 2 * k2);
#pragma unroll
      for (int i = j; i < 8; ++i) acc[i][j] = fma(-ri[i].y, cj.y, fma(-ri[i].x, cj.x, acc[i][j]));
    }
#pragma unroll
    for (int j = 0; j <= P; ++j) {
      const double2 yj = *reinterpret_cast<const double2*>(YT + (tx + 16 * j) * PB_LD + 2 * k2);
#pragma unroll
      for (int i = P + 1; i < 8; ++i) yac[i][j] = fma(-ri[i].y, yj.y, fma(-ri[i].x, yj.x, yac[i][j]));
    }
  }
}

__global__ void __launch_bounds__(256, 1) potf2inv_blocked_kernel(double* __restrict__ A, int64_t ld, double* __restrict__ Linv,
                                                                  int64_t ldi, int j0, int* __restrict__ info) {
  extern __shared__ __align__(16) double pb_smem[];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  if (tid == 0) *reinterpret_cast<unsigned*>(pb_smem + 4 * 128 * PB_LD + 16 * PB_LD) = 0u;     // released-column counter
  double acc[8][8], yac[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t r = ty + 16 * i, c = tx + 16 * j;
      if (j <= i) {
        acc[i][j] = A[r * ld + c];
        yac[i][j] = (i == j && ty == tx) ? 1.0 : 0.0;
      } else {                                    // strictly-upper 16x16 blocks of both outputs are zero
        A[r * ld + c] = 0.0;
        Linv[r * ldi + c] = 0.0;
      }
    }
  pb_panel<0>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<1>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<2>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<3>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<4>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<5>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<6>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
  pb_panel<7>(acc, yac, pb_smem, A, ld, Linv, ldi, ty, tx, tid, j0, info);
}

