// K3 / K5: information-gain scoring of candidate sample sets and paths.
//
// Replaces the per-candidate Python loops of Agent.greedy (reference
// agent.py:313-347) and Agent.best_path (agent.py:373-400), each of which
// fancy-indexes a fresh (|S|+k)^2 matrix and calls np.linalg.slogdet
// (utils.py:193).  With the base set factored once (A_B = L L^T) and
// Wt = Sigma_{:,B} L^-T resident in HBM (location-major: row `loc` holds the
// N numbers L^-1 Sigma_{B,loc}), a candidate C only needs
//     G = Wt_C Wt_C^T (k x k Gram over N),  P_CC = Sigma_CC - G,
//     logdet(I + D P_CC D),  D = diag(sqrt(delta))
// (SURVEY.md 9.3).  The Gram is one DMMA.8x8x4 per 4 values of N with the SAME
// register as both operands, fed by 256-bit loads of whole 128-byte lines.
//
//   score_sets_k8      one warp per candidate, k <= 8; 8x8 elimination by shuffles
//   score_sets_generic one CTA per candidate, k <= 128; Gram blocks per warp, elimination in smem
//   greedy_utilities   k = 1 closed form over all n locations (agent.py:341)
//   argmax             (max value, then min index) == np.argmax first-max (agent.py:349,402)
//   append             posterior rank-1 downdate after an acquisition, as one new column of Wt
#include "common.cuh"
#include "argmax.cuh"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

#define ALGP_CONST 1.4189385332046727   // 0.5*log(2*pi*e), utils.py:10
// workspace of algp_score_sets_tiled: [arrival counters: 2 x SCORE_SPLIT_MAX_B uint32][fragments].  One counter array per
// PARTS value (2 and 4): a counter must be a multiple of PARTS when a call starts, which mixing the two on one array
// would break.
#define SCORE_SPLIT_MAX_B 16384
#define SCORE_COUNTER_DOUBLES SCORE_SPLIT_MAX_B

struct ScoreArgs {
  KernelParams kp;
  double noise;            // sigma_n^2 (on the diagonal of Sigma, agent.py:90)
  const double* Wt;        // [n x ldw]
  int64_t ldw;
  int ncols16;             // number of valid Wt columns rounded up to 16 (tail columns are zero)
  const double* X;         // [n x d] field coordinates
  const double* pi0;       // [n] base precisions (0 = unsampled)
  const int32_t* idx;      // [B x k]   (-1 = empty slot)
  const double* delta;     // [B x k] or null
  double delta_scalar;
  const uint8_t* skip;     // [n] or null: slots at these locations add nothing (already-mobile, agent.py:377)
  int k;
  int64_t B;
  double H_base;
  double* scores;          // [B]
  double* work;            // k > 128: [grid][kk][kk+1] elimination scratch in global memory
  // k <= 8, columns in L2-sized chunks (one launch per chunk, algp_score_sets_tiled): this launch covers columns
  // [col0, col1); the accumulator fragments travel between launches through gpart [B][32 lanes][2]
  int col0, col1, first, last;
  double* gpart;
  unsigned int* counters;  // PARTS > 1: arrivals per candidate (zero at allocation; every call adds PARTS to each)
};

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld256_l1(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// the same two loads without .nc, for a kernel that writes (other words of) the lines it reads: append_block_dmma_kernel
__device__ __forceinline__ void ld256_rw(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld256_rw_l1(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}

// ---------------------------------------------------------------------------
// k <= 8: one warp per candidate
//
// Lane (g,t) streams row g of the candidate with 256-bit loads: per 128-byte line the four
// t-lanes take 32 B each, so every request is a whole line and the 4 doubles a lane holds feed
// 4 DMMAs as BOTH operands (the k-order inside a line is permuted by t, which a Gram matrix
// does not care about).  UNROLL lines per row are requested before the first MMA of an
// iteration; PREFETCH additionally issues the next iteration's loads before this one's MMAs
// (register double buffering).  A shared-memory cp.async ring was tried and was slower: the
// data then crosses the LSU twice (LDGSTS + LDS) and the kernel becomes LSU-bound.
// ---------------------------------------------------------------------------
template <int UNROLL>
__device__ __forceinline__ void sc_load(double (&v)[4 * UNROLL], const double* p, bool active, int k0, int ncols16) {
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    v[4 * u] = v[4 * u + 1] = v[4 * u + 2] = v[4 * u + 3] = 0.0;
    if (active && k0 + 16 * u < ncols16) ld256(p + k0 + 16 * u, v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  }
}

// Slot g of candidate `cand` (4 lanes per slot): location, precision increment, and whether the slot adds anything
// (empty, zero-precision, already-mobile and repeated slots do not: agent.py:377 sets a boolean flag).
__device__ __forceinline__ bool score_k8_slot(const ScoreArgs& a, int64_t cand, int g, int& my_idx, double& my_delta) {
  my_idx = (g < a.k) ? a.idx[cand * a.k + g] : -1;
  my_delta = (g < a.k) ? (a.delta ? a.delta[cand * a.k + g] : a.delta_scalar) : 0.0;
  bool active = (my_idx >= 0) && (my_delta > 0.0);
  if (active && a.skip && a.skip[my_idx]) active = false;
  // duplicates inside a set are idempotent (agent.py:377): keep the first
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    int o_idx = __shfl_sync(0xffffffffu, my_idx, 4 * s);
    int o_act = __shfl_sync(0xffffffffu, (int)active, 4 * s);
    if (s < g && o_act && o_idx == my_idx) active = false;
  }
  return active;
}

// Epilogue of the k <= 8 kernels: Sigma_CC from the coordinates, the 8 x 8 un-normalised elimination by shuffles and the
// size / precision bookkeeping.  (G0, G1) = Gram entries (g, 2t), (g, 2t+1); only the lower triangle is read.
__device__ __forceinline__ void score_k8_epilogue(const ScoreArgs& a, int64_t cand, int lane, int g, int t, int d, int my_idx,
                                                  double my_delta, bool active, double G0, double G1) {
  // Sigma_CC from coordinates: lane needs x of slot g (row) and slots 2t, 2t+1 (cols)
  double r2a = 0.0, r2b = 0.0;
  for (int j = 0; j < d; ++j) {
    double xg = (my_idx >= 0) ? a.X[(int64_t)my_idx * d + j] * a.kp.inv_ls[j] : 0.0;
    double xa = __shfl_sync(0xffffffffu, xg, 4 * (2 * t));
    double xb = __shfl_sync(0xffffffffu, xg, 4 * (2 * t + 1));
    r2a = fma(xg - xa, xg - xa, r2a);
    r2b = fma(xg - xb, xg - xb, r2b);
  }
  const double sq = active ? sqrt(my_delta) : 0.0;
  const double sqa = __shfl_sync(0xffffffffu, sq, 4 * (2 * t));
  const double sqb = __shfl_sync(0xffffffffu, sq, 4 * (2 * t + 1));
  double m0 = (kern_from_r2(r2a, a.kp.kind, a.kp.outputscale) + ((g == 2 * t) ? a.noise : 0.0) - G0) * sq * sqa;
  double m1 = (kern_from_r2(r2b, a.kp.kind, a.kp.outputscale) + ((g == 2 * t + 1) ? a.noise : 0.0) - G1) * sq * sqb;
  if (g == 2 * t) m0 += 1.0;
  if (g == 2 * t + 1) m1 += 1.0;

  // 8x8 un-normalised elimination; element (i,j) lives in lane 4i + (j>>1), register j&1
  double logdet = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double src = (c & 1) ? m1 : m0;
    const double piv = __shfl_sync(0xffffffffu, src, 4 * c + (c >> 1));
    const double lic = __shfl_sync(0xffffffffu, src, 4 * g + (c >> 1));
    const double lj0 = __shfl_sync(0xffffffffu, src, 4 * (2 * t) + (c >> 1));
    const double lj1 = __shfl_sync(0xffffffffu, src, 4 * (2 * t + 1) + (c >> 1));
    const double f = lic / piv;
    if (g > c) {
      if (2 * t > c) m0 = fma(-f, lj0, m0);
      if (2 * t + 1 > c) m1 = fma(-f, lj1, m1);
    }
    logdet += log(piv);
  }

  // size / precision bookkeeping
  double term = 0.0, nnew = 0.0;
  if (t == 0 && active) {
    const double p0 = a.pi0[my_idx];
    term = log(p0 + my_delta) - (p0 > 0.0 ? log(p0) : 0.0);
    nnew = (p0 > 0.0) ? 0.0 : 1.0;
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    term += __shfl_xor_sync(0xffffffffu, term, o);
    nnew += __shfl_xor_sync(0xffffffffu, nnew, o);
  }
  if (lane == 0) a.scores[cand] = a.H_base + nnew * ALGP_CONST + 0.5 * (logdet - term);
}

// PARTS > 1: a candidate is scored by PARTS independent warps, each over a slice of the columns.  A warp stores its
// accumulator fragment to gpart[cand][part], fences, and bumps the candidate's arrival counter; the warp that arrives
// last adds the other fragments and runs the epilogue ("last one out" -- no barrier, no extra launch; the counters are
// never reset: every call adds exactly PARTS to each, so arrival order is the old value modulo PARTS).  A candidate is
// ~90 us of one warp's time under load, so a batch smaller than the resident warp slots leaves most of the machine idle
// and ends with a long tail: finer work items fix that for short path lists (up to ~1.5 x the warp slots; beyond, whole
// candidates per warp stream better).
template <int UNROLL, bool PREFETCH, int THREADS, int PARTS>
__global__ void __launch_bounds__(THREADS, 3) score_sets_k8_kernel(const ScoreArgs a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.kp.d;
  constexpr int STEP = 16 * UNROLL;

  for (int64_t it = warp0; it < a.B * PARTS; it += nwarps) {
    int64_t cand = it;
    int col0 = a.col0, col1 = a.col1, part = 0;
    if (PARTS > 1) {
      cand = it / PARTS;
      part = (int)(it - cand * PARTS);
      const int span = ((a.ncols16 + PARTS - 1) / PARTS + STEP - 1) / STEP * STEP;
      col0 = part * span;
      col1 = col0 + span;
    }
    int my_idx;
    double my_delta;
    const bool active = score_k8_slot(a, cand, g, my_idx, my_delta);
    const double* row = a.Wt + (int64_t)(active ? my_idx : 0) * a.ldw + 4 * t;

    double c0[4], c1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c0[q] = c1[q] = 0.0;
    const int cend = col1 < a.ncols16 ? col1 : a.ncols16;
    if (!a.first) {                                        // continue the Gram of the earlier column chunks
      const double2 p = *reinterpret_cast<const double2*>(a.gpart + cand * 64 + 2 * lane);
      c0[0] = p.x;
      c1[0] = p.y;
    }
    // empty / duplicate slots issue no loads; the MMA itself is warp-wide
    if (PREFETCH) {
      double cur[4 * UNROLL], nxt[4 * UNROLL];
      sc_load<UNROLL>(cur, row, active, col0, cend);
      for (int k0 = col0; k0 < cend; k0 += STEP) {
        sc_load<UNROLL>(nxt, row, active, k0 + STEP, cend);
#pragma unroll
        for (int q = 0; q < 4 * UNROLL; ++q) dmma884(c0[q & 3], c1[q & 3], cur[q], cur[q]);
#pragma unroll
        for (int q = 0; q < 4 * UNROLL; ++q) cur[q] = nxt[q];
      }
    } else {
      for (int k0 = col0; k0 < cend; k0 += STEP) {
        double v[4 * UNROLL];
        sc_load<UNROLL>(v, row, active, k0, cend);
#pragma unroll
        for (int q = 0; q < 4 * UNROLL; ++q) dmma884(c0[q & 3], c1[q & 3], v[q], v[q]);
      }
    }
    double G0 = (c0[0] + c0[1]) + (c0[2] + c0[3]);         // G[g][2t]
    double G1 = (c1[0] + c1[1]) + (c1[2] + c1[3]);         // G[g][2t+1]
    if (PARTS > 1) {
      double* mine = a.gpart + (cand * PARTS + part) * 64 + 2 * lane;
      *reinterpret_cast<double2*>(mine) = make_double2(G0, G1);
      __threadfence();
      __syncwarp();
      unsigned int old = 0;
      if (lane == 0) old = atomicAdd(a.counters + cand, 1u);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old % PARTS != PARTS - 1) continue;              // another warp finishes this candidate
      __threadfence();
#pragma unroll
      for (int p = 0; p < PARTS; ++p) {
        if (p == part) continue;
        const double2 o = __ldcg(reinterpret_cast<const double2*>(a.gpart + (cand * PARTS + p) * 64 + 2 * lane));
        G0 += o.x;
        G1 += o.y;
      }
    }
    if (!a.last) {                                         // more column chunks to come: park the fragment
      *reinterpret_cast<double2*>(a.gpart + cand * 64 + 2 * lane) = make_double2(G0, G1);
      continue;
    }

    score_k8_epilogue(a, cand, lane, g, t, d, my_idx, my_delta, active, G0, G1);
  }
}

template <int UNROLL, bool PREFETCH, int THREADS>
static void score_k8_launch(const ScoreArgs& a, int sms, int blocks_per_sm, int parts, cudaStream_t st) {
  const int wpb = THREADS / 32;
  int64_t want = (a.B * parts + wpb - 1) / wpb;
  int64_t cap = (int64_t)sms * blocks_per_sm;
  const int grid = (int)(want < cap ? want : cap);
  if (parts == 4) score_sets_k8_kernel<UNROLL, PREFETCH, THREADS, 4><<<grid, THREADS, 0, st>>>(a);
  else if (parts == 2) score_sets_k8_kernel<UNROLL, PREFETCH, THREADS, 2><<<grid, THREADS, 0, st>>>(a);
  else score_sets_k8_kernel<UNROLL, PREFETCH, THREADS, 1><<<grid, THREADS, 0, st>>>(a);
}

// ---------------------------------------------------------------------------
// k <= 8, large batches: ONE persistent launch that sweeps the columns of Wt in L2-sized chunks with the candidates'
// partial Grams resident in shared memory (the form algp_score_sets_tiled picks for batches that stream >= 2 GB).
//
// The chunked launches further down park every candidate's accumulator fragment in global memory between launches
// (34 MB written and 34 MB read per chunk at 65 536 candidates: as much L2 traffic as a 512-column slice of Wt) and pay
// a ramp and a tail per launch, so their best chunk is 1024 columns -- a 100 MB slice of which the 126 MB L2 keeps
// 59 % (4.8 GB of DRAM reads per call for 0.4 GB of distinct rows).  Here an SM OWNS up to SR_CAP candidates for the
// whole call: the lower triangle of a candidate's Gram (36 doubles), its eight row indices after the slot rule, its id
// and a progress word wait in shared memory between chunks -- 65 536 x 328 B = 21 MB over the 148 SMs -- so no
// fragment leaves the SM and the chunk can be as narrow as the L2 likes: 768 columns for the 16 384 rows of
// configs[2], 80 % sector hits, 1.9 GB of DRAM reads (ncu: profiles/r02_prof_score_resident_summary.csv).  With the
// slice in L2 the loop is latency-bound, so a warp keeps FOUR 128-byte lines per row in flight (96 KB per SM) where
// the DRAM-heavy single launch was best with two.
//
// Work items are (candidate, chunk) pairs.  Chunk-0 items are created by taking a shared-memory slot and a ticket
// from a global counter, so a fast SM (fewer SMs sharing its GPC's path to the L2) takes more candidates than a slow
// one and keeps that share for the remaining chunks, which a static partition cannot do.  After one CTA barrier (the
// slot list is final) the 24 warps take the remaining items of their SM chunk-major from a shared-memory counter:
// the whole SM moves through the chunks together, and a batch of only a few candidates per warp (8192 sets: 2.3)
// still splits evenly.  Item (slot, c) needs item (slot, c-1), which has a smaller index and is therefore already
// taken by a warp that is running: the progress word of the slot is polled, nothing can deadlock.  Sums are in
// chunk order whatever the schedule: results are deterministic.  Shared memory is kept to 189 KB: at 218 KB the L1
// that stages the loads in flight shrinks and the kernel is 60 % slower.
// Measured (profiles/r02_score_resident.log): 1.22-1.29 ms per 65 536 sets where the per-chunk launches take
// 1.45-1.55 and the single launch 1.51-1.67 on the same box; 8192 sets 0.214 against 0.234.
// ---------------------------------------------------------------------------
static int g_tile_cols = 0;       // 0 = derive from the L2 size; algp_set_score_tile_cols overrides (tuning / tests)
static int g_resident = 0;        // 0 = auto (resident when the chunk is derived, launches when it is forced), 1 always, -1 never
static bool score_use_resident() { return g_resident > 0 || (g_resident == 0 && g_tile_cols == 0); }

#define SR_WARPS 24
#define SR_CAP (24 * 24)                                   // candidates per SM
#define SR_UNROLL 4
#define SR_SLOT_BYTES (36 * sizeof(double) + 10 * sizeof(int))
#define SR_NO_CAND 0xffffffffu
__global__ void __launch_bounds__(SR_WARPS * 32, 1) score_sets_k8_resident_kernel(const ScoreArgs a, int chunk, unsigned int* next) {
  // [SR_CAP][36] Grams, [SR_CAP][8] rows, [SR_CAP] ids, [SR_CAP] chunks done, {slots taken, items taken}
  extern __shared__ __align__(16) double sr_gram[];
  int* rows = reinterpret_cast<int*>(sr_gram + (size_t)SR_CAP * 36);
  unsigned int* ids = reinterpret_cast<unsigned int*>(rows + SR_CAP * 8);
  volatile int* done = reinterpret_cast<volatile int*>(ids + SR_CAP);
  int* ctr = const_cast<int*>(done) + SR_CAP;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int d = a.kp.d;
  constexpr int STEP = 16 * SR_UNROLL;
  const int tri = g * (g + 1) / 2 + 2 * t;                 // entry (g, 2t) of a slot's lower triangle
  const bool keep0 = 2 * t <= g, keep1 = 2 * t + 1 <= g;   // lower triangle incl. the diagonal: all the epilogue reads
  const int nchunks = a.ncols16 > chunk ? (a.ncols16 + chunk - 1) / chunk : 1;
  if (threadIdx.x == 0) ctr[0] = ctr[1] = 0;
  __syncthreads();

  // item (slot, c): chunk c of the candidate in `slot`.  c == 0 items are created by taking a slot and a ticket.
  int count = 0, total = 0;
  for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1) {
      if (nchunks == 1) break;
      __syncthreads();                                     // every chunk-0 item of this SM is done, the slot list is final
      count = ctr[0] < SR_CAP ? ctr[0] : SR_CAP;
      total = count * (nchunks - 1);
    }
    for (;;) {
      int slot = 0, c = 0;
      unsigned int cu = 0;
      if (phase == 0) {
        if (lane == 0) {
          slot = atomicAdd(&ctr[0], 1);
          cu = slot < SR_CAP ? atomicAdd(next, 1u) : SR_NO_CAND;
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        cu = __shfl_sync(0xffffffffu, cu, 0);
        if (slot >= SR_CAP) break;
        if ((int64_t)cu >= a.B) {                          // tickets ran out: the slot stays empty
          if (lane == 0) ids[slot] = SR_NO_CAND;
          break;
        }
        if (lane == 0) ids[slot] = cu;
      } else {
        int j = 0;
        if (lane == 0) j = atomicAdd(&ctr[1], 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= total) break;
        c = 1 + j / count;
        slot = j - (c - 1) * count;
        cu = ids[slot];
        if (cu == SR_NO_CAND) continue;
        while (done[slot] < c) {}                          // chunk c-1 of this slot is an earlier item: some warp is on it
        __threadfence_block();
      }
      const int64_t cand = cu;
      const int col0 = c * chunk;
      const int cend = (col0 + chunk < a.ncols16) ? col0 + chunk : a.ncols16;
      int my_idx = -1;
      double my_delta = 0.0;
      bool active;
      int my_row;
      double c0[4], c1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) c0[q] = c1[q] = 0.0;
      if (c == 0) {                                        // the slot rule once per candidate; later chunks read the row back
        active = score_k8_slot(a, cand, g, my_idx, my_delta);
        my_row = active ? my_idx : -1;
        if (t == 0) rows[slot * 8 + g] = my_row;
      } else {
        my_row = rows[slot * 8 + g];
        active = my_row >= 0;
        if (keep0) c0[0] = sr_gram[slot * 36 + tri];
        if (keep1) c1[0] = sr_gram[slot * 36 + tri + 1];
      }
      const double* row = a.Wt + (int64_t)(active ? my_row : 0) * a.ldw + 4 * t;
      for (int k0 = col0; k0 < cend; k0 += STEP) {
        double v[4 * SR_UNROLL];
        sc_load<SR_UNROLL>(v, row, active, k0, cend);
#pragma unroll
        for (int q = 0; q < 4 * SR_UNROLL; ++q) dmma884(c0[q & 3], c1[q & 3], v[q], v[q]);
      }
      const double G0 = (c0[0] + c0[1]) + (c0[2] + c0[3]);
      const double G1 = (c1[0] + c1[1]) + (c1[2] + c1[3]);
      if (c + 1 < nchunks) {
        if (keep0) sr_gram[slot * 36 + tri] = G0;
        if (keep1) sr_gram[slot * 36 + tri + 1] = G1;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) done[slot] = c + 1;
        continue;
      }
      if (c > 0) active = score_k8_slot(a, cand, g, my_idx, my_delta);
      score_k8_epilogue(a, cand, lane, g, t, d, my_idx, my_delta, active, G0, G1);
    }
  }
}

// Candidates one resident launch takes: 80 % of the shared-memory slots.  An SM stops taking tickets when its SR_CAP
// slots are full, so every candidate is taken as long as the batch is below the slot count; the 20 % are the room the
// fast SMs need to take more than their even share.
static int64_t score_resident_batch(int sms) { return (int64_t)sms * SR_CAP * 4 / 5; }

static int score_k8_resident(ScoreArgs a, int sms, int chunk, unsigned int* next, cudaStream_t st) {
  const size_t smem = (size_t)SR_CAP * SR_SLOT_BYTES + 2 * sizeof(int);
  static AlgpPerDevice configured;
  if (configured.raise(smem)) {
    ALGP_CUDA(cudaFuncSetAttribute(score_sets_k8_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int64_t batch = score_resident_batch(sms), B = a.B;
  const int32_t* idx = a.idx;
  const double* delta = a.delta;
  double* scores = a.scores;
  for (int64_t b0 = 0; b0 < B; b0 += batch) {
    a.B = B - b0 < batch ? B - b0 : batch;
    a.idx = idx + b0 * a.k;
    a.delta = delta ? delta + b0 * a.k : nullptr;
    a.scores = scores + b0;
    ALGP_CUDA(cudaMemsetAsync(next, 0, sizeof(unsigned int), st));
    score_sets_k8_resident_kernel<<<sms, SR_WARPS * 32, smem, st>>>(a, chunk, next);
    ALGP_LAUNCH_CHECK();
  }
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// k <= 128: one CTA per candidate, the k x k matrix in shared memory.
// 128 < k <= 2048 (long paths, agent.py:373-400 scores paths of "tens to hundreds" of mobile locations):
// the same kernel with the matrix in a per-CTA global scratch (L2 resident).
// ---------------------------------------------------------------------------
#define SG_MAXK 128
#define SG_MAXK_LARGE 2048
template <bool GLOBAL_M>
__global__ void __launch_bounds__(256) score_sets_generic_kernel(const ScoreArgs a) {
  extern __shared__ __align__(16) double sg_smem[];
  const int k = a.k;
  const int kp = (k + 7) >> 3, kk = kp * 8, pitch = kk + 1;
  double* M = GLOBAL_M ? a.work + (size_t)blockIdx.x * kk * pitch : sg_smem;   // [kk][kk+1]
  double* sqd = GLOBAL_M ? sg_smem : M + kk * pitch;                           // [kk] sqrt(delta) or 0
  double* sx = sqd + kk;                     // [kk][d] scaled coordinates
  int* sidx = (int*)(sx + kk * a.kp.d);      // [kk]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int d = a.kp.d;

  for (int64_t cand = blockIdx.x; cand < a.B; cand += gridDim.x) {
    __syncthreads();
    for (int s = tid; s < kk; s += 256) {
      int ix = (s < k) ? a.idx[cand * k + s] : -1;
      double dl = (s < k) ? (a.delta ? a.delta[cand * k + s] : a.delta_scalar) : 0.0;
      bool act = ix >= 0 && dl > 0.0;
      if (act && a.skip && a.skip[ix]) act = false;
      for (int q = 0; q < s && act; ++q) {
        int oix = a.idx[cand * k + q];
        double odl = a.delta ? a.delta[cand * k + q] : a.delta_scalar;
        if (oix == ix && odl > 0.0) act = false;
      }
      sidx[s] = act ? ix : -1;
      sqd[s] = act ? sqrt(dl) : 0.0;
      for (int j = 0; j < d; ++j) sx[s * d + j] = act ? a.X[(int64_t)ix * d + j] * a.kp.inv_ls[j] : 0.0;
    }
    __syncthreads();

    const int npairs = kp * (kp + 1) / 2;
    for (int pr = warp; pr < npairs; pr += 8) {
      int bi = (int)((sqrtf(8.0f * pr + 1.0f) - 1.0f) * 0.5f);
      while ((bi + 1) * (bi + 2) / 2 <= pr) ++bi;
      while (bi * (bi + 1) / 2 > pr) --bi;
      const int bj = pr - bi * (bi + 1) / 2;
      const int ia = sidx[bi * 8 + g], ib = sidx[bj * 8 + g];
      const double* ra = a.Wt + (int64_t)(ia >= 0 ? ia : 0) * a.ldw + 4 * t;
      const double* rb = a.Wt + (int64_t)(ib >= 0 ? ib : 0) * a.ldw + 4 * t;
      double c0[4], c1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) c0[q] = c1[q] = 0.0;
      for (int k0 = 0; k0 < a.ncols16; k0 += 16) {
        double va[4], vb[4];
        ld256_l1(ra + k0, va[0], va[1], va[2], va[3]);
        ld256_l1(rb + k0, vb[0], vb[1], vb[2], vb[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(c0[q], c1[q], ia >= 0 ? va[q] : 0.0, ib >= 0 ? vb[q] : 0.0);
      }
      const double G0 = (c0[0] + c0[1]) + (c0[2] + c0[3]);
      const double G1 = (c1[0] + c1[1]) + (c1[2] + c1[3]);
      const int r = bi * 8 + g;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = bj * 8 + 2 * t + e;
        double r2 = 0.0;
        for (int j = 0; j < d; ++j) {
          double df = sx[r * d + j] - sx[c * d + j];
          r2 = fma(df, df, r2);
        }
        double kv = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((r == c) ? a.noise : 0.0);
        double m = (kv - (e ? G1 : G0)) * sqd[r] * sqd[c] + ((r == c) ? 1.0 : 0.0);
        M[r * pitch + c] = m;
      }
    }
    __syncthreads();

    // un-normalised elimination on the lower triangle of M
    double logdet = 0.0;
    for (int c = 0; c < kk; ++c) {
      const double piv = M[c * pitch + c];
      const double inv = 1.0 / piv;
      if (tid == 0) logdet += log(piv);
      const int rem = kk - 1 - c;
      for (int e = tid; e < rem * rem; e += 256) {
        const int r = c + 1 + e / rem, cc = c + 1 + e % rem;
        if (cc <= r) M[r * pitch + cc] = fma(-M[r * pitch + c] * inv, M[cc * pitch + c], M[r * pitch + cc]);
      }
      __syncthreads();
    }
    // bookkeeping
    double term = 0.0, nnew = 0.0;
    for (int s = tid; s < kk; s += 256)
      if (sidx[s] >= 0) {
        const double p0 = a.pi0[sidx[s]];
        const double dl = sqd[s] * sqd[s];
        term += log(p0 + dl) - (p0 > 0.0 ? log(p0) : 0.0);
        nnew += (p0 > 0.0) ? 0.0 : 1.0;
      }
    // k <= 128 < 256 threads: a single pass of warp sums through shared memory
    term = warp_sum(term);
    nnew = warp_sum(nnew);
    __shared__ double s_t[8], s_n[8];
    if (lane == 0) { s_t[warp] = term; s_n[warp] = nnew; }
    __syncthreads();
    if (tid == 0) {
      double tt = 0.0, nn = 0.0;
      for (int w = 0; w < 8; ++w) { tt += s_t[w]; nn += s_n[w]; }
      a.scores[cand] = a.H_base + nn * ALGP_CONST + 0.5 * (logdet - tt);
    }
  }
}

int make_kernel_params(KernelParams* kp, int d, const double* log_ls_host, double log_os, int kind);

extern "C" int algp_score_sets_large(const double* Wt, int64_t ldw, int64_t ncols, const double* X, int d,
                                     const double* log_ls_host, double log_os, int kind, double noise,
                                     const double* pi0, const int32_t* idx, const double* delta, double delta_scalar,
                                     const uint8_t* skip, int k, int64_t B, double H_base, double* scores, double* work,
                                     int64_t work_doubles, void* stream);

extern "C" int algp_score_sets(const double* Wt, int64_t ldw, int64_t ncols, const double* X, int d,
                               const double* log_ls_host, double log_os, int kind, double noise,
                               const double* pi0, const int32_t* idx, const double* delta, double delta_scalar,
                               const uint8_t* skip, int k, int64_t B, double H_base, double* scores, void* stream) {
  return algp_score_sets_large(Wt, ldw, ncols, X, d, log_ls_host, log_os, kind, noise, pi0, idx, delta, delta_scalar, skip,
                               k <= SG_MAXK ? k : -1, B, H_base, scores, nullptr, 0, stream);
}

extern "C" int64_t algp_score_sets_large_work_doubles(int k, int64_t B) {
  if (k <= SG_MAXK || B <= 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t kk = (k + 7) / 8 * 8, grid = B < (int64_t)sms * 2 ? B : (int64_t)sms * 2;
  return grid * kk * (kk + 1);
}

extern "C" int algp_score_sets_large(const double* Wt, int64_t ldw, int64_t ncols, const double* X, int d,
                                     const double* log_ls_host, double log_os, int kind, double noise,
                                     const double* pi0, const int32_t* idx, const double* delta, double delta_scalar,
                                     const uint8_t* skip, int k, int64_t B, double H_base, double* scores, double* work,
                                     int64_t work_doubles, void* stream) {
  if (!Wt || !X || !pi0 || !idx || !scores || k < 1 || k > SG_MAXK_LARGE || B < 0 || ncols < 0) return ALGP_ERR_INVALID;
  if ((ldw & 3) || ((uintptr_t)Wt & 31)) return ALGP_ERR_INVALID;       // 256-bit row loads
  // k > 128: the scratch may be smaller than the preferred size -- fewer CTAs run (at least one kk x (kk+1) matrix)
  const int64_t kk_l = (k + 7) / 8 * 8, per_cta_l = kk_l * (kk_l + 1);
  if (k > SG_MAXK && B > 0 && (!work || work_doubles < per_cta_l)) return ALGP_ERR_INVALID;
  if (work_doubles < 0 && (k > 8 || (-work_doubles) % 32 != 0)) return ALGP_ERR_INVALID;   // chunked k <= 8 (algp_score_sets_tiled)
  ScoreArgs a;
  a.work = work;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.noise = noise; a.Wt = Wt; a.ldw = ldw;
  a.ncols16 = (int)((ncols + 15) / 16 * 16);
  if (a.ncols16 > ldw) return ALGP_ERR_INVALID;
  a.X = X; a.pi0 = pi0; a.idx = idx; a.delta = delta; a.delta_scalar = delta_scalar; a.skip = skip;
  a.k = k; a.B = B; a.H_base = H_base; a.scores = scores;
  a.col0 = 0; a.col1 = a.ncols16; a.first = 1; a.last = 1; a.gpart = nullptr; a.counters = nullptr;
  if (B == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (k <= 8) {
    // 256-thread CTAs (74 registers: 3 resident per SM), up to 32 CTAs per SM in the grid, warps stride over the
    // candidates.  Two 128-byte lines per row in flight and no register prefetch was the fastest load schedule
    // (profiles/r01_score_variants.log); MANY short CTAs beat a persistent resident grid (profiles/r02_score_tiling.log):
    // CTAs that start at different times keep the warps of an SM out of phase, so loads and DMMAs overlap.
    if (work && work_doubles < 0) {
      // algp_score_sets_tiled, swept in chunks of -work_doubles columns: the persistent launch (its ticket counter is the
      // first word behind the arrival counters), or one launch per chunk with the accumulator fragments in work [B][64]
      const int chunk = (int)(-work_doubles);
      a.gpart = work + SCORE_COUNTER_DOUBLES;
      if (score_use_resident()) return score_k8_resident(a, sms, chunk, (unsigned int*)a.gpart, st);
      for (int c0 = 0; c0 < a.ncols16 || c0 == 0; c0 += chunk) {
        a.col0 = c0;
        a.col1 = c0 + chunk;
        a.first = c0 == 0;
        a.last = c0 + chunk >= a.ncols16;
        score_k8_launch<2, false, 256>(a, sms, 32, 1, st);
        ALGP_LAUNCH_CHECK();
      }
      return ALGP_OK;
    }
    if (work && work_doubles > 0) {
      // algp_score_sets_tiled, one launch: small batches as PARTS column slices per candidate (see the kernel)
      const int64_t slots = (int64_t)sms * 3 * 8;                  // resident warps: 3 CTAs of 8 per SM
      // measured (profiles/r02_score_tiling.log 6d): 256 sets 0.081 -> 0.042 ms and 1000 sets 0.068 -> 0.049 ms with 4 parts,
      // 4096 sets 0.145 -> 0.133 ms with 2; from 8192 sets on whole candidates per warp win (0.225 vs 0.251 ms)
      int parts = 1;
      if (2 * B <= slots) parts = 4;
      else if (2 * B <= 3 * slots) parts = 2;
      const char* e = getenv("ALGP_SCORE_PARTS");                  // tuning
      if (e) parts = atoi(e);
      if (parts != 2 && parts != 4) parts = 1;
      if (parts > 1 && (a.ncols16 < 32 * parts || B > SCORE_SPLIT_MAX_B || work_doubles < SCORE_COUNTER_DOUBLES + B * parts * 64))
        parts = 1;
      if (parts > 1) {
        a.counters = (unsigned int*)work + (parts == 4 ? SCORE_SPLIT_MAX_B : 0);
        a.gpart = work + SCORE_COUNTER_DOUBLES;
      }
      score_k8_launch<2, false, 256>(a, sms, 32, parts, st);
      ALGP_LAUNCH_CHECK();
      return ALGP_OK;
    }
    score_k8_launch<2, false, 256>(a, sms, 32, 1, st);
  } else if (k <= SG_MAXK) {
    const int kp = (k + 7) / 8, kk = kp * 8;
    size_t smem = ((size_t)kk * (kk + 1) + kk + (size_t)kk * d) * 8 + (size_t)kk * 4 + 16;
    static AlgpPerDevice configured;
    if (configured.raise(smem)) {
      ALGP_CUDA(cudaFuncSetAttribute(score_sets_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int grid = (int)(B < (int64_t)sms * 2 ? B : (int64_t)sms * 2);
    score_sets_generic_kernel<false><<<grid, 256, smem, st>>>(a);
  } else {
    const int kp = (k + 7) / 8, kk = kp * 8;
    size_t smem = ((size_t)kk + (size_t)kk * d) * 8 + (size_t)kk * 4 + 16;
    static AlgpPerDevice configured;
    if (configured.raise(smem)) {
      ALGP_CUDA(cudaFuncSetAttribute(score_sets_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int grid = (int)(B < (int64_t)sms * 2 ? B : (int64_t)sms * 2);
    if (work_doubles / per_cta_l < grid) grid = (int)(work_doubles / per_cta_l);
    score_sets_generic_kernel<true><<<grid, 256, smem, st>>>(a);
  }
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// ---------------------------------------------------------------------------
// k <= 8 with the columns of Wt in L2-sized chunks.
//
// Large batches of random sets re-read every row of Wt many times (configs[2]: 65 536 sets over 12 288 distinct rows of
// 32 KB, 403 MB against a 126 MB L2; ncu round 1: 7.7 GB of DRAM reads per launch for 0.4 GB of compulsory traffic).
// With one launch per column chunk the slice of Wt a launch touches (n_rows x chunk x 8 B) is L2-sized and the
// accumulator fragments wait in `work` between launches.  Measured on B200 (profiles/r02_score_tiling.log): 1.50 ms
// at a 1024-column chunk against 1.65-1.70 ms for the single launch; smaller chunks lose more to the per-chunk
// prologue than they gain in hit rate, and a 36-entry DFMA Gram (instead of the 64-entry DMMA tile) or a persistent
// resident grid were both slower.
// ---------------------------------------------------------------------------
extern "C" int algp_set_score_tile_cols(int cols) {
  if (cols < -1 || (cols > 0 && (cols & 63))) return ALGP_ERR_INVALID;
  g_tile_cols = cols;
  return ALGP_OK;
}

extern "C" int algp_set_score_resident(int mode) {
  if (mode < -1 || mode > 1) return ALGP_ERR_INVALID;
  g_resident = mode;
  return ALGP_OK;
}

// counters + fragments: [B][64] between chunk launches, [B][4 parts][64] for batches small enough to be split
extern "C" int64_t algp_score_sets_tiled_work_doubles(int64_t B) {
  if (B <= 0) return 0;
  return SCORE_COUNTER_DOUBLES + (B <= SCORE_SPLIT_MAX_B ? B * 4 * 64 : B * 64);
}

// The column chunk algp_score_sets_tiled uses for this call: > 0 chunked launches, 0 one launch (with the workspace:
// split candidates for small batches), -1 the plain single launch.
static int tiled_chunk_cols(int k, int64_t B, int64_t ncols, int64_t n_rows, int* chunk_out) {
  int chunk = g_tile_cols;
  if (chunk == 0) {
    const double bytes = 8.0 * (double)B * k * (double)ncols;
    // measured on configs[2] (profiles/r02_score_resident.log): the persistent sweep wins from ~8000 sets of 8 on
    // (2 GB of row reads: 8192 sets 0.214 vs 0.234 ms; 16 384 sets 0.357 vs 0.410; 65 536 sets 1.22-1.29 vs 1.48-1.61;
    // 4096 sets are level with the split candidates), the per-chunk launches only from ~12 GB
    if (bytes >= (score_use_resident() ? 2.0e9 : 12.0e9) && ncols >= 2048) {
      int dev = 0, l2 = 64 << 20;
      ALGP_CUDA(cudaGetDevice(&dev));
      ALGP_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
      // the slice of a chunk about the size of the L2 for the chunked launches (1024 columns for the 16 384 rows of
      // configs[2]), three quarters of it for the resident form, whose warps are not all on the same chunk (768)
      const int64_t budget = score_use_resident() ? (int64_t)l2 / 4 * 3 : (int64_t)l2;
      int64_t c = (budget / (8 * n_rows) + 32) / 64 * 64;
      chunk = (int)(c < 256 ? 256 : c);
    }
  }
  *chunk_out = chunk;
  return ALGP_OK;
}

// Kernel launches one algp_score_sets_tiled call of this shape makes (the chunked form launches the scoring kernel
// once per column chunk); <= 0 on error.  For callers that count launches (bench.py's gpu_launches).
extern "C" int algp_score_sets_tiled_launches(int k, int64_t B, int64_t ncols, int64_t n_rows) {
  if (k < 1 || k > 8 || B <= 0 || n_rows < 1 || ncols < 0) return 0;
  int chunk = 0;
  if (tiled_chunk_cols(k, B, ncols, n_rows, &chunk)) return 0;
  if (chunk <= 0) return 1;
  if (score_use_resident()) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t batch = score_resident_batch(sms);
    return (int)((B + batch - 1) / batch);
  }
  const int64_t ncols16 = (ncols + 15) / 16 * 16;
  int launches = 0;
  for (int64_t c0 = 0; c0 < ncols16 || c0 == 0; c0 += chunk) ++launches;
  return launches;
}

extern "C" int algp_score_sets_tiled(const double* Wt, int64_t ldw, int64_t ncols, int64_t n_rows, const double* X, int d,
                                     const double* log_ls_host, double log_os, int kind, double noise, const double* pi0,
                                     const int32_t* idx, const double* delta, double delta_scalar, const uint8_t* skip,
                                     int k, int64_t B, double H_base, double* scores, double* work, int64_t work_doubles,
                                     void* stream) {
  if (k < 1 || k > 8 || B < 0 || n_rows < 1) return ALGP_ERR_INVALID;
  if (B > 0 && (!work || work_doubles < algp_score_sets_tiled_work_doubles(B) || ((uintptr_t)work & 15))) return ALGP_ERR_INVALID;
  // g_tile_cols > 0: that chunk; -1: plain single launch; 0 (auto): a sweep in L2-sized column chunks where ONE call
  // streams enough for the L2 hit rate to pay for the per-chunk prologues (persistent form: >= ~2 GB), else the single
  // launch (small batches: split candidates)
  int chunk = 0;
  const int rc = tiled_chunk_cols(k, B, ncols, n_rows, &chunk);
  if (rc) return rc;
  if (chunk > 0)
    return algp_score_sets_large(Wt, ldw, ncols, X, d, log_ls_host, log_os, kind, noise, pi0, idx, delta, delta_scalar, skip,
                                 k, B, H_base, scores, B > 0 ? work : (double*)1, -(int64_t)chunk, stream);
  return algp_score_sets_large(Wt, ldw, ncols, X, d, log_ls_host, log_os, kind, noise, pi0, idx, delta, delta_scalar, skip, k,
                               B, H_base, scores, chunk < 0 ? nullptr : work, chunk < 0 ? 0 : work_doubles, stream);
}

// ---------------------------------------------------------------------------
// greedy (k = 1) utilities, argmax, rank-1 append
// ---------------------------------------------------------------------------
// entropy criterion, agent.py:341 with the restructuring of SURVEY.md 9.3:
// ut_i = [pi_i == 0]*CONST + 0.5*(log1p(dS*P_ii) - log(pi_i + dS) + [pi_i > 0] log pi_i);  -inf where already static
__global__ void greedy_utilities_kernel(const double* __restrict__ diagP, const double* __restrict__ pi,
                                        const uint8_t* __restrict__ is_static, double d_static, int64_t n,
                                        double* __restrict__ ut) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u;
  if (is_static[i]) {
    u = -INFINITY;
  } else {
    const double p = pi[i];
    u = (p > 0.0 ? 0.0 : ALGP_CONST) + 0.5 * (log1p(d_static * diagP[i]) - log(p + d_static) + (p > 0.0 ? log(p) : 0.0));
  }
  ut[i] = u;
}

__global__ void argmax_stage2_kernel(const ArgPair* __restrict__ part, int np, ArgPair* __restrict__ out) {
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (int p = threadIdx.x; p < np; p += 32)
    if (arg_better(part[p].v, part[p].i, bv, bi)) { bv = part[p].v; bi = part[p].i; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (threadIdx.x == 0) { out->v = bv; out->i = bi; }
}

extern "C" int64_t algp_argmax_work_bytes(void) { return (int64_t)ARGMAX_BLOCKS * sizeof(ArgPair); }

// out_pair: {double value; int64 index} on the device; index = position + idx_offset (global id of a shard)
// one launch for vectors a single CTA scans in a few microseconds (the per-step argmax of a scoring batch)
__global__ void __launch_bounds__(1024) argmax_one_cta_kernel(const double* __restrict__ x, int64_t n, int64_t idx_offset,
                                                              ArgPair* __restrict__ out) {
  double bv;
  long long bi;
  argmax_scan(x, n, idx_offset, bv, bi);
  argmax_block_reduce_1024(bv, bi);
  if (threadIdx.x == 0) { out->v = bv; out->i = bi; }
}

extern "C" int algp_argmax(const double* x, int64_t n, int64_t idx_offset, void* out_pair, void* work, void* stream) {
  if (!x || !out_pair || !work || n <= 0) return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= ARGMAX_ONE_CTA_MAX) {
    argmax_one_cta_kernel<<<1, 1024, 0, st>>>(x, n, idx_offset, (ArgPair*)out_pair);
    ALGP_LAUNCH_CHECK();
    return ALGP_OK;
  }
  int blocks = (int)((n + 255) / 256 < ARGMAX_BLOCKS ? (n + 255) / 256 : ARGMAX_BLOCKS);
  argmax_stage1_kernel<<<blocks, 256, 0, st>>>(x, n, idx_offset, (ArgPair*)work);
  ALGP_LAUNCH_CHECK();
  argmax_stage2_kernel<<<1, 32, 0, st>>>((const ArgPair*)work, blocks, (ArgPair*)out_pair);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// Candidate slots outside [-1, n): the reference indexes NumPy arrays with them (agent.py:377) and gets an IndexError;
// the scoring kernels would read out of bounds.  One pass over the (device copy of the) slot array -- 2 MB for
// configs[2], ~3 us -- counts them and turns them into empty slots, so that the scoring launched behind it is safe;
// the count travels to the host together with the winner, so the check costs no extra synchronisation.
__global__ void check_indices_kernel(int32_t* __restrict__ idx, int64_t count, int64_t n, unsigned long long* __restrict__ bad) {
  unsigned mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx[i];
    if (v < -1 || v >= n) {
      idx[i] = -1;
      ++mine;
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(bad, (unsigned long long)mine);
}

extern "C" int algp_check_indices(int32_t* idx, int64_t count, int64_t n, int64_t* bad_count, void* stream) {
  if (!bad_count || count < 0 || n < 0 || (count > 0 && !idx)) return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  ALGP_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int64_t), st));
  if (count == 0) return ALGP_OK;
  int64_t blocks = (count + 4 * 256 - 1) / (4 * 256);
  if (blocks > 4 * 148) blocks = 4 * 148;
  check_indices_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, count, n, (unsigned long long*)bad_count);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

extern "C" int algp_greedy_utilities(const double* diagP, const double* pi, const uint8_t* is_static, double d_static,
                                     int64_t n, double* ut, void* stream) {
  if (!diagP || !pi || !is_static || !ut || n < 0) return ALGP_ERR_INVALID;
  if (n == 0) return ALGP_OK;
  greedy_utilities_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(diagP, pi, is_static, d_static, n, ut);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

// append: P <- P - P_:j P_j: / (P_jj + 1/delta) expressed as one new column of Wt:
//   w_loc = (Sigma[loc,j] - Wt[loc,:ncols] . Wt[j,:ncols]) / sqrt(P_jj + 1/delta),  diagP[loc] -= w_loc^2
struct AppendArgs {
  KernelParams kp;
  double noise;
  double* Wt; int64_t ldw; int ncols;      // current number of columns; the new one is written at column ncols
  const double* X; int64_t n;
  double* diagP; double* pi; uint8_t* is_static;
  const long long* j_dev;                  // chosen location (device scalar, e.g. the index of an ArgPair)
  double delta; int mark_static;
  double* scratch;                         // [2]: denom, (unused)
};

__global__ void append_prepare_kernel(const AppendArgs a) {
  const long long j = *a.j_dev;
  a.scratch[0] = sqrt(a.diagP[j] + 1.0 / a.delta);
  a.pi[j] += a.delta;
  if (a.mark_static) a.is_static[j] = 1;
}

__global__ void __launch_bounds__(256) append_column_kernel(const AppendArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t loc = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (loc >= a.n) return;
  const long long j = *a.j_dev;
  const double* rl = a.Wt + loc * a.ldw;
  const double* rj = a.Wt + (int64_t)j * a.ldw;
  double s0 = 0.0, s1 = 0.0;
  int k = 2 * lane;
  for (; k + 1 < a.ncols; k += 64) {
    double2 x = *reinterpret_cast<const double2*>(rl + k);
    double2 y = *reinterpret_cast<const double2*>(rj + k);
    s0 = fma(x.x, y.x, s0);
    s1 = fma(x.y, y.y, s1);
  }
  if (k < a.ncols) s0 = fma(rl[k], rj[k], s0);
  const double dot = warp_sum(s0 + s1);
  if (lane == 0) {
    const int d = a.kp.d;
    double r2 = 0.0;
    for (int q = 0; q < d; ++q) {
      double df = (a.X[loc * d + q] - a.X[(int64_t)j * d + q]) * a.kp.inv_ls[q];
      r2 = fma(df, df, r2);
    }
    double sig = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((loc == j) ? a.noise : 0.0);
    double w = (sig - dot) / a.scratch[0];
    a.scratch[2 + loc] = w;                // staged: row j of Wt is still being read by other warps
    a.diagP[loc] -= w * w;
  }
}

__global__ void append_commit_kernel(double* __restrict__ Wt, int64_t ldw, int col, const double* __restrict__ w, int64_t n) {
  const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (loc < n) Wt[loc * ldw + col] = w[loc];
}

// ---------------------------------------------------------------------------
// Block append: commit k <= 16 locations at once (the mobile readings of the chosen path, agent.py:179-192).
// k rank-1 appends read Wt k times (HBM-bound, 8 n N bytes each); here Wt is read ONCE:
//   1. the k x k block P_CC = Sigma_CC - Wt_C Wt_C^T (one warp per pair),
//   2. its sequential elimination in one warp: denom_c and T[r][c] = w_c[j_r], exactly the values the k
//      successive rank-1 appends would produce at the chosen rows,
//   3. one pass over the locations: each warp takes 4 rows of Wt at a time, the k chosen rows are staged in
//      shared memory by 256-column chunk, 4 x 16 dot products accumulate in registers, then
//      w_c[loc] = (P0[loc,j_c] - sum_{c'<c} w_c'[loc] T[c][c']) / denom_c, written as k new columns.
// ---------------------------------------------------------------------------
#define AB_K 16          // slots per block
#define AB_LB 4          // rows of Wt a warp carries at a time
#define AB_CH 256        // columns staged per chunk
struct AppendBlockArgs {
  KernelParams kp;
  double noise;
  double* Wt; int64_t ldw; int ncols;
  const double* X; int64_t n;
  double* diagP; double* pi; uint8_t* is_static;
  const long long* idx;                    // [k] chosen locations (device), distinct
  const double* delta; double delta_scalar;
  int k, mark_static;
  double* pcc;                             // [AB_K x AB_K] P_CC, then T
  double* denom;                           // [AB_K]
};

__global__ void __launch_bounds__(256) append_block_pairs_kernel(const AppendBlockArgs a) {
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int r = pair / AB_K, c = pair % AB_K;
  if (r >= a.k || c > r) return;
  const long long jr = a.idx[r], jc = a.idx[c];
  const double* rr = a.Wt + jr * a.ldw;
  const double* rc = a.Wt + jc * a.ldw;
  double s = 0.0;
  for (int q = lane; q < a.ncols; q += 32) s = fma(rr[q], rc[q], s);
  s = warp_sum(s);
  if (lane == 0) {
    const int d = a.kp.d;
    double r2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double df = (a.X[jr * d + q] - a.X[jc * d + q]) * a.kp.inv_ls[q];
      r2 = fma(df, df, r2);
    }
    const double v = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((jr == jc) ? a.noise : 0.0) - s;
    a.pcc[r * AB_K + c] = v;
    a.pcc[c * AB_K + r] = v;
  }
}

__global__ void append_block_eliminate_kernel(const AppendBlockArgs a) {
  // one warp; lane r owns row r of the k x k block
  __shared__ double P[AB_K][AB_K + 1];
  const int r = threadIdx.x;
  const int k = a.k;
  if (r < k)
    for (int c = 0; c < k; ++c) P[r][c] = a.pcc[r * AB_K + c];
  __syncwarp();
  for (int c = 0; c < k; ++c) {
    const double dl = a.delta ? a.delta[c] : a.delta_scalar;
    const double den = sqrt(P[c][c] + 1.0 / dl);
    __syncwarp();
    double w = 0.0;
    if (r < k) w = P[r][c] / den;                  // w_c at location j_r
    __syncwarp();
    if (r < k) {
      P[r][c] = w;                                 // column c now holds T[r][c] = w_c[j_r]
      if (r == 0) a.denom[c] = den;
    }
    __syncwarp();
    if (r < k)
      for (int s2 = c + 1; s2 < k; ++s2) P[r][s2] = fma(-w, P[s2][c], P[r][s2]);
    __syncwarp();
  }
  if (r < k) {
    for (int c = 0; c < k; ++c) a.pcc[r * AB_K + c] = P[r][c];
    const long long j = a.idx[r];
    a.pi[j] += a.delta ? a.delta[r] : a.delta_scalar;
    if (a.mark_static) a.is_static[j] = 1;
  }
}

__global__ void __launch_bounds__(256) append_block_kernel(const AppendBlockArgs a) {
  extern __shared__ __align__(16) double ab_smem[];
  double* rows = ab_smem;                          // [AB_K][AB_CH] chunk of the chosen rows
  double* Ts = rows + AB_K * AB_CH;                // [AB_K][AB_K] T
  double* dn = Ts + AB_K * AB_K;                   // [AB_K] denom
  double* xj = dn + AB_K;                          // [AB_K][d] scaled coordinates of the chosen locations
  double* stage = xj + AB_K * ALGP_MAX_D;          // [8 warps][AB_LB][AB_K] per-warp scratch
  long long* jloc = (long long*)(stage + 8 * AB_LB * AB_K);   // [AB_K]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = a.k, d = a.kp.d;
  for (int e = tid; e < AB_K * AB_K; e += 256) Ts[e] = a.pcc[e];
  if (tid < AB_K) {
    dn[tid] = tid < k ? a.denom[tid] : 1.0;
    jloc[tid] = tid < k ? a.idx[tid] : -1;
  }
  for (int e = tid; e < AB_K * d; e += 256) {
    const int c = e / d, q = e % d;
    xj[c * ALGP_MAX_D + q] = c < k ? a.X[a.idx[c] * d + q] * a.kp.inv_ls[q] : 0.0;
  }
  double* mine = stage + warp * AB_LB * AB_K;
  const int64_t groups = (a.n + 8 * AB_LB - 1) / (8 * AB_LB);
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
    const int64_t loc0 = g * (8 * AB_LB) + warp * AB_LB;
    double acc[AB_LB][AB_K];
#pragma unroll
    for (int l = 0; l < AB_LB; ++l)
#pragma unroll
      for (int c = 0; c < AB_K; ++c) acc[l][c] = 0.0;
    for (int c0 = 0; c0 < a.ncols; c0 += AB_CH) {
      __syncthreads();                             // the previous chunk is consumed
      for (int e = tid; e < AB_K * (AB_CH / 2); e += 256) {
        const int c = e / (AB_CH / 2), q = (e % (AB_CH / 2)) * 2;
        double2 v = make_double2(0.0, 0.0);
        if (c < k) {
          const double* src = a.Wt + a.idx[c] * a.ldw + c0 + q;
          if (c0 + q + 1 < a.ncols) v = *reinterpret_cast<const double2*>(src);
          else if (c0 + q < a.ncols) v.x = src[0];
        }
        *reinterpret_cast<double2*>(rows + c * AB_CH + q) = v;
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < AB_CH / 64; ++it) {
        const int q = it * 64 + 2 * lane;
        double2 x[AB_LB];
#pragma unroll
        for (int l = 0; l < AB_LB; ++l) {
          x[l] = make_double2(0.0, 0.0);
          const int64_t loc = loc0 + l;
          if (loc < a.n) {
            const double* src = a.Wt + loc * a.ldw + c0 + q;
            if (c0 + q + 1 < a.ncols) x[l] = *reinterpret_cast<const double2*>(src);
            else if (c0 + q < a.ncols) x[l].x = src[0];
          }
        }
#pragma unroll
        for (int c = 0; c < AB_K; ++c) {
          const double2 y = *reinterpret_cast<const double2*>(rows + c * AB_CH + q);
#pragma unroll
          for (int l = 0; l < AB_LB; ++l) acc[l][c] = fma(x[l].x, y.x, fma(x[l].y, y.y, acc[l][c]));
        }
      }
    }
    // butterfly: every lane ends with all AB_LB x AB_K dot products; lane t keeps entry t (and t + 32)
#pragma unroll
    for (int l = 0; l < AB_LB; ++l)
#pragma unroll
      for (int c = 0; c < AB_K; ++c) {
        const double v = warp_sum(acc[l][c]);
        if (lane == ((l * AB_K + c) & 31)) mine[l * AB_K + c] = v;
      }
    __syncwarp();
    // P0[loc, j_c] = Sigma[loc, j_c] - dot
    for (int t = lane; t < AB_LB * AB_K; t += 32) {
      const int l = t / AB_K, c = t % AB_K;
      const int64_t loc = loc0 + l;
      double p = 0.0;
      if (loc < a.n && c < k) {
        double r2 = 0.0;
        for (int q = 0; q < d; ++q) {
          const double df = a.X[loc * d + q] * a.kp.inv_ls[q] - xj[c * ALGP_MAX_D + q];
          r2 = fma(df, df, r2);
        }
        p = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((loc == jloc[c]) ? a.noise : 0.0) - mine[t];
      }
      mine[t] = p;
    }
    __syncwarp();
    // the k successive rank-1 columns of this row: lane l < AB_LB owns row loc0 + l
    if (lane < AB_LB && loc0 + lane < a.n) {
      double* pr = mine + lane * AB_K;
      double ss = 0.0;
      for (int c = 0; c < k; ++c) {
        double v = pr[c];
        for (int c1 = 0; c1 < c; ++c1) v = fma(-pr[c1], Ts[c * AB_K + c1], v);
        v /= dn[c];
        pr[c] = v;                                   // w_c[loc]
        ss = fma(v, v, ss);
      }
      a.diagP[loc0 + lane] -= ss;
    }
    __syncwarp();
    for (int t = lane; t < AB_LB * AB_K; t += 32) {
      const int l = t / AB_K, c = t % AB_K;
      if (loc0 + l < a.n && c < k) a.Wt[(loc0 + l) * a.ldw + a.ncols + c] = mine[t];
    }
    __syncwarp();
  }
}

// Step 3 on DMMA fragments (the form algp_append_block launches when the rows of Wt are 32-byte aligned and the pitch
// is a multiple of 16): the pass is the skinny product D[n x 16] = Wt[n x ncols] . Wt_C[16 x ncols]^T.  A warp owns 16
// rows of Wt (two A fragments) for all columns and multiplies them with the two B fragments of the 16 chosen rows,
// which every warp re-reads through L1: per 16 columns four 256-bit loads (a lane takes 32 bytes of its row, whose 4
// doubles feed 4 DMMAs; A and B lanes take the same columns, so the k-order inside a line is permuted alike on both
// sides) and 16 DMMAs.  No shared-memory staging and no block barrier in the column loop: the scalar version above
// staged the chosen rows chunk by chunk between two barriers with 16 warps per SM and reached 1.07 TB/s on the
// 40 000 x 2264 matrix of the configs[4] episode (0.68 ms per path commit; profiles/r02_episode_batch.log).
// Columns [ncols, ncols rounded up to 16) of every row must be zero, as they are in a zero-filled Wt that only ever
// gains columns (the scoring kernels read the same tail); a warp reads its own rows before it writes their new
// columns, so the A side of the tail is zero whatever a chosen row's owner has written there meanwhile.
#define ABD_ROWS 16
#define ABD_WARPS 4
__global__ void __launch_bounds__(ABD_WARPS * 32) append_block_dmma_kernel(const AppendBlockArgs a) {
  __shared__ double Ts[AB_K * AB_K], dn[AB_K], xj[AB_K * ALGP_MAX_D];
  __shared__ long long jloc[AB_K];
  __shared__ double stage_all[ABD_WARPS][ABD_ROWS][AB_K + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k = a.k, d = a.kp.d;
  for (int e = tid; e < AB_K * AB_K; e += ABD_WARPS * 32) Ts[e] = a.pcc[e];
  if (tid < AB_K) {
    dn[tid] = tid < k ? a.denom[tid] : 1.0;
    jloc[tid] = tid < k ? a.idx[tid] : -1;
  }
  for (int e = tid; e < AB_K * d; e += ABD_WARPS * 32) {
    const int c = e / d, q = e % d;
    xj[c * ALGP_MAX_D + q] = c < k ? a.X[a.idx[c] * d + q] * a.kp.inv_ls[q] : 0.0;
  }
  __syncthreads();
  const int64_t loc0 = ((int64_t)blockIdx.x * ABD_WARPS + warp) * ABD_ROWS;
  if (loc0 >= a.n) return;
  double (*stage)[AB_K + 1] = stage_all[warp];

  const int64_t ra0 = loc0 + g < a.n ? loc0 + g : a.n - 1;           // rows past the end repeat the last row (discarded)
  const int64_t ra1 = loc0 + 8 + g < a.n ? loc0 + 8 + g : a.n - 1;
  const bool b0_on = g < k, b1_on = g + 8 < k;
  const double* pa0 = a.Wt + ra0 * a.ldw + 4 * t;
  const double* pa1 = a.Wt + ra1 * a.ldw + 4 * t;
  const double* pb0 = a.Wt + (b0_on ? a.idx[g] : 0) * a.ldw + 4 * t;
  const double* pb1 = a.Wt + (b1_on ? a.idx[g + 8] : 0) * a.ldw + 4 * t;
  const int ncols16 = (a.ncols + 15) / 16 * 16;
  // acc[A half][B half][chain]: D[A half * 8 + g][B half * 8 + 2t + {0, 1}]
  double c0[2][2][2], c1[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) c0[i][j][0] = c0[i][j][1] = c1[i][j][0] = c1[i][j][1] = 0.0;
  for (int k0 = 0; k0 < ncols16; k0 += 16) {
    double va0[4], va1[4], vb0[4], vb1[4];
    ld256_rw(pa0 + k0, va0[0], va0[1], va0[2], va0[3]);
    ld256_rw(pa1 + k0, va1[0], va1[1], va1[2], va1[3]);
    vb0[0] = vb0[1] = vb0[2] = vb0[3] = 0.0;
    vb1[0] = vb1[1] = vb1[2] = vb1[3] = 0.0;
    if (b0_on) ld256_rw_l1(pb0 + k0, vb0[0], vb0[1], vb0[2], vb0[3]);
    if (b1_on) ld256_rw_l1(pb1 + k0, vb1[0], vb1[1], vb1[2], vb1[3]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dmma884(c0[0][0][q & 1], c1[0][0][q & 1], va0[q], vb0[q]);
      dmma884(c0[0][1][q & 1], c1[0][1][q & 1], va0[q], vb1[q]);
      dmma884(c0[1][0][q & 1], c1[1][0][q & 1], va1[q], vb0[q]);
      dmma884(c0[1][1][q & 1], c1[1][1][q & 1], va1[q], vb1[q]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      stage[i * 8 + g][j * 8 + 2 * t] = c0[i][j][0] + c0[i][j][1];
      stage[i * 8 + g][j * 8 + 2 * t + 1] = c1[i][j][0] + c1[i][j][1];
    }
  __syncwarp();
  // P0[loc, j_c] = Sigma[loc, j_c] - dot
  for (int e = lane; e < ABD_ROWS * AB_K; e += 32) {
    const int l = e / AB_K, c = e % AB_K;
    const int64_t loc = loc0 + l;
    double p = 0.0;
    if (loc < a.n && c < k) {
      double r2 = 0.0;
      for (int q = 0; q < d; ++q) {
        const double df = a.X[loc * d + q] * a.kp.inv_ls[q] - xj[c * ALGP_MAX_D + q];
        r2 = fma(df, df, r2);
      }
      p = kern_from_r2(r2, a.kp.kind, a.kp.outputscale) + ((loc == jloc[c]) ? a.noise : 0.0) - stage[l][c];
    }
    stage[l][c] = p;
  }
  __syncwarp();
  // the k successive rank-1 columns of a row: lane l < 16 owns row loc0 + l
  if (lane < ABD_ROWS && loc0 + lane < a.n) {
    double* pr = stage[lane];
    double ss = 0.0;
    for (int c = 0; c < k; ++c) {
      double v = pr[c];
      for (int c1 = 0; c1 < c; ++c1) v = fma(-pr[c1], Ts[c * AB_K + c1], v);
      v /= dn[c];
      pr[c] = v;                                     // w_c[loc]
      ss = fma(v, v, ss);
    }
    a.diagP[loc0 + lane] -= ss;
  }
  __syncwarp();
  for (int e = lane; e < ABD_ROWS * AB_K; e += 32) {
    const int l = e / AB_K, c = e % AB_K;
    if (loc0 + l < a.n && c < k) a.Wt[(loc0 + l) * a.ldw + a.ncols + c] = stage[l][c];
  }
}

extern "C" int64_t algp_append_block_work_doubles(void) { return AB_K * AB_K + AB_K; }

// Commit k (1..16) DISTINCT locations idx[k] (device int64) with precision increments delta[k] (device, or NULL for
// delta_scalar): the same k new columns of Wt, diagP, pi and flags as k successive algp_append calls in that order.
static int g_append_block_scalar = 0;      // tuning / tests: 1 = the scalar shared-memory pass instead of the DMMA one
extern "C" int algp_set_append_block_scalar(int on) {
  g_append_block_scalar = on != 0;
  return ALGP_OK;
}

extern "C" int algp_append_block(double* Wt, int64_t ldw, int64_t ncols, const double* X, int64_t n, int d,
                                 const double* log_ls_host, double log_os, int kind, double noise, double* diagP,
                                 double* pi, uint8_t* is_static, const void* idx_dev, int k, const double* delta_dev,
                                 double delta_scalar, int mark_static, double* work, void* stream) {
  if (!Wt || !X || !diagP || !pi || !is_static || !idx_dev || !work || ncols < 0 || k < 1 || k > AB_K || ncols + k > ldw ||
      (ldw & 1) || (!delta_dev && !(delta_scalar > 0.0)))
    return ALGP_ERR_INVALID;
  AppendBlockArgs a;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.noise = noise; a.Wt = Wt; a.ldw = ldw; a.ncols = (int)ncols; a.X = X; a.n = n;
  a.diagP = diagP; a.pi = pi; a.is_static = is_static; a.idx = (const long long*)idx_dev;
  a.delta = delta_dev; a.delta_scalar = delta_scalar; a.k = k; a.mark_static = mark_static;
  a.pcc = work; a.denom = work + AB_K * AB_K;
  if (n == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  append_block_pairs_kernel<<<AB_K * AB_K / 8, 256, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  append_block_eliminate_kernel<<<1, 32, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  if (a.ldw % 16 == 0 && ((uintptr_t)Wt & 31) == 0 && !g_append_block_scalar) {
    const int64_t warps = (n + ABD_ROWS - 1) / ABD_ROWS;
    append_block_dmma_kernel<<<(unsigned)((warps + ABD_WARPS - 1) / ABD_WARPS), ABD_WARPS * 32, 0, st>>>(a);
    ALGP_LAUNCH_CHECK();
    return ALGP_OK;
  }
  const size_t smem = (size_t)(AB_K * AB_CH + AB_K * AB_K + AB_K + AB_K * ALGP_MAX_D + 8 * AB_LB * AB_K) * 8 + AB_K * 8;
  static AlgpPerDevice configured;
  if (configured.raise(1)) {
    ALGP_CUDA(cudaFuncSetAttribute(append_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t groups = (n + 8 * AB_LB - 1) / (8 * AB_LB);
  const int grid = (int)(groups < (int64_t)sms * 2 ? groups : (int64_t)sms * 2);
  append_block_kernel<<<grid, 256, smem, st>>>(a);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}

extern "C" int64_t algp_append_work_doubles(int64_t n) { return n + 2; }

extern "C" int algp_append(double* Wt, int64_t ldw, int64_t ncols, const double* X, int64_t n, int d,
                           const double* log_ls_host, double log_os, int kind, double noise,
                           double* diagP, double* pi, uint8_t* is_static, const void* j_dev, double delta,
                           int mark_static, double* work, void* stream) {
  if (!Wt || !X || !diagP || !pi || !is_static || !j_dev || !work || ncols < 0 || ncols >= ldw || (ldw & 1) || !(delta > 0.0)) return ALGP_ERR_INVALID;
  AppendArgs a;
  int rc = make_kernel_params(&a.kp, d, log_ls_host, log_os, kind);
  if (rc) return rc;
  a.noise = noise; a.Wt = Wt; a.ldw = ldw; a.ncols = (int)ncols; a.X = X; a.n = n;
  a.diagP = diagP; a.pi = pi; a.is_static = is_static; a.j_dev = (const long long*)j_dev;
  a.delta = delta; a.mark_static = mark_static; a.scratch = work;
  if (n == 0) return ALGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  append_prepare_kernel<<<1, 1, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  append_column_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a);
  ALGP_LAUNCH_CHECK();
  append_commit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Wt, ldw, (int)ncols, work + 2, n);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
