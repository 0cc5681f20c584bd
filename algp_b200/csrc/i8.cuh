// Internal interface of the INT8 digit GEMM (i8.cu), shared with the recursive factorisation (i8chol.cu).
#pragma once
#include "common.cuh"

#define I8_TM 128                       // rows of the left operand per tile
#define I8_TN 64                        // rows of the right operand (= output columns) per tile
#define I8_KC 32                        // bytes (= int8 digits) of k per row per stage = one MMA's K
#define I8_MAX_S 8

enum { I8_K_FULL = 0, I8_KE_NT = 1, I8_KE_MT = 2, I8_KB_NT = 1 };

struct I8Gemm {
  int MT, NT;                          // 128-row tiles of A, 64-row tiles of B
  const int8_t* a_tiles;               // [MT][kchunks][S][128 x 32 B]
  const int8_t* b_tiles;               // [NT][kchunks][S][ 64 x 32 B]
  int kchunks;                         // K / 32
  const double* scale_a;               // [MT x 128]  2^eA_m
  const double* scale_b;               // [NT x 64]   2^eB_j
  int kbeg_rule, kend_rule;            // triangular operands: I8_KB_NT (k >= j), I8_KE_NT (k <= j), I8_KE_MT (k <= i)
  int lower_only;                      // skip tiles entirely above the diagonal (square outputs)
  int mt_desc, nt_desc;                // tile order: longest k-ranges first
  double* rn_partial; int rn_nt;       // MODE 0: [MT x 128][rn_nt] row sums of squares per 64-column tile
  double* C; int64_t ldc;              // MODE 1: C = alpha A B^T + beta C
  double alpha, beta;
  int transposed;                      // MODE 1: store C^T (ldc = row stride of the transposed matrix)
  const uint8_t* a_mask;               // [MT][mask_ld] plane-occupancy byte per 32-byte k chunk (NULL: all occupied)
  const uint8_t* b_mask;               // [NT][mask_ld]
  int64_t mask_ld;                     // i8_mask_ld(K)
};

inline int64_t i8_mask_ld(int64_t cols) { return (cols / I8_KC + 7) / 8 * 8; }

inline I8Gemm i8_gemm_default() {
  I8Gemm g;
  g.MT = g.NT = 0; g.a_tiles = g.b_tiles = nullptr; g.kchunks = 0; g.scale_a = g.scale_b = nullptr;
  g.kbeg_rule = g.kend_rule = I8_K_FULL; g.lower_only = 0; g.mt_desc = g.nt_desc = 0;
  g.rn_partial = nullptr; g.rn_nt = 0; g.C = nullptr; g.ldc = 0; g.alpha = 1.0; g.beta = 0.0; g.transposed = 0;
  g.a_mask = g.b_mask = nullptr; g.mask_ld = 0;
  return g;
}

int i8_split(const double* src, int64_t rows, int64_t cols, int64_t ld, int nslices, int tile_rows, int8_t* planes,
             double* row_scale, uint8_t* mask, cudaStream_t st);
int i8_gemm(const I8Gemm& a, int nslices, cudaStream_t st);
