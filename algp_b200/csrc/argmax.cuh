// np.argmax on the device (first maximum wins: reference agent.py:349,402), shared by score.cu (single GPU) and
// p2p.cu (winner exchange between the GPUs of one box).
#pragma once
#include "common.cuh"
#include <math.h>

struct ArgPair { double v; long long i; };

__device__ __forceinline__ bool arg_better(double v, long long i, double bv, long long bi) {
  return (v > bv) || (v == bv && i < bi);
}

static __global__ void argmax_stage1_kernel(const double* __restrict__ x, int64_t n, int64_t idx_offset, ArgPair* __restrict__ part) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = x[i];
    if (arg_better(v, i + idx_offset, bv, bi)) { bv = v; bi = i + idx_offset; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (arg_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
    part[blockIdx.x].v = bv;
    part[blockIdx.x].i = bi;
  }
}


#define ARGMAX_BLOCKS 148
#define ARGMAX_ONE_CTA_MAX (1 << 14)   // up to this many values one 1024-thread CTA scans the vector itself (65536 values: 37 us in one CTA, 10 us in two kernels)

// block-wide first-maximum of (bv, bi) over a 1024-thread CTA; the result is valid in every thread of warp 0
__device__ __forceinline__ void argmax_block_reduce_1024(double& bv, long long& bi) {
  __shared__ double sv1k[32];
  __shared__ long long si1k[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv1k[threadIdx.x >> 5] = bv; si1k[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = sv1k[threadIdx.x];
    bi = si1k[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  }
}

// scan x[n] (global ids = position + idx_offset) with the whole CTA: the first-maximum pair of this thread
__device__ __forceinline__ void argmax_scan(const double* __restrict__ x, int64_t n, int64_t idx_offset, double& bv, long long& bi) {
  bv = -INFINITY;
  bi = 0x7fffffffffffffffLL;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    if (arg_better(v, i + idx_offset, bv, bi)) { bv = v; bi = i + idx_offset; }
  }
}
