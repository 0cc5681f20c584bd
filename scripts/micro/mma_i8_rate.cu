// Micro-benchmark: issue rate of small tcgen05.mma.kind::i8 instructions (M128, N = 64..256, K32) from one thread,
// as a function of how consecutive instructions' accumulator ranges in TMEM relate:
//   pattern 0: the same D range every time (a GEMM k-loop)
//   pattern 1: D shifted by 64 columns each time, wrapping over 7 groups (what a sparse chunk of the digit GEMM issues)
//   pattern 2: disjoint D ranges cycling over the TMEM columns
// Operands are whatever shared memory holds (all zero): only timing matters.   nvcc -arch=sm_100a -o mma_i8_rate mma_i8_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int pattern, int nplanes, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t da = desc_kmajor(s_u32(smem)), db = desc_kmajor(s_u32(smem) + 16384);
    const uint32_t id = idesc_i8(128, 64 * nplanes);
    const long long t0 = clock64();
    // the column advances incrementally (no division in the issue loop: a lone thread runs ~20 cycles per dependent
    // instruction, so any scalar work per MMA shows up in the rate)
    const uint32_t step = pattern == 0 ? 0u : (pattern == 1 ? 64u : 64u * nplanes);
    const uint32_t limit = pattern == 1 ? (uint32_t)((8 - nplanes) * 64) : (uint32_t)((512 / (64 * nplanes)) * 64 * nplanes);
    uint32_t col = 0;
#pragma unroll 4
    for (int r = 0; r < reps; ++r) {
      mma(tmem + col, da, db, id);
      col += step;
      col = col >= limit ? 0u : col;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar)) : "memory");
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s_u32(&bar)) : "memory");
    } while (!ok);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 4096;
  printf("cycles per tcgen05.mma.kind::i8 (M128, K32), one issuing thread, 1 CTA/SM on all SMs; floor = N/2 cycles\n");
  for (int np = 1; np <= 4; ++np)
    for (int pat = 0; pat < 3; ++pat) {
      bench<<<148, 128, 64 * 1024>>>(pat, np, reps, d);
      bench<<<148, 128, 64 * 1024>>>(pat, np, reps, d);
      long long h = 0;
      cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("N=%3d pattern %d (%s): %.1f cycles / MMA\n", 64 * np, pat,
             pat == 0 ? "same D" : (pat == 1 ? "D shifted by 64" : "disjoint D"), (double)h / reps);
    }
  return 0;
}
