// Micro-benchmark: what one SPARSE chunk of the INT8 digit GEMM costs on the tensor side.  One elected thread issues,
// per chunk, the MMAs of an (L, W) schedule (L A planes against up to W B planes: N = 64 min(L - i, W), split at 256)
// from a 4-stage operand ring in shared memory, then -- optionally -- one tcgen05.commit to that stage's mbarrier, as
// the kernel does to release the stage.  No loads, nobody waits on the barriers: only issue + tensor pipe + commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_i8_chunks mma_i8_chunks.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define STAGE_BYTES 43008      // 7 planes x (4 KB + 2 KB)
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((128u >> 4) << 16); }
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}"
               ::"r"(d), "r"(a), "r"(b), "r"(id), "r"((256u >> 4) | (1u << 14)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int L, int W, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) bench(int chunks, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 4 * STAGE_BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar[s])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&done)));
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
  if (warp == 1) {
    const long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      const int s = c & 3;
      const uint32_t a0 = s_u32(smem) + (uint32_t)s * STAGE_BYTES, b0 = a0 + 7 * 4096;
      const uint32_t da = desc_lo(a0), db = desc_lo(b0);
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < L; ++i) {
          const int n = (L - i < W) ? L - i : W;
          const int c0 = n < 4 ? n : 4;
          mma(tmem + (uint32_t)(i * 64), da + (uint32_t)(i * 256), db, idesc_i8(128, c0 * 64));
          if (n > 4) mma(tmem + (uint32_t)((i + 4) * 64), da + (uint32_t)(i * 256), db + 512u, idesc_i8(128, (n - 4) * 64));
        }
        if (COMMIT_EVERY > 0 && (c % COMMIT_EVERY) == COMMIT_EVERY - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar[s])) : "memory");
      }
    }
    if (elect_one()) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&done)) : "memory");
    }
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s_u32(&done)) : "memory");
    } while (!ok);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && (tid & 31) == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int L, int W>
static void run(long long* d) {
  const int chunks = 4096;
  double floor_ = 0;
  int nm = 0;
  for (int i = 0; i < L; ++i) {
    int n = (L - i < W) ? L - i : W;
    int c0 = n < 4 ? n : 4;
    floor_ += c0 * 32 < 60 ? 60 : c0 * 32; ++nm;
    if (n > 4) { floor_ += (n - 4) * 32 < 60 ? 60 : (n - 4) * 32; ++nm; }
  }
  long long h[3] = {0, 0, 0};
  const size_t smem = 4 * STAGE_BYTES + 1024;
  cudaFuncSetAttribute(bench<L, W, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<L, W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<L, W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    bench<L, W, 0><<<148, 128, smem>>>(chunks, d); cudaMemcpy(&h[0], d, 8, cudaMemcpyDeviceToHost);
    bench<L, W, 1><<<148, 128, smem>>>(chunks, d); cudaMemcpy(&h[1], d, 8, cudaMemcpyDeviceToHost);
    bench<L, W, 4><<<148, 128, smem>>>(chunks, d); cudaMemcpy(&h[2], d, 8, cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("L=%d W=%d (%2d MMAs, model %4.0f cycles): no commit %6.1f, commit per chunk %6.1f, commit per 4 chunks %6.1f cycles / chunk  %s\n",
         L, W, nm, floor_, (double)h[0] / chunks, (double)h[1] / chunks, (double)h[2] / chunks, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  printf("cycles per chunk of tcgen05.mma.kind::i8 (M128 K32) schedules, one elected thread, 1 CTA/SM on all SMs\n");
  run<1, 1>(d); run<2, 1>(d); run<2, 2>(d); run<3, 3>(d); run<4, 2>(d); run<4, 4>(d); run<5, 5>(d); run<7, 3>(d); run<7, 7>(d);
  return 0;
}
