// Micro-benchmark: cycles per iteration of the per-chunk control steps of the digit GEMM's single-lane roles, one warp alone
// on its SM: (0) empty loop, (1) ballot + ffs + redux.sync broadcast of a plan word, (2) try_wait on a barrier phase that has
// already completed, (3) tcgen05.fence::after_thread_sync, (4) elect.sync + tcgen05.commit to a barrier nobody waits on,
// (5) elect.sync + plain mbarrier.arrive, (6) 1 + 2 + 3 + 4 together (the issuer's skeleton), (7) a producer/consumer ping-pong
// between two warps through a 4-deep full/empty ring where the consumer releases with tcgen05.commit, (8) the same releasing
// with a plain mbarrier.arrive.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o control_costs control_costs.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wait_parity(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void commit_to(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}

__global__ void __launch_bounds__(64, 1) bench(int mode, int iters, const uint32_t* __restrict__ words, long long* out) {
  __shared__ uint64_t done_bar, sink_bar, full_bar[4], empty_bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&done_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&sink_bar)));
    for (int s = 0; s < 4; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&full_bar[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&empty_bar[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (tid == 0) arrive(&done_bar);            // phase 0 of done_bar is complete from here on
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (mode <= 6) {
    if (warp == 0) {
      for (int it = 0; it < iters; ++it) {
        if (mode == 1 || mode == 6) {
          const uint32_t w = words[(it * 32 + lane) & 1023];
          uint32_t bits = __ballot_sync(0xffffffffu, w != 0u);
          const int l = __ffs(bits | 0x80000000u) - 1;
          const uint32_t ws = __reduce_or_sync(0xffffffffu, lane == l ? w : 0u);
          acc += (ws & 0xf) + ((ws >> 4) & 0xf);
        }
        if (mode == 2 || mode == 6) wait_parity(&done_bar, 0);
        if (mode == 3 || mode == 6) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode == 4 || mode == 6) { if (elect_one()) commit_to(&sink_bar); }
        if (mode == 5) { if (elect_one()) arrive(&sink_bar); }
        if (mode == 0) asm volatile("" ::"r"(it) : "memory");
      }
    }
  } else {
    // ping-pong: warp 0 = producer (waits empty, arrives full), warp 1 = consumer (waits full, releases empty)
    if (warp == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it & 3, u = it >> 2;
        if (u > 0) wait_parity(&empty_bar[s], (u - 1) & 1);
        if (elect_one()) arrive(&full_bar[s]);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        const int s = it & 3, u = it >> 2;
        wait_parity(&full_bar[s], u & 1);
        if (mode == 7) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) { if (mode == 7) commit_to(&empty_bar[s]); else arrive(&empty_bar[s]); }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(slot));
}

int main() {
  long long* d; uint32_t* w;
  cudaMalloc(&d, 16); cudaMalloc(&w, 4096);
  uint32_t h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = (i % 3 == 0) ? 0u : (0x80000000u | (i & 7) | ((i >> 3 & 3) << 4) | (6u << 8));
  cudaMemcpy(w, h, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"empty loop", "ballot + ffs + redux.sync plan broadcast (+ L1-resident word load)", "try_wait on a completed barrier phase",
                         "tcgen05.fence::after_thread_sync", "elect + tcgen05.commit", "elect + mbarrier.arrive",
                         "issuer skeleton (1 + 2 + 3 + 4)", "4-deep ping-pong, consumer releases with tcgen05.commit",
                         "4-deep ping-pong, consumer releases with mbarrier.arrive"};
  const int iters = 20000;
  for (int mode = 0; mode <= 8; ++mode) {
    long long r[2];
    bench<<<148, 64>>>(mode, iters, w, d);
    bench<<<148, 64>>>(mode, iters, w, d);
    cudaError_t e = cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
    printf("%-72s %7.1f cycles / iteration %s\n", names[mode], (double)r[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
