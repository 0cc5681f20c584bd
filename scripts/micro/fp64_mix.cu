// Do the DMMA (tensor sub-pipe) and DFMA (fp64 pipe) issue concurrently on sm_100a, or do they share the same units?
// Three kernels over the same grid: DMMA only, DFMA only, and a mix with equal flops of each per thread
// (1 DMMA = 256 FMA per warp = 8 DFMA warp instructions), both interleaved inside a warp and split over warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../algp_b200/csrc/common.cuh"
extern "C" int algp_set_cuda_error(cudaError_t, const char*, int) { return 2; }

template <int NM, int NF, int SPLIT>
__global__ void mix_kernel(double* out, int iters) {
  double c0[4], c1[4], f[16];
  for (int i = 0; i < 4; ++i) c0[i] = c1[i] = 0;
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  const double m = 1.0000001, c = 1e-9;
  const int warp = threadIdx.x >> 5;
  const bool do_m = SPLIT ? (warp & 1) == 0 : true, do_f = SPLIT ? (warp & 1) == 1 : true;
  for (int it = 0; it < iters; ++it) {
    if (NM > 0 && do_m) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma884(c0[i & 3], c1[i & 3], a, b);
    }
    if (NF > 0 && do_f) {
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i & 15] = fma(f[i & 15], m, c);
    }
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 4000;
  for (int bps : {2, 4}) {
    int grid = sms * bps;
    double warps = (double)grid * 8;
    float tm = timeit([&] { mix_kernel<4, 0, 0><<<grid, 256>>>(out, iters); });
    float tf = timeit([&] { mix_kernel<0, 32, 0><<<grid, 256>>>(out, iters); });
    float tx = timeit([&] { mix_kernel<4, 32, 0><<<grid, 256>>>(out, iters); });
    float ts = timeit([&] { mix_kernel<4, 32, 1><<<grid, 256>>>(out, iters); });
    float tq = timeit([&] { mix_kernel<4, 16, 0><<<grid, 256>>>(out, iters); });
    double fm = warps * iters * 4 * 256 * 2, ff = warps * iters * 32 * 32 * 2;
    printf("blocks/SM %d: DMMA alone %.2f TF | DFMA alone %.2f TF | interleaved 1:1 flops %.2f TF (%.3f ms vs %.3f + %.3f) | "
           "split over warps %.2f TF | interleaved 2:1 %.2f TF\n", bps, fm / tm / 1e9, ff / tf / 1e9, (fm + ff) / tx / 1e9, tx, tm, tf,
           (fm + ff) / 2 / ts / 1e9, (fm + ff / 2) / tq / 1e9);
  }
  return 0;
}
