// FP64 pipe ceilings on this GPU: DFMA rate, exp_nonpos rate, DMMA rate (register operands only).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../algp_b200/csrc/common.cuh"
extern "C" int algp_set_cuda_error(cudaError_t, const char*, int) { return 2; }

__global__ void dfma_kernel(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void exp_kernel(double* out, int iters) {
  double x = -(threadIdx.x * 1e-2 + blockIdx.x * 1e-4), s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s += exp_nonpos(x - 0.37 * i - 1e-3 * it);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double* out, int iters) {
  double c0[8], c1[8];
  for (int i = 0; i < 8; ++i) c0[i] = c1[i] = 0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], a, b);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
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
  const int iters = 20000;
  for (int bps : {1, 2, 4, 8}) {
    int grid = sms * bps;
    float t1 = timeit([&] { dfma_kernel<<<grid, 256>>>(out, iters); });
    double fma = (double)grid * 256 * iters * 8;
    float t2 = timeit([&] { exp_kernel<<<grid, 256>>>(out, iters / 8); });
    double ex = (double)grid * 256 * (iters / 8) * 4;
    float t3 = timeit([&] { dmma_kernel<<<grid, 256>>>(out, iters / 4); });
    double mm = (double)grid * 8 * (iters / 4) * 8 * 256;      // warps * iters * 8 dmma * 256 fma
    printf("blocks/SM %d: DFMA %.2f TFLOP/s (%.1f FMA/clk/SM @1.965GHz) | exp_nonpos %.1f Gexp/s | DMMA %.2f TFLOP/s\n", bps,
           2 * fma / t1 / 1e9, fma / (t1 * 1e-3) / sms / 1.965e9, ex / t2 / 1e6, 2 * mm / t3 / 1e9);
  }
  return 0;
}
