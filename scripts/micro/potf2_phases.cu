// Phase times of potf2inv_blocked_kernel (clock64 of thread 0 at the phase boundaries) and the dependent-issue
// latencies of the instructions on its column-to-column chain.  Build: see scripts/micro/build.sh.
#define PB_PROF
#include "potf2_blocked.cuh"
#include <stdio.h>
#include <vector>
#include <math.h>

extern "C" int algp_set_cuda_error(cudaError_t e, const char* file, int line) {
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line);
  return ALGP_ERR_CUDA;
}

template <int OP>
__global__ void lat_kernel(double* out, long long* cyc, double x0, int iters) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000000001;
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) x = fma(x, y, 1e-30);                       // DFMA chain
      if (OP == 1) x = x * y;                                  // DMUL chain
      if (OP == 2) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.0; }   // MUFU.RCP64H + DADD
      if (OP == 3) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);      // 64-bit shuffle chain
      if (OP == 4) { sm[threadIdx.x & 63] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 63]; __syncwarp(); }  // STS -> LDS
      if (OP == 5) { x = fma(x, y, 1e-30); __syncthreads(); }  // DFMA + CTA barrier
      if (OP == 6) x = 1.0 / x;                                // IEEE division
      if (OP == 7) x = pb_rcp(x) + 0.5;                        // pb_rcp + DADD
      if (OP == 8) x = x + y;                                  // DADD chain
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void lat(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  lat_kernel<OP><<<1, threads>>>(out, cyc, 1.5, iters);
  lat_kernel<OP><<<1, threads>>>(out, cyc, 1.5, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %3d threads: %7.1f cycles per op\n", name, threads, (double)h / (iters * 8.0));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {32, 256}) {
    lat<0>("DFMA dependent chain", threads);
    lat<1>("DMUL dependent chain", threads);
    lat<8>("DADD dependent chain", threads);
    lat<2>("MUFU.RCP64H + DADD", threads);
    lat<7>("pb_rcp (MUFU + 3 DFMA) + DADD", threads);
    lat<6>("IEEE 1.0 / x", threads);
    lat<3>("64-bit SHFL chain", threads);
    lat<4>("STS, syncwarp, LDS, syncwarp", threads);
    lat<5>("DFMA + __syncthreads", threads);
  }
  // ---- the blocked kernel on a well-conditioned SPD block ----
  const int n = 128;
  std::vector<double> A(n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[i * n + j] = exp(-0.02 * (i - j) * (i - j)) + (i == j ? 0.05 : 0.0);
  double *dA, *dL; int* info;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&info, 4);
  cudaFuncSetAttribute(potf2inv_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_BYTES);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaMemset(info, 0, 4);
    potf2inv_blocked_kernel<<<1, 256, PB_SMEM_BYTES>>>(dA, n, dL, n, 0, info);
    cudaDeviceSynchronize();
  }
  long long prof[64];
  cudaMemcpyFromSymbol(prof, g_pb_prof, sizeof(prof));
  printf("potf2inv_blocked_kernel, thread 0 clock64 deltas (cycles): panel: publish+barrier | eliminate | barrier | update(+next publish start)\n");
  for (int p = 0; p < 8; ++p) {
    long long a = prof[4 * p], b = prof[4 * p + 1], c = prof[4 * p + 2], d = prof[4 * p + 3];
    long long nxt = (p < 7) ? prof[4 * p + 4] : c;
    printf("  panel %d: %6lld | %6lld | %6lld | %6lld\n", p, b - a, c - b, (p < 7) ? d - c : 0, (p < 7) ? nxt - d : 0);
  }
  printf("  total %lld cycles from first publish to end of last elimination\n", prof[30] - prof[0]);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
