// Measured roofs for the kernels that are bound by neither HBM nor a tensor pipe (diagnostic entry points used by
// bench.py; nothing on the product path calls them).
//
// algp_probe_l2_read: every CTA streams a DIFFERENT part of an L2-sized buffer with the load instruction of the
// scoring kernels (16-byte ld.global.nc.L1::no_allocate), `passes` times, so after the first pass every request is an
// L2 hit and no two CTAs ask for the same line at the same time (no request merging in the L2).  bytes x passes /
// time is the L2 -> SM delivery rate the row-streaming scoring kernel (score.cu) is measured against.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) probe_l2_read_kernel(const double2* __restrict__ buf, int64_t n16, int passes,
                                                            double* __restrict__ sink) {
  const int64_t per_cta = n16 / gridDim.x;                 // 16-byte elements owned by a CTA per pass
  const int64_t start = (int64_t)blockIdx.x * per_cta;
  double acc = 0.0;
  for (int p = 0; p < passes; ++p) {
    // rotate the CTA's window by a prime number of CTAs per pass: the data stays L2-resident, the address stream of
    // an SM differs from pass to pass
    const int64_t base = (start + (int64_t)p * 37 * per_cta) % (per_cta * gridDim.x);
    for (int64_t i = threadIdx.x; i < per_cta; i += 256 * 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = make_double2(0.0, 0.0);
        const int64_t j = i + 256 * u;
        if (j < per_cta) {
          const double2* p2 = buf + base + j;
          asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(p2));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    }
  }
  if (acc == 1.2345e-300) sink[0] = acc;                   // keeps the loads alive
}

}  // namespace

// Launch the probe once: `bytes` (multiple of 16 x grid) of buf are read `passes` times by sms x ctas_per_sm CTAs.
extern "C" int algp_probe_l2_read(const void* buf, int64_t bytes, int passes, int ctas_per_sm, void* sink, void* stream) {
  if (!buf || !sink || bytes < (1 << 20) || passes < 1 || ctas_per_sm < 1 || ctas_per_sm > 8 || ((uintptr_t)buf & 15))
    return ALGP_ERR_INVALID;
  int dev = 0, sms = 148;
  ALGP_CUDA(cudaGetDevice(&dev));
  ALGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * ctas_per_sm;
  const int64_t n16 = bytes / 16 / grid * grid;
  probe_l2_read_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2*)buf, n16, passes, (double*)sink);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
