// Shared device/host helpers for the algp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ALGP_MAX_D 16          // ARD input dims (reference default field is d=6, utils.py:143-155)
#define ALGP_BLK 128           // every factor-side matrix dimension is padded to a multiple of this

// status codes returned by every C-ABI entry point (include/algp_b200.h)
enum {
  ALGP_OK = 0,
  ALGP_ERR_INVALID = 1,        // bad argument (null pointer, d > ALGP_MAX_D, unpadded dimension ...)
  ALGP_ERR_CUDA = 2,           // a CUDA runtime call failed; see algp_last_cuda_error()
  ALGP_ERR_NOT_PD = 3,         // matrix not positive definite (info > 0, LAPACK potrf convention)
  ALGP_ERR_UNSUPPORTED = 4,
};

extern "C" int algp_set_cuda_error(cudaError_t e, const char* file, int line);

#define ALGP_CUDA(call)                                                    \
  do {                                                                     \
    cudaError_t _e = (call);                                               \
    if (_e != cudaSuccess) return algp_set_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)

#define ALGP_LAUNCH_CHECK() ALGP_CUDA(cudaGetLastError())

struct KernelParams {
  int kind;                    // 0 = RBF, 1 = Matern nu=1.5
  int d;
  double inv_ls[ALGP_MAX_D];   // 1 / lengthscale_j
  double outputscale;          // s^2
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA.  On sm_100a every f64 mma.sync shape lowers to
// DMMA.8x8x4 (checked with cuobjdump), so m8n8k4 is the native unit.
// Fragments: a = A[lane>>2][lane&3], b = B[k=lane&3][n=lane>>2],
// c{0,1} = C[lane>>2][2*(lane&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// k(x,x') for already length-scaled squared distance r2
__device__ __forceinline__ double kern_from_r2(double r2, int kind, double os) {
  if (kind == 0) return os * exp(-0.5 * r2);
  const double s3 = 1.7320508075688772;
  double r = sqrt(r2);
  return os * (1.0 + s3 * r) * exp(-s3 * r);
}
__device__ __forceinline__ float kern_from_r2f(float r2, int kind, float os) {
  if (kind == 0) return os * expf(-0.5f * r2);
  const float s3 = 1.7320508f;
  float r = sqrtf(r2);
  return os * (1.0f + s3 * r) * expf(-s3 * r);
}
