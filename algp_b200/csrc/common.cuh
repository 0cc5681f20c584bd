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

// Per-device bookkeeping for cudaFuncSetAttribute (the attribute belongs to one device context, so a process-wide
// "already configured" flag is wrong as soon as one process drives two GPUs).  Atomic, so concurrent host threads at
// worst set the attribute twice.
#include <atomic>
#define ALGP_MAX_DEVICES 64
struct AlgpPerDevice {
  std::atomic<size_t> v[ALGP_MAX_DEVICES];
  // true if `want` exceeds what this device was configured with so far (and records it)
  bool raise(size_t want) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ALGP_MAX_DEVICES) return true;
    size_t cur = v[dev].load(std::memory_order_relaxed);
    while (want > cur)
      if (v[dev].compare_exchange_weak(cur, want, std::memory_order_relaxed)) return true;
    return false;
  }
};

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

// ---- mbarrier (split-phase CTA barrier) ------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbarrier_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(b)) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_addr_u32(b);
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
  } while (!ok);
}

// one non-blocking poll of a phase (true once the phase with this parity has completed)
__device__ __forceinline__ bool mbarrier_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_addr_u32(b)), "r"(parity)
               : "memory");
  return ok != 0;
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

// exp(x) for x <= 0, branch-free: k = rint(x log2 e), r = x - k ln2 (two-part Cody-Waite),
// degree-13 Taylor polynomial on |r| <= 0.347 (truncation 4e-18), scaled by 2^k through the
// exponent field.  ~18 fp64 ops against ~30 plus a special-case branch for exp(); error < 2 ulp.
// Arguments below -708 (result < 3.4e-308) are clamped, so the result is never denormal.
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, -708.0);
  const double magic = 6755399441055744.0;   // 1.5 * 2^52
  double kf = fma(x, 1.4426950408889634, magic);
  const int k = __double2loint(kf);
  kf -= magic;
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;          // 1/13!
  p = fma(p, r, 2.08767569878681e-09);        // 1/12!
  p = fma(p, r, 2.505210838544172e-08);       // 1/11!
  p = fma(p, r, 2.755731922398589e-07);       // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);      // 1/9!
  p = fma(p, r, 2.48015873015873e-05);        // 1/8!
  p = fma(p, r, 1.984126984126984e-04);       // 1/7!
  p = fma(p, r, 1.388888888888889e-03);       // 1/6!
  p = fma(p, r, 8.333333333333333e-03);       // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);      // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);      // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// exp(x) * tab-scale for x <= 0 with a 16-entry table tab[j] = scale * 2^(j/16) (shared memory):
// n = rint(16 x / ln2), r = x - n ln2/16 (|r| <= 0.0217, two-part Cody-Waite), degree-7 Taylor
// polynomial (truncation 1.2e-18), result = tab[n & 15] * p(r) * 2^(n >> 4).  11 fp64 ops; 1.5 ulp.
// x is clamped at -650 through its high word (x <= 0: more negative = larger unsigned high word) so
// the exponent-field add never leaves the normal range for any sane scale.
__device__ __forceinline__ double exp_nonpos_tab(double x, const double* __restrict__ tab) {
  if ((unsigned)__double2hiint(x) > 0xC0845000u) x = -650.0;
  const double magic = 6755399441055744.0;   // 1.5 * 2^52
  double nf = fma(x, 23.083120654223414, magic);
  const int n = __double2loint(nf);
  nf -= magic;
  double r = fma(nf, -0.04332169877307024, x);
  r = fma(nf, -1.1926343307941173e-11, r);
  double p = 1.984126984126984e-04;           // 1/7!
  p = fma(p, r, 1.388888888888889e-03);
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= tab[n & 15];
  return __hiloint2double(__double2hiint(p) + ((n >> 4) << 20), __double2loint(p));
}

// 2^(j/16), j = 0..15 (correctly rounded)
static __device__ __constant__ double c_exp2_16th[16] = {
    1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484,
    1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407,
    1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474};

// k(x,x') for already length-scaled squared distance r2
__device__ __forceinline__ double kern_from_r2(double r2, int kind, double os) {
  if (kind == 0) return os * exp_nonpos(-0.5 * r2);
  const double s3 = 1.7320508075688772;
  double r = sqrt(r2);
  return os * (1.0 + s3 * r) * exp_nonpos(-s3 * r);
}
__device__ __forceinline__ float kern_from_r2f(float r2, int kind, float os) {
  if (kind == 0) return os * expf(-0.5f * r2);
  const float s3 = 1.7320508f;
  float r = sqrtf(r2);
  return os * (1.0f + s3 * r) * expf(-s3 * r);
}
