// C-ABI glue: status strings, last CUDA error, hyper-parameter packing.
#include "common.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>

static thread_local char g_last_err[512] = "";

extern "C" int algp_set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_last_err, sizeof(g_last_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
  return ALGP_ERR_CUDA;
}

extern "C" const char* algp_last_cuda_error(void) { return g_last_err; }

extern "C" const char* algp_strerror(int code) {
  switch (code) {
    case ALGP_OK: return "ok";
    case ALGP_ERR_INVALID: return "invalid argument";
    case ALGP_ERR_CUDA: return "CUDA runtime error";
    case ALGP_ERR_NOT_PD: return "matrix is not positive definite";
    case ALGP_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" int algp_version(void) { return 100; }

// theta arrives as the reference stores it: logs (models.py:180, run.py:36-37)
int make_kernel_params(KernelParams* kp, int d, const double* log_ls_host, double log_os, int kind) {
  if (!kp || !log_ls_host || d < 1 || d > ALGP_MAX_D) return ALGP_ERR_INVALID;
  if (kind != 0 && kind != 1) return ALGP_ERR_UNSUPPORTED;   // models.py:226-227 raises NotImplementedError
  memset(kp, 0, sizeof(*kp));
  kp->kind = kind;
  kp->d = d;
  for (int j = 0; j < d; ++j) kp->inv_ls[j] = exp(-log_ls_host[j]);
  kp->outputscale = exp(log_os);
  return ALGP_OK;
}
