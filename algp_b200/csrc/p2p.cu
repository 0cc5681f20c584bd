// Winner exchange between the GPUs of one box over NVLink peer memory.
//
// Candidate scoring shards over the GPUs; the only exchange per scoring step is one 16-byte {score, global index}
// pair per rank.  Through NCCL that is an all-gather launch (~20-45 us on an 8-GPU box, round 1) after the two argmax
// kernels -- a fifth of a strong-scaled step of 8192 candidate sets.  Here the LAST kernel of the argmax writes the
// rank's pair straight into a mailbox in every peer's memory (cudaIpc-mapped, plain stores over NVLink), waits for the
// peers' pairs to land in its own mailbox and reduces them with np.argmax's first-maximum rule (agent.py:349,402:
// highest score, ties to the lowest global index): no collective launch, one kernel.
//
// Mailbox of rank r (device memory of GPU r, algp_p2p_mailbox_bytes(world) bytes, zero-initialised):
//     slot[parity][src] = {double v; int64 i; uint64 epoch; pad}     parity = epoch & 1, src = writing rank
// A writer stores the payload, fences (membar.sys) and stores the epoch; the reader polls the epoch with acquire loads.
// Two parities suffice: a rank can only reach epoch e + 2 after every peer has sent its e + 1 pair, i.e. after every
// peer has finished reading epoch e.  Epochs start at 1 and advance by one per call on every rank.
#include "argmax.cuh"
#include <string.h>

#define P2P_MAX_WORLD 16

struct P2PSlot { double v; long long i; unsigned long long epoch; unsigned long long pad; };
struct P2PMailbox { P2PSlot slot[2][P2P_MAX_WORLD]; };

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// out3 = {double value; int64 index; int64 status}: status 0 = ok, 1 = a peer's pair did not arrive in time.
// DIRECT: the (1024-thread) CTA scans this rank's score block x[n] itself -- one launch for argmax + exchange;
// otherwise one warp reduces the np block partials of argmax_stage1_kernel.
template <bool DIRECT>
__global__ void argmax_exchange_kernel(const ArgPair* __restrict__ part, int np, const double* __restrict__ x, int64_t n,
                                       int64_t idx_offset, P2PMailbox* const* __restrict__ peers, int rank, int world,
                                       unsigned long long epoch, long long timeout_cycles, long long* __restrict__ out3) {
  const int lane = threadIdx.x;
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  if (DIRECT) {
    argmax_scan(x, n, idx_offset, bv, bi);
    argmax_block_reduce_1024(bv, bi);
    if (threadIdx.x >= 32) return;                       // warp 0 carries the exchange
  } else {
    for (int p = lane; p < np; p += 32)
      if (arg_better(part[p].v, part[p].i, bv, bi)) { bv = part[p].v; bi = part[p].i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (arg_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  }
  // every lane holds this rank's winner; lane r delivers it to rank r (its own mailbox included)
  const int par = (int)(epoch & 1ull);
  if (lane < world) {
    P2PSlot* dst = &peers[lane]->slot[par][rank];
    *reinterpret_cast<volatile double*>(&dst->v) = bv;
    *reinterpret_cast<volatile long long*>(&dst->i) = bi;
    __threadfence_system();
    st_release_sys(&dst->epoch, epoch);
  }
  // lane r waits for rank r's pair in the local mailbox
  double v = -INFINITY;
  long long i = 0x7fffffffffffffffLL;
  int bad = 0;
  if (lane < world) {
    P2PSlot* src = &peers[rank]->slot[par][lane];
    const long long t0 = clock64();
    while (ld_acquire_sys(&src->epoch) != epoch) {
      if (clock64() - t0 > timeout_cycles) { bad = 1; break; }
    }
    if (!bad) {
      v = *reinterpret_cast<volatile double*>(&src->v);
      i = *reinterpret_cast<volatile long long*>(&src->i);
    }
  }
  bad = __any_sync(0xffffffffu, bad);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (arg_better(ov, oi, v, i)) { v = ov; i = oi; }
  }
  if (lane == 0) {
    out3[0] = __double_as_longlong(v);
    out3[1] = i;
    out3[2] = bad;
  }
}

}  // namespace

extern "C" int64_t algp_p2p_mailbox_bytes(int world) {
  return (world >= 1 && world <= P2P_MAX_WORLD) ? (int64_t)sizeof(P2PMailbox) : 0;
}

// Allocate this rank's zeroed mailbox with cudaMalloc (cudaIpc needs a whole allocation, which a caching allocator
// does not hand out) and export its 64-byte IPC handle.
extern "C" int algp_p2p_create(int64_t bytes, void** local_ptr, void* handle64) {
  if (bytes <= 0 || !local_ptr || !handle64) return ALGP_ERR_INVALID;
  void* p = nullptr;
  ALGP_CUDA(cudaMalloc(&p, (size_t)bytes));
  ALGP_CUDA(cudaMemset(p, 0, (size_t)bytes));
  ALGP_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return algp_set_cuda_error(e, __FILE__, __LINE__);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  *local_ptr = p;
  return ALGP_OK;
}

// Map a peer's mailbox into this process (peer access over NVLink is enabled lazily by the runtime).
extern "C" int algp_p2p_open(const void* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) return ALGP_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  ALGP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return ALGP_OK;
}

extern "C" int algp_p2p_close(void* peer_ptr) {
  if (!peer_ptr) return ALGP_ERR_INVALID;
  ALGP_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return ALGP_OK;
}

extern "C" int algp_p2p_destroy(void* local_ptr) {
  if (!local_ptr) return ALGP_ERR_INVALID;
  ALGP_CUDA(cudaFree(local_ptr));
  return ALGP_OK;
}

// np.argmax of this rank's block x[n] (global ids = position + idx_offset; n = 0: an empty shard that only takes
// part in the exchange), exchanged with the `world` ranks whose mailboxes are listed in peers_dev (device array of
// `world` pointers, entry r = rank r's mailbox as mapped into this process), reduced with the first-maximum rule.
// out3 (device, 24 bytes) = {double value; int64 index; int64 status}; identical on every rank.  `epoch` must be
// 1, 2, 3, ... in step on all ranks.  work: algp_argmax_work_bytes() bytes.  timeout_ms bounds the wait for the peers.
extern "C" int algp_argmax_exchange(const double* x, int64_t n, int64_t idx_offset, void* work, const void* peers_dev,
                                    int rank, int world, int64_t epoch, double timeout_ms, void* out3, void* stream) {
  if ((n > 0 && !x) || n < 0 || !work || !peers_dev || !out3 || world < 1 || world > P2P_MAX_WORLD || rank < 0 ||
      rank >= world || epoch < 1)
    return ALGP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = 0;
  const bool direct = n > 0 && n <= ARGMAX_ONE_CTA_MAX;
  if (n > 0 && !direct) {
    blocks = (int)((n + 255) / 256 < ARGMAX_BLOCKS ? (n + 255) / 256 : ARGMAX_BLOCKS);
    argmax_stage1_kernel<<<blocks, 256, 0, st>>>(x, n, idx_offset, (ArgPair*)work);
    ALGP_LAUNCH_CHECK();
  }
  // the clock-rate attribute is a slow driver query (~1 ms): read it once per device
  static int khz_of[64] = {0};
  int dev = 0;
  ALGP_CUDA(cudaGetDevice(&dev));
  int khz = (dev >= 0 && dev < 64) ? khz_of[dev] : 0;
  if (khz == 0) {
    khz = 1965000;
    ALGP_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    if (dev >= 0 && dev < 64) khz_of[dev] = khz;
  }
  const long long cycles = (long long)(timeout_ms * (double)khz);
  if (direct)
    argmax_exchange_kernel<true><<<1, 1024, 0, st>>>((const ArgPair*)work, 0, x, n, idx_offset, (P2PMailbox* const*)peers_dev,
                                                     rank, world, (unsigned long long)epoch, cycles, (long long*)out3);
  else
    argmax_exchange_kernel<false><<<1, 32, 0, st>>>((const ArgPair*)work, blocks, nullptr, 0, 0, (P2PMailbox* const*)peers_dev,
                                                    rank, world, (unsigned long long)epoch, cycles, (long long*)out3);
  ALGP_LAUNCH_CHECK();
  return ALGP_OK;
}
