// One-shot all-reduce (sum) of a SMALL fp32 vector over peer memory — the sync-BatchNorm statistics exchange (SURVEY.md 8e C2:
// keras BatchNormalization(synchronized=True), reference conformer.py:326-333 / subsampling.py norms=["batch", ...]; here fused.bn_train_fwd /
// bn_train_bwd).  36 of these sit on the critical path of every Conformer-S training step (18 BatchNorms, forward + backward), 2 x C
// floats each: an NCCL all-reduce costs its launch + protocol latency per call; this kernel is one launch of one CTA and no fence:
//   1. PUSH   every rank stores (value, epoch) PAIRS — one 8-byte store each, single-copy atomic over NVLink — into ITS lane of every
//             rank's window: the epoch half vouches for the value half, so there is no flag, no release fence and no second round
//             (the low-latency protocol of NCCL's LL mode; round 2's first version, data + fence.sys + flag, measured 12.1 us per
//             exchange on two B200s, slower than NCCL's 10.6),
//   2. GATHER thread i polls slot i of every rank's lane in its own window (ld.relaxed.sys: the point of coherence, never a stale L1
//             line; bounded spin) and sums the values in RANK ORDER (every rank computes bit-identical sums — replicas must not drift).
// Window of one rank:  uint64 slot[2][world][capacity]  (value bits | epoch << 32).
// The epoch counter lives in device memory and is advanced by the kernel itself, so a captured CUDA graph replays correctly.
// Two parities: a rank can be at most one call ahead of a peer (it needs the peer's slots of call k to finish call k), so the lane
// it overwrites in call k+1 was last read in call k-1; a slot of call k-2 carries epoch k-2 and can never be mistaken for call k's.
// Plain SIMT: builds twice (nvcc -> libtfasr_b200.so; g++ -DTFASR_EMU_BUILD -> tests/emu, where the ranks are host threads sharing one
// address space and ThreadSanitizer checks the protocol).
#include <string.h>

#include "simt.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

constexpr int kPeerThreads = 256;
constexpr int kPeerMaxWorld = 64;

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const uint64_t* __restrict__ windows, int rank, int world, int capacity,
                                                                      uint32_t* __restrict__ state, float* __restrict__ data, int n,
                                                                      long long timeout_ticks) {
  const int tid = threadIdx.x;
  const uint32_t epoch = state[0] + 1u;  // written by the previous launch of this kernel (stream order)
  __syncthreads();                       // every thread has read the counter before thread 0 advances it
  const size_t par = epoch & 1u;
  const unsigned long long tag = (unsigned long long)epoch << 32;
  // 1. push (value, epoch) pairs into every rank's window (the own one included)
  for (int i = tid; i < n; i += blockDim.x) {
    const unsigned long long packed = tag | (unsigned long long)__float_as_uint(data[i]);
    for (int p = 0; p < world; ++p) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(windows[p]) + (par * world + rank) * capacity;
      simt_st_relaxed_sys_u64(dst + i, packed);
    }
  }
  // 2. gather: wait for every rank's slot i, sum in rank order
  const unsigned long long* lanes = reinterpret_cast<const unsigned long long*>(windows[rank]) + par * world * capacity;
  const long long t0 = simt_ticks();
  bool late = false;
  for (int i = tid; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* slot = lanes + (size_t)r * capacity + i;
      unsigned long long w = simt_ld_relaxed_sys_u64(slot);
      while ((uint32_t)(w >> 32) != epoch && !late) {
        if (simt_ticks() - t0 > timeout_ticks) late = true;  // a peer never arrived: flag it and fall through rather than hang the GPU
        w = simt_ld_relaxed_sys_u64(slot);
      }
      s += __uint_as_float((uint32_t)w);
    }
    data[i] = s;
  }
  if (late) atomicAdd(reinterpret_cast<int*>(state) + 1, 1);  // any non-zero value = a peer missed the wait (one atomic per late thread: no plain-store race)
  if (tid == 0) state[0] = epoch;
}

}  // namespace tfasr

using namespace tfasr;

extern "C" size_t tfasr_peer_window_bytes(int32_t world, int32_t capacity) {
  if (world < 1 || world > kPeerMaxWorld || capacity < 1) return 0;
  return (size_t)2 * world * capacity * sizeof(unsigned long long);
}

// Window management: plain CUDA IPC on the CURRENT device (one process per GPU; the 64-byte handles travel through the host-side process group).
extern "C" int tfasr_peer_window_alloc(size_t bytes, void** window, void* ipc_handle) {
  if (!bytes || !window || !ipc_handle) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  return TFASR_EXECUTION_FAILED;  // the emulator's ranks share host memory; nothing to map
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == TFASR_PEER_IPC_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  cudaIpcMemHandle_t h;
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    cudaFree(p);
    (void)cudaGetLastError();
    return TFASR_EXECUTION_FAILED;
  }
  memcpy(ipc_handle, &h, sizeof(h));
  *window = p;
  return TFASR_OK;
#endif
}

extern "C" int tfasr_peer_window_open(const void* ipc_handle, void** window) {
  if (!ipc_handle || !window) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  return TFASR_EXECUTION_FAILED;
#else
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void)cudaGetLastError();
    return TFASR_EXECUTION_FAILED;
  }
  *window = p;
  return TFASR_OK;
#endif
}

extern "C" int tfasr_peer_window_close(void* window, int32_t owned) {
  if (!window) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  (void)owned;
  return TFASR_EXECUTION_FAILED;
#else
  const cudaError_t e = owned ? cudaFree(window) : cudaIpcCloseMemHandle(window);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return TFASR_EXECUTION_FAILED;
  }
  return TFASR_OK;
#endif
}

extern "C" int tfasr_peer_allreduce_f32(const tfasr_peer_comm* comm, float* data, int32_t n, tfasr_stream_t stream_) {
  if (!comm || !comm->windows || !comm->state || comm->world < 1 || comm->world > kPeerMaxWorld || comm->rank < 0 || comm->rank >= comm->world ||
      comm->capacity < 1 || n < 0 || n > comm->capacity || (n > 0 && !data))
    return TFASR_INVALID_VALUE;
  if (n == 0) return TFASR_OK;  // every rank skips the same call: the epochs stay aligned
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  const long long timeout = comm->timeout_ms > 0 ? (long long)comm->timeout_ms * 1000000LL : 60LL * 1000000000LL;
  emu::launch(dim3(1), dim3(kPeerThreads), 0,
              [&] { peer_allreduce_kernel(comm->windows, comm->rank, comm->world, comm->capacity, comm->state, data, n, timeout); });
  return TFASR_OK;
#else
  // SM clock ticks at ~2 GHz; the default (30 s) only has to outlast a peer's slowest host-side hiccup (lazy library initialisation in the first eager step)
  const long long timeout = (comm->timeout_ms > 0 ? (long long)comm->timeout_ms : 30000LL) * 2000000LL;
  peer_allreduce_kernel<<<1, kPeerThreads, 0, static_cast<cudaStream_t>(stream_)>>>(comm->windows, comm->rank, comm->world, comm->capacity, comm->state,
                                                                                   data, n, timeout);
  count_launch();
  return check_launch();
#endif
}
