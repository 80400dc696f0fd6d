// Emulator stand-in for csrc/common.cuh — TEST INFRASTRUCTURE ONLY.  The first-generation kernel files (rnnt_loss.cu, ctc_loss.cu,
// fused_rows.cu, conv_bn.cu, dwconv.cu, joint.cu, param_ops.cu, logmel.cu, specaugment.cu) were written against common.cuh and launch
// with <<<>>>; tests/emu/cu2emu.py rewrites the launch statements, the dynamic shared memory declarations and the one inline-PTX helper,
// swaps `#include "common.cuh"` for this header, and the result compiles with g++ against the SIMT emulator (cuda_emu.h).
#pragma once
#include "cuda_emu.h"

#include <atomic>
#include <mutex>
#include <utility>

#include "../../include/tfasr_b200.h"

// ---- the slice of the CUDA runtime API the launchers use
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}

template <class T>
inline T __ldg(const T* p) { return *p; }
inline float log1pf_(float x) { return log1pf(x); }
namespace tfasr {
inline void grid_dependency_wait() {}  // launches are serialised on the emulator
inline void grid_launch_dependents() {}
inline bool pdl_trigger_enabled() { return false; }
template <class... KArgs, class... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args&&... args) {
  emu::launch(grid, block, smem, [&] { kernel(static_cast<KArgs>(args)...); });
  return 0;
}
#ifndef TFASR_KLOG2E_DEFINED
#define TFASR_KLOG2E_DEFINED
constexpr float kLog2e = 1.4426950408889634f;
#endif
inline float fast_ex2(float x) { return exp2f(x); }
}  // namespace tfasr
// mixed-width min / max as CUDA's overload set resolves them
inline int64_t min(int64_t a, int b) { return a < b ? a : b; }
inline int64_t min(int a, int64_t b) { return a < b ? a : b; }
inline int64_t max(int64_t a, int b) { return a > b ? a : b; }
inline int64_t max(int a, int64_t b) { return a > b ? a : b; }
inline size_t min(size_t a, int b) { return a < (size_t)b ? a : (size_t)b; }

// ---- cooperative_groups: the cluster slice the LSTM kernels use
namespace cooperative_groups {
struct cluster_group {
  unsigned block_rank() const { return (unsigned)emu::t_cluster_rank; }
  template <class T>
  T* map_shared_rank(T* p, unsigned rank) const {  // the same offset inside the peer CTA's shared memory
    return reinterpret_cast<T*>(emu::t_cluster->smem[rank] + (reinterpret_cast<uint8_t*>(p) - emu::t_dyn_smem));
  }
  void sync() const { emu::t_cluster->bar.arrive_and_wait(); }
};
inline cluster_group this_cluster() { return {}; }
}  // namespace cooperative_groups

namespace tfasr {

inline std::atomic<int> g_launch_count{0};
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
inline int check_launch() { return TFASR_OK; }

constexpr int kNumSMs = 148;
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

static inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
static inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline float log_add_exp(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + log1pf(expf(-fabsf(a - b)));
}
static inline float log_add_exp_fast(float a, float b) {
  const float m = fmaxf(a, b);
  const float d = -fabsf(a - b);
  const float r = m + __logf(1.0f + __expf(d));
  return (m == -CUDART_INF_F) ? -CUDART_INF_F : r;
}
static inline float4 ld_stream_f4(const float4* p) {
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
  return *p;
}
static inline uint4 ld_stream_u4(const void* p) {
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
  return *reinterpret_cast<const uint4*>(p);
}
static inline void st_stream_f4(float4* p, const float4& v) {
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
  *p = v;
}

}  // namespace tfasr
