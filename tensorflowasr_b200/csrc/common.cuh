// Shared helpers for libtfasr_b200 kernels (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>

#include "../../include/tfasr_b200.h"

namespace tfasr {

extern std::atomic<int> g_launch_count;  // kernels launched by this library (bench.py's gpu_launches)

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TFASR_OK : TFASR_EXECUTION_FAILED;
}

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log(exp(a) + exp(b)) with -inf handled (no NaN from inf - inf).
__device__ __forceinline__ float log_add_exp(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + log1pf(expf(-fabsf(a - b)));
}
// Same, on the MUFU fast path (ex2 / lg2): this sits on the critical path of the lattice wavefronts, where one step's
// result feeds the next.  The argument of the log is in [1, 2] and of the exp in (-inf, 0]: absolute error ~1e-7, far
// below the fp32 ulp of the lattice values themselves (|alpha| ~ 1e3 -> ulp 6e-5).
__device__ __forceinline__ float log_add_exp_fast(float a, float b) {
  const float m = fmaxf(a, b);
  const float d = -fabsf(a - b);  // -inf - -inf = NaN only when both are -inf
  const float r = m + __logf(1.0f + __expf(d));
  return (m == -CUDART_INF_F) ? -CUDART_INF_F : r;
}

// ---- programmatic dependent launch (TFASR_PDL, default on): a kernel launched through launch_pdl may be scheduled while its predecessor in
// the stream is still draining, so its launch latency (and whatever it does before grid_dependency_wait()) overlaps the predecessor's tail;
// grid_dependency_wait() returns once the predecessor has completed and its memory operations are visible (a no-op for an ordinary launch).
// RULE: every kernel launched through launch_pdl calls grid_dependency_wait() before its first global-memory access.  The step graph keeps
// these launches as programmatic edges; the training step is a chain of ~700 dependent, few-microsecond kernels.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// The other half: a kernel that executes this in every CTA lets its programmatic dependents be SCHEDULED from then on (they still block in
// their grid_dependency_wait() until this grid has completed).  Used by the persistent tcgen05 kernels only — their grid is at most one CTA
// per SM, all resident from the start, so an early-resident dependent can never hold back a CTA of the primary.  TFASR_PDL_TRIGGER=0: off.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled();
inline bool pdl_trigger_enabled() {
  static const bool on = [] {
    const char* v = getenv("TFASR_PDL_TRIGGER");
    return !(v && v[0] == '0');
  }();
  return on && pdl_enabled();
}
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("TFASR_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 2^x as ONE MUFU (flush-to-zero: no range fix-up around it) — for softmax terms, where underflow to zero is the right answer
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// streaming 128-bit global accesses (read-once / write-once data should not pollute L1)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

}  // namespace tfasr
