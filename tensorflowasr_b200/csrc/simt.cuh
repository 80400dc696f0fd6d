// Portability shim for the plain-SIMT kernel headers (*_kernels.cuh).
//
// Device build (nvcc, sm_100a): common.cuh + 16-byte access helpers.
// Emulation build (g++ -DTFASR_EMU_BUILD -I tests/emu): tests/emu/cuda_emu.h supplies threadIdx / __syncthreads / shuffles / atomics / bf16
// so the SAME kernel source runs on the CPU inside the `-m "not gpu"` suite (tests/test_emu_kernels.py).  Test infrastructure only: the
// emulation build never produces libtfasr_b200.so.
#pragma once
#ifdef TFASR_EMU_BUILD
#include "cuda_emu.h"
namespace tfasr {
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
constexpr int kNumSMs = 148;
#ifndef TFASR_KLOG2E_DEFINED
#define TFASR_KLOG2E_DEFINED
constexpr float kLog2e = 1.4426950408889634f;
#endif
}  // namespace tfasr
#else
#include "common.cuh"
#define TFASR_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace tfasr {

// 16-byte global accesses (8 bf16).  The emulation build aborts on a misaligned address, as the hardware would fault.
__device__ __forceinline__ uint4 ld16(const void* p) {
#ifdef TFASR_EMU_BUILD
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
#endif
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void st16(void* p, const uint4& v) {
#ifdef TFASR_EMU_BUILD
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
#endif
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void bf8_unpack(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 bf8_pack(const float* f) {
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    pk[j] = (uint32_t)(*reinterpret_cast<const uint16_t*>(&t.x)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&t.y)) << 16);
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// read-once 16-byte global load (device: ld.global.L1::no_allocate; emulator: plain load with the alignment check)
__device__ __forceinline__ float4 simt_ld_stream_f4(const float4* p) {
#ifdef TFASR_EMU_BUILD
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
  return *p;
#else
  return ld_stream_f4(p);
#endif
}

// read-once 16-byte global load of raw bits
__device__ __forceinline__ uint4 simt_ld_stream_u4(const uint4* p) {
#ifdef TFASR_EMU_BUILD
  if (reinterpret_cast<uintptr_t>(p) & 15) abort();
  return *p;
#else
  return ld_stream_u4(p);
#endif
}

// read-once 8-byte global load
__device__ __forceinline__ uint2 simt_ld_stream_u2(const uint2* p) {
#ifdef TFASR_EMU_BUILD
  if (reinterpret_cast<uintptr_t>(p) & 7) abort();
  return *p;
#else
  uint2 r;
  asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
#endif
}

// 2^x, flush-to-zero approximation (one MUFU, no range fix-up): for softmax terms whose underflow to zero is the right answer
__device__ __forceinline__ float simt_ex2(float x) {
#ifdef TFASR_EMU_BUILD
  return exp2f(x);
#else
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

__device__ __forceinline__ float simt_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// activation codes shared with the BatchNorm kernels: 0 none, 1 swish, 2 relu
__device__ __forceinline__ float simt_act(float z, int act) {
  if (act == 1) return z * simt_sigmoid(z);
  if (act == 2) return fmaxf(z, 0.f);
  return z;
}
__device__ __forceinline__ float simt_act_grad(float z, int act) {
  if (act == 1) {
    const float g = simt_sigmoid(z);
    return g * (1.0f + z * (1.0f - g));
  }
  if (act == 2) return z > 0.f ? 1.0f : 0.f;
  return 1.0f;
}

}  // namespace tfasr

// ---- system-scope flag accesses for exchanges over peer memory (csrc/peer_allreduce.cu) ----
// Device: PTX release / acquire at .sys scope (the flag lives in another GPU's memory, reached over NVLink).
// Emulator: std::atomic_ref with the same orders, so ThreadSanitizer checks exactly the ordering the protocol relies on.
namespace tfasr {
__device__ __forceinline__ void simt_st_release_sys(uint32_t* p, uint32_t v) {
#ifdef TFASR_EMU_BUILD
  std::atomic_ref<uint32_t>(*p).store(v, std::memory_order_release);
#else
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t simt_ld_acquire_sys(const uint32_t* p) {
#ifdef TFASR_EMU_BUILD
  return std::atomic_ref<uint32_t>(*const_cast<uint32_t*>(p)).load(std::memory_order_acquire);
#else
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
// data written by a peer GPU into this GPU's memory: read at the point of coherence (never from a stale L1 line)
__device__ __forceinline__ float simt_ld_relaxed_sys_f32(const float* p) {
#ifdef TFASR_EMU_BUILD
  return *p;
#else
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
#endif
}
// 8-byte (value, epoch) slots of the LL exchange: single-copy atomic over NVLink, so the epoch half vouches for the value half
__device__ __forceinline__ void simt_st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
#ifdef TFASR_EMU_BUILD
  std::atomic_ref<unsigned long long>(*p).store(v, std::memory_order_relaxed);
#else
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long simt_ld_relaxed_sys_u64(const unsigned long long* p) {
#ifdef TFASR_EMU_BUILD
  return std::atomic_ref<unsigned long long>(*const_cast<unsigned long long*>(p)).load(std::memory_order_relaxed);
#else
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void simt_fence_sys() {
#ifdef TFASR_EMU_BUILD
  std::atomic_thread_fence(std::memory_order_seq_cst);
#else
  __threadfence_system();
#endif
}
// monotonic tick counter for bounded spins (device: SM clock cycles; emulator: nanoseconds)
__device__ __forceinline__ long long simt_ticks() {
#ifdef TFASR_EMU_BUILD
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
#else
  return clock64();
#endif
}
}  // namespace tfasr
