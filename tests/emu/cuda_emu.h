// CPU emulation of the CUDA SIMT execution model — TEST INFRASTRUCTURE ONLY (never linked into libtfasr_b200.so).
//
// Purpose: the plain-SIMT kernels of tensorflowasr_b200/csrc/*_kernels.cuh (no TMA / tcgen05 / cp.async: thread and block indices,
// shared memory, __syncthreads, warp shuffles, atomics, bf16 conversions) are compiled a second time with g++ against this header
// and executed block by block with one OS thread per CUDA thread.  That gives the `-m "not gpu"` suite a bit-level check of the
// kernels' indexing, masking, reductions and launch geometry against torch CPU references — and, under AddressSanitizer with
// exact-size heap buffers, what compute-sanitizer memcheck would report (out-of-bounds global / shared accesses, misaligned vector
// accesses through the ld/st helpers below).  It says nothing about performance or about device-only intrinsics' rounding
// (__expf etc. map to the libm functions here).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <chrono>
#include <memory>
#include <thread>
#include <vector>

#define TFASR_EMU 1
#define __global__ static  // internal linkage: kernels / device helpers of different files may share names
#define __device__ static
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static  // blocks run one after another, so a function-local static is exactly one block's shared variable

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 {
  float x, y;
};
struct alignas(8) uint2 {
  uint32_t x, y;
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(16) uint4 {
  uint32_t x, y, z, w;
};
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }

// ------------------------------------------------------------------------------------------------ bf16 (round-to-nearest-even)
struct __nv_bfloat16 {
  uint16_t v;
  operator float() const {  // as cuda_bf16.h: bf16 -> f32 is exact
    uint32_t u = (uint32_t)v << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  }
};
struct alignas(4) __nv_bfloat162 {
  __nv_bfloat16 x, y;
};
inline float __uint_as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
inline float __bfloat162float(__nv_bfloat16 h) { return __uint_as_float((uint32_t)h.v << 16); }
inline __nv_bfloat16 __float2bfloat16(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return {(uint16_t)0x7fff};  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return {(uint16_t)(u >> 16)};
}
inline __nv_bfloat16 __float2bfloat16_rn(float f) { return __float2bfloat16(f); }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return {__float2bfloat16(a), __float2bfloat16(b)}; }
inline float __low2float(__nv_bfloat162 v) { return __bfloat162float(v.x); }
inline float __high2float(__nv_bfloat162 v) { return __bfloat162float(v.y); }
inline __nv_bfloat16 __low2bfloat16(__nv_bfloat162 v) { return v.x; }
inline __nv_bfloat16 __high2bfloat16(__nv_bfloat162 v) { return v.y; }
inline float2 __bfloat1622float2(__nv_bfloat162 v) { return {__bfloat162float(v.x), __bfloat162float(v.y)}; }

// ------------------------------------------------------------------------------------------------ device math on the host
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T>
inline T min(T a, T b) { return a < b ? a : b; }
template <class T>
inline T max(T a, T b) { return a > b ? a : b; }
#define CUDART_INF_F (__builtin_inff())

// ------------------------------------------------------------------------------------------------ execution model
namespace emu {

struct Warp {
  int n;  // live threads in this warp
  std::barrier<> bar;
  uint64_t buf[32];
  explicit Warp(int n_) : n(n_), bar(n_) {}
};

struct BlockCtx {
  std::barrier<>* bar;
  std::vector<std::unique_ptr<Warp>>* warps;
};

inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local Warp* t_warp;
inline thread_local int t_lane;
inline thread_local std::barrier<>* t_bar;
inline thread_local uint8_t* t_dyn_smem = nullptr;  // this CTA's dynamic shared memory: an exact-size heap block, so ASan sees overruns
struct ClusterCtx {  // thread-block cluster (launch_cluster): the CTAs of one cluster run concurrently
  int size;
  std::vector<uint8_t*> smem;  // dynamic shared memory of every CTA of the cluster (distributed shared memory)
  std::barrier<> bar;
  ClusterCtx(int size_, int threads) : size(size_), smem(size_, nullptr), bar(threads) {}
};
inline thread_local ClusterCtx* t_cluster = nullptr;
inline thread_local int t_cluster_rank = 0;
inline std::atomic<long> g_launches{0};

// One OS thread per CUDA thread of a block, created once per launch; the blocks of the grid run one after another on them (a second
// barrier separates consecutive blocks, so shared memory — dynamic and function-local statics — always belongs to exactly one block).
// A kernel in which some threads return before a __syncthreads() that the others still execute would deadlock here (on the GPU exited
// threads stop counting); none of the kernels does that.
template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
  const int nth = (int)(block.x * block.y * block.z);
  if (nth <= 0 || nth > 1024) abort();
  g_launches++;
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  std::barrier<> bar(nth), done_bar(nth);
  std::vector<std::unique_ptr<Warp>> warps;
  for (int w = 0; w * 32 < nth; ++w) warps.emplace_back(new Warp(nth - w * 32 < 32 ? nth - w * 32 : 32));
  // exact size (rounded up to 16 by aligned_alloc: the tail of < 16 bytes is the only slack ASan cannot see)
  uint8_t* smem = smem_bytes ? (uint8_t*)aligned_alloc(16, (smem_bytes + 15) / 16 * 16) : nullptr;
  std::vector<std::thread> ths;
  ths.reserve(nth);
  for (int t = 0; t < nth; ++t)
    ths.emplace_back([&, t] {
      t_threadIdx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
      t_blockDim = block;
      t_gridDim = grid;
      t_warp = warps[t / 32].get();
      t_lane = t % 32;
      t_bar = &bar;
      t_dyn_smem = smem;
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            t_blockIdx = {bx, by, bz};
            body();
            done_bar.arrive_and_wait();
          }
    });
  for (auto& th : ths) th.join();
  free(smem);
}

// Thread-block clusters of `csize` CTAs along x (__cluster_dims__(csize, 1, 1)): the CTAs of one cluster run CONCURRENTLY — csize x nth OS
// threads — each with its own dynamic shared memory, CTA barrier and warps, plus one cluster-wide barrier (cluster.sync()) and the
// table of peers' shared memory (cluster.map_shared_rank).  Clusters run one after another.  Kernels launched this way must keep their
// shared data in DYNAMIC shared memory: a function-local __shared__ static would be common to the concurrent CTAs.
template <class F>
void launch_cluster(dim3 grid, dim3 block, size_t smem_bytes, int csize, F&& body) {
  const int nth = (int)(block.x * block.y * block.z);
  if (nth <= 0 || nth > 1024 || csize < 1 || grid.x % csize) abort();
  g_launches++;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned cx = 0; cx < grid.x / csize; ++cx) {
        ClusterCtx ctx(csize, csize * nth);
        std::vector<std::unique_ptr<std::barrier<>>> bars;
        std::vector<std::vector<std::unique_ptr<Warp>>> warps(csize);
        for (int r = 0; r < csize; ++r) {
          bars.emplace_back(new std::barrier<>(nth));
          for (int w = 0; w * 32 < nth; ++w) warps[r].emplace_back(new Warp(nth - w * 32 < 32 ? nth - w * 32 : 32));
          ctx.smem[r] = smem_bytes ? (uint8_t*)aligned_alloc(16, (smem_bytes + 15) / 16 * 16) : nullptr;
        }
        std::vector<std::thread> ths;
        ths.reserve((size_t)csize * nth);
        for (int r = 0; r < csize; ++r)
          for (int t = 0; t < nth; ++t)
            ths.emplace_back([&, r, t] {
              t_threadIdx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
              t_blockIdx = {cx * csize + r, by, bz};
              t_blockDim = block;
              t_gridDim = grid;
              t_warp = warps[r][t / 32].get();
              t_lane = t % 32;
              t_bar = bars[r].get();
              t_dyn_smem = ctx.smem[r];
              t_cluster = &ctx;
              t_cluster_rank = r;
              body();
            });
        for (auto& th : ths) th.join();
        for (int r = 0; r < csize; ++r) free(ctx.smem[r]);
      }
}

template <class T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Warp& w = *t_warp;
  memcpy(&w.buf[t_lane], &v, sizeof(T));
  w.bar.arrive_and_wait();
  T r = v;
  if (src_lane >= 0 && src_lane < w.n) memcpy(&r, &w.buf[src_lane], sizeof(T));
  w.bar.arrive_and_wait();
  return r;
}

}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

inline void __syncthreads() { emu::t_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::t_warp->bar.arrive_and_wait();
}
// full-mask, convergent shuffles only (what the kernels use)
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::shfl(v, emu::t_lane ^ lane_mask); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta) { return emu::shfl(v, emu::t_lane + delta); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta) { return emu::shfl(v, emu::t_lane - delta); }

inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline void atomicAdd(float4* p, float4 v) {  // red.global.add.v4.f32: four independent float adds
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
  atomicAdd(&p->z, v.z);
  atomicAdd(&p->w, v.w);
}
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }

// dynamic shared memory: kernels declare it with TFASR_DYN_SMEM(name)
#define TFASR_DYN_SMEM(name) uint8_t* name = emu::t_dyn_smem
#define __cluster_dims__(...)
#define __grid_constant__
