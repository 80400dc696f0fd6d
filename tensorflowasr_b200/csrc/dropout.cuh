// Counter-based dropout mask shared by the GEMM epilogue (forward) and the row kernels (backward): the mask of element
// (row, col) of an [M, N] tensor is a pure function of (*seed_ptr + salt, row, col), so the backward regenerates it instead of
// storing it.  One 32-bit hash yields two 16-bit uniforms (an even/odd column pair); keep iff u16 >= round(p * 65536).
#pragma once
#include <stdint.h>

namespace tfasr {

// 64-bit work once per row: (seed, row) -> 32-bit row key
__device__ __forceinline__ uint32_t dropout_rowkey(unsigned long long seed, unsigned long long row) {
  const unsigned long long k = seed * 0xD6E8FEB86659FD93ull + row * 0x9E3779B97F4A7C15ull;
  uint32_t h = (uint32_t)k ^ (uint32_t)(k >> 32);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// per column pair: Weyl step on the row key + murmur3 finaliser (32-bit ops only: the GEMM epilogue runs this 8 times per 16 columns)
__device__ __forceinline__ uint32_t dropout_hash(uint32_t rowkey, uint32_t even_col) {
  uint32_t h = rowkey + (even_col >> 1) * 0x9E3779B9u;
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

__device__ __forceinline__ uint32_t dropout_threshold(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
}

}  // namespace tfasr
