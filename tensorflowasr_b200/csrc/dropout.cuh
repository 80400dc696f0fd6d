// Counter-based dropout mask shared by the GEMM epilogue (forward) and the row kernels (backward): the mask of element
// (row, col) of an [M, N] tensor is a pure function of (*seed_ptr + salt, row * N + col), so the backward regenerates it instead of
// storing it.  One 32-bit hash yields two 16-bit uniforms (an even/odd column pair); keep iff u16 >= round(p * 65536).
#pragma once
#include <stdint.h>

namespace tfasr {

__device__ __forceinline__ uint32_t dropout_hash(unsigned long long seed, unsigned long long even_idx) {
  // fold the 64-bit (seed, pair index) into 32 bits, then a murmur3-style finaliser
  const unsigned long long k = seed + (even_idx >> 1) * 0x9E3779B97F4A7C15ull;
  uint32_t h = (uint32_t)k ^ (uint32_t)(k >> 32);
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
