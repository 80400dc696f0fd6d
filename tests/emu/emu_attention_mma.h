// Emulator stand-in for csrc/attention_mma.cuh — TEST INFRASTRUCTURE ONLY.  The warp-collective tensor-core primitives of the attention
// kernels (ldmatrix, mma.sync m16n8k16 bf16 -> f32, cp.async) restated on the SIMT emulator with the PTX-documented fragment layouts:
// every lane publishes its operands in the warp's exchange area, the warp synchronises, every lane computes its own result registers.
//   ldmatrix x4 / x2 : lanes 8i..8i+7 supply the row addresses of 8x8 matrix i; lane l receives row l/4, columns 2(l%4), 2(l%4)+1 of each
//                      matrix (.trans: rows 2(l%4), 2(l%4)+1 of column l/4)
//   mma m16n8k16     : g = lane/4, t = lane%4.  A (16x16, row): a0 = (g, 2t..2t+1), a1 = (g+8, 2t..), a2 = (g, 2t+8..), a3 = (g+8, 2t+8..);
//                      B (16x8, col): b0 = (k = 2t..2t+1, n = g), b1 = (k = 2t+8.., n = g); C/D: c0,c1 = (g, 2t..2t+1), c2,c3 = (g+8, 2t..)
//   cp.async         : an immediate copy (a legal, stronger ordering); commit / wait are no-ops
// fp32 accumulation runs in ascending k (the hardware's order is unspecified: parity is to tolerance, as on the GPU).
#pragma once
#include <stdint.h>

namespace tfasr {
namespace mma {

namespace detail {
struct WarpX {  // per-warp exchange area, protected by the emulator's warp barrier
  const uint16_t* addr[32];
  uint32_t a[32][4];
  uint32_t b[32][2];
};
inline WarpX* warp_area() {
  // one area per emulated warp, keyed by the emulator's Warp object (lives for the whole launch)
  static thread_local emu::Warp* cached_w = nullptr;
  static thread_local WarpX* cached_x = nullptr;
  if (cached_w != emu::t_warp) {
    static std::mutex mu;
    static std::vector<std::pair<emu::Warp*, std::unique_ptr<WarpX>>> table;
    std::lock_guard<std::mutex> lk(mu);
    WarpX* found = nullptr;
    for (auto& e : table)
      if (e.first == emu::t_warp) found = e.second.get();
    if (!found) {
      table.emplace_back(emu::t_warp, std::make_unique<WarpX>());
      found = table.back().second.get();
    }
    cached_w = emu::t_warp;
    cached_x = found;
  }
  return cached_x;
}
inline float bf(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
inline void wsync() { emu::t_warp->bar.arrive_and_wait(); }

template <int NM, bool TRANS>
inline void ldsm(uint32_t* r, const void* row_ptr) {
  WarpX* x = warp_area();
  const int l = emu::t_lane;
  x->addr[l] = static_cast<const uint16_t*>(row_ptr);
  wsync();
  for (int i = 0; i < NM; ++i) {
    uint16_t lo, hi;
    if (!TRANS) {
      const uint16_t* row = x->addr[8 * i + l / 4];
      lo = row[2 * (l % 4)];
      hi = row[2 * (l % 4) + 1];
    } else {
      lo = x->addr[8 * i + 2 * (l % 4)][l / 4];
      hi = x->addr[8 * i + 2 * (l % 4) + 1][l / 4];
    }
    r[i] = (uint32_t)lo | ((uint32_t)hi << 16);
  }
  wsync();
}
}  // namespace detail

static inline void ldsm_x4(uint32_t (&r)[4], const void* p) { detail::ldsm<4, false>(r, p); }
static inline void ldsm_x4_trans(uint32_t (&r)[4], const void* p) { detail::ldsm<4, true>(r, p); }
static inline void ldsm_x2(uint32_t (&r)[2], const void* p) { detail::ldsm<2, false>(r, p); }
static inline void ldsm_x2_trans(uint32_t (&r)[2], const void* p) { detail::ldsm<2, true>(r, p); }

static inline void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  detail::WarpX* x = detail::warp_area();
  const int l = emu::t_lane, g = l / 4, t = l % 4;
  for (int i = 0; i < 4; ++i) x->a[l][i] = a[i];
  for (int i = 0; i < 2; ++i) x->b[l][i] = b[i];
  detail::wsync();
  auto A = [&](int row, int k) {  // element (row, k) of the 16 x 16 A tile
    const int reg = (row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0);
    const uint32_t w = x->a[(row % 8) * 4 + (k % 8) / 2][reg];
    return detail::bf((uint16_t)((k & 1) ? (w >> 16) : (w & 0xFFFF)));
  };
  auto B = [&](int k, int n) {  // element (k, n) of the 16 x 8 B tile
    const uint32_t w = x->b[n * 4 + (k % 8) / 2][k >= 8 ? 1 : 0];
    return detail::bf((uint16_t)((k & 1) ? (w >> 16) : (w & 0xFFFF)));
  };
  for (int i = 0; i < 4; ++i) {
    const int row = g + (i >= 2 ? 8 : 0), n = 2 * t + (i & 1);
    float acc = c[i];
    for (int k = 0; k < 16; ++k) acc = fmaf(A(row, k), B(k, n), acc);
    c[i] = acc;
  }
  detail::wsync();
}

static inline void cp_async4(void* smem, const void* gmem) { memcpy(smem, gmem, 4); }
static inline void cp_async8(void* smem, const void* gmem) { memcpy(smem, gmem, 8); }
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}

}  // namespace mma
}  // namespace tfasr
