// Emulator stand-in for csrc/sm100.cuh — TEST INFRASTRUCTURE ONLY.  The sm_100a features the GEMM uses, restated on the SIMT emulator:
//   mbarrier        arrival count + transaction bytes + phase parity, in the 8 bytes the kernel reserves (std::atomic on them)
//   TMA             cp.async.bulk.tensor 2-D tile load / store / f32 reduce-add as immediate copies between global memory and the box layout
//                   in shared memory, with the tensor map's swizzle (128B / 64B / none: XOR of the 16-byte chunk index with address bits
//                   [7,10) resp. [7,9)), zero fill of out-of-bounds elements on loads and clipping on stores
//   TMEM            128 lanes x 512 32-bit columns per CTA (keyed by the CTA's shared-memory base)
//   tcgen05.mma     kind::f16, cta_group::1: D[128 x N] (+)= A[128 x 16] . B[N x 16]^T, operands read from shared memory through the
//                   UMMA descriptors the kernel builds (start address, LBO, SBO, SWIZZLE_128B; K-major or MN-major per the instruction
//                   descriptor), fp32 accumulation in ascending k; commit = arrive on the mbarrier (the emulated MMA is synchronous)
//   tcgen05.ld      32x32b.x16: lane l of the warp reads TMEM lane (addr >> 16) + l, 16 consecutive columns
// Shared-memory "addresses" (smem_u32) are byte offsets from the CTA's dynamic shared memory base, which the emulator allocates 1024-byte
// aligned for launches of this kernel — the swizzle patterns are functions of absolute address bits.
#pragma once
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <thread>

struct alignas(64) CUtensorMap {
  const uint8_t* base;
  uint64_t rows, cols, ld_bytes;
  uint32_t box_rows, box_cols, elem_bytes, swizzle;  // swizzle: 0 none, 2 = 64B, 3 = 128B (log2 of the 16-byte chunks spanned)
  uint64_t pad[3];
};

namespace tfasr {

inline bool make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                         uint32_t box_cols) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15)) return false;
  const uint32_t inner = box_cols * elem_bytes;
  if (inner != 32 && inner != 64 && inner != 128) return false;  // cuTensorMapEncodeTiled: the swizzle span must cover the box's inner extent
  if (box_rows > 256 || box_cols > 256) return false;
  map->base = static_cast<const uint8_t*>(base);
  map->rows = rows;
  map->cols = cols;
  map->ld_bytes = ld * elem_bytes;
  map->box_rows = box_rows;
  map->box_cols = box_cols;
  map->elem_bytes = (uint32_t)elem_bytes;
  map->swizzle = inner == 32 ? 0 : (inner == 64 ? 2 : 3);
  return true;
}
inline bool make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(map, base, 2, rows, cols, ld, box_rows, box_cols);
}

namespace sm100 {

static inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const uint8_t*>(p) - emu::t_dyn_smem); }
static inline uint8_t* smem_ptr(uint32_t a) { return emu::t_dyn_smem + a; }

// physical byte offset of logical offset `off` inside a swizzled region (off and the region base relative to the 1024-aligned smem base)
static inline uint32_t swz(uint32_t addr, uint32_t mode) {
  if (mode == 3) return addr ^ (((addr >> 7) & 7u) << 4);
  if (mode == 2) return addr ^ (((addr >> 7) & 3u) << 4);
  return addr;
}

// ---------------------------------------------------------------------------------------------- mbarrier
struct MBar {  // lives in the kernel's 8-byte slot
  std::atomic<int32_t> state;  // bits [0,1): phase parity | bits [1,12): pending arrivals | bits [12,32): pending tx bytes / 16... kept simple below
  std::atomic<int32_t> aux;
};
// Representation: `pending` arrivals and `tx` bytes packed under one mutex-free protocol would be intricate; a per-barrier spin lock is enough.
struct MBarState {
  std::atomic_flag lock = ATOMIC_FLAG_INIT;
  int32_t count = 0, pending = 0;
  int64_t tx = 0;
  uint32_t phase = 0;  // completed phases
};
static_assert(sizeof(uint64_t) == 8, "");
inline std::mutex& mbar_table_mu() {
  static std::mutex mu;
  return mu;
}
inline std::map<const void*, std::unique_ptr<MBarState>>& mbar_table() {
  static std::map<const void*, std::unique_ptr<MBarState>> t;
  return t;
}
static inline MBarState* mbar_get(const uint64_t* bar) {
  // the kernel's 8-byte slot stores the pointer of the state object (written by mbar_init)
  MBarState* s;
  memcpy(&s, bar, sizeof(s));
  return s;
}
static inline void mbar_lock(MBarState* s) {
  while (s->lock.test_and_set(std::memory_order_acquire)) std::this_thread::yield();
}
static inline void mbar_unlock(MBarState* s) { s->lock.clear(std::memory_order_release); }
static inline void mbar_check_complete(MBarState* s) {  // called with the lock held
  if (s->pending == 0 && s->tx == 0) {
    s->pending = s->count;
    ++s->phase;
  }
}
static inline void mbar_init(uint64_t* bar, uint32_t count) {
  MBarState* s;
  {
    std::lock_guard<std::mutex> lk(mbar_table_mu());
    auto& slot = mbar_table()[bar];
    slot = std::make_unique<MBarState>();
    s = slot.get();
  }
  s->count = s->pending = (int32_t)count;
  memcpy(bar, &s, sizeof(s));
}
static inline void fence_barrier_init() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void fence_proxy_async() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void mbar_arrive(uint64_t* bar) {
  MBarState* s = mbar_get(bar);
  mbar_lock(s);
  --s->pending;
  mbar_check_complete(s);
  mbar_unlock(s);
}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  MBarState* s = mbar_get(bar);
  mbar_lock(s);
  s->tx += bytes;
  --s->pending;
  mbar_check_complete(s);
  mbar_unlock(s);
}
static inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  MBarState* s = mbar_get(bar);
  mbar_lock(s);
  s->tx -= bytes;
  mbar_check_complete(s);
  mbar_unlock(s);
}
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {  // true once the phase with this parity has completed
  MBarState* s = mbar_get(bar);
  mbar_lock(s);
  const bool ok = ((s->phase & 1u) != parity);
  mbar_unlock(s);
  if (!ok) std::this_thread::yield();
  return ok;
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint64_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1ull << 26)) abort();  // a mis-programmed pipeline: fail instead of hanging the test
}

// ---------------------------------------------------------------------------------------------- TMA
static inline void prefetch_tensormap(const CUtensorMap*) {}
static inline void grid_dependency_wait() {}  // launches are serialised on the emulator
static inline void grid_launch_dependents() {}
// box element (r, c) <-> global element (c1 + r, c0 + c); shared memory: row r at r * inner bytes, swizzled
static inline void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  const uint32_t inner = m->box_cols * m->elem_bytes, base = smem_u32(smem_dst);
  if (((int64_t)c0 * m->elem_bytes) % 16 != 0) abort();  // seen on the B200: a box that does not start on a 16-byte boundary of the row is an illegal instruction
  for (uint32_t r = 0; r < m->box_rows; ++r)
    for (uint32_t c = 0; c < m->box_cols; ++c) {
      const int64_t gr = (int64_t)c1 + r, gc = (int64_t)c0 + c;
      uint8_t* d = smem_ptr(swz(base + r * inner + c * m->elem_bytes, m->swizzle));
      if (gr >= 0 && gc >= 0 && (uint64_t)gr < m->rows && (uint64_t)gc < m->cols) memcpy(d, m->base + (uint64_t)gr * m->ld_bytes + (uint64_t)gc * m->elem_bytes, m->elem_bytes);
      else memset(d, 0, m->elem_bytes);
    }
  mbar_complete_tx(bar, m->box_rows * inner);
}
static inline void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  const uint32_t inner = m->box_cols * m->elem_bytes, base = smem_u32(smem_src);
  if (((int64_t)c0 * m->elem_bytes) % 16 != 0) abort();
  for (uint32_t r = 0; r < m->box_rows; ++r)
    for (uint32_t c = 0; c < m->box_cols; ++c) {
      const int64_t gr = (int64_t)c1 + r, gc = (int64_t)c0 + c;
      if (gr < 0 || gc < 0 || (uint64_t)gr >= m->rows || (uint64_t)gc >= m->cols) continue;  // clipped by the hardware
      memcpy(const_cast<uint8_t*>(m->base) + (uint64_t)gr * m->ld_bytes + (uint64_t)gc * m->elem_bytes,
             smem_ptr(swz(base + r * inner + c * m->elem_bytes, m->swizzle)), m->elem_bytes);
    }
}
static inline void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  const uint32_t inner = m->box_cols * m->elem_bytes, base = smem_u32(smem_src);
  if (m->elem_bytes != 4) abort();
  for (uint32_t r = 0; r < m->box_rows; ++r)
    for (uint32_t c = 0; c < m->box_cols; ++c) {
      const int64_t gr = (int64_t)c1 + r, gc = (int64_t)c0 + c;
      if (gr < 0 || gc < 0 || (uint64_t)gr >= m->rows || (uint64_t)gc >= m->cols) continue;
      float v;
      memcpy(&v, smem_ptr(swz(base + r * inner + c * 4, m->swizzle)), 4);
      float* g = reinterpret_cast<float*>(const_cast<uint8_t*>(m->base) + (uint64_t)gr * m->ld_bytes + (uint64_t)gc * 4);
      std::atomic_ref<float>(*g).fetch_add(v, std::memory_order_relaxed);
    }
}
static inline void tma_store_commit() {}
template <int N>
static inline void tma_store_wait_read() {}
static inline void tma_store_wait_all() {}

// ---------------------------------------------------------------------------------------------- TMEM
struct Tmem {
  uint32_t cell[128][512];
};
static inline Tmem* tmem_of_cta() {
  static std::mutex mu;
  static std::map<const void*, std::unique_ptr<Tmem>> table;
  static thread_local const void* cached_key = nullptr;
  static thread_local Tmem* cached = nullptr;
  if (cached_key != emu::t_dyn_smem) {
    std::lock_guard<std::mutex> lk(mu);
    auto& slot = table[emu::t_dyn_smem];
    if (!slot) slot = std::make_unique<Tmem>();
    cached = slot.get();
    cached_key = emu::t_dyn_smem;
  }
  return cached;
}
static inline void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  if (ncols > 512) abort();
  if (emu::t_lane == 0) *smem_result = 0;  // column 0, lane 0
  emu::t_warp->bar.arrive_and_wait();      // .sync.aligned: the whole warp
}
static inline void tmem_dealloc(uint32_t, uint32_t) {}
static inline void tc_fence_before() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void tc_fence_after() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t make_idesc_bf16_f32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// element (i = row of the M / N dimension, k in 0..15) of an operand tile described by `desc`
static inline float umma_operand(uint64_t desc, bool mn_major, int i, int k) {
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  if (((desc >> 61) & 7) != 2) abort();  // SWIZZLE_128B is the only layout the kernel uses
  uint32_t off;
  if (!mn_major) off = (uint32_t)(i / 8) * sbo + (uint32_t)(i % 8) * 128 + (uint32_t)k * 2;            // K-major: rows of 64 k (128 B), 8-row atoms
  else off = (uint32_t)(i / 64) * lbo + (uint32_t)(k / 8) * sbo + (uint32_t)(k % 8) * 128 + (uint32_t)(i % 64) * 2;  // MN-major: k-rows of 64 mn
  uint16_t h;
  memcpy(&h, smem_ptr(swz(start + off, 3)), 2);
  return __uint_as_float((uint32_t)h << 16);
}
static inline void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (M != 128 || (tmem_d >> 16) != 0) abort();
  Tmem* t = tmem_of_cta();
  const uint32_t col0 = tmem_d & 0xFFFF;
  float a[128][16];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 16; ++k) a[m][k] = umma_operand(adesc, a_mn, m, k);
  for (int n = 0; n < N; ++n) {
    float b[16];
    for (int k = 0; k < 16; ++k) b[k] = umma_operand(bdesc, b_mn, n, k);
    for (int m = 0; m < 128; ++m) {
      float acc = accumulate ? __uint_as_float(t->cell[m][col0 + n]) : 0.f;
      for (int k = 0; k < 16; ++k) acc = fmaf(a[m][k], b[k], acc);
      t->cell[m][col0 + n] = __float_as_uint(acc);
    }
  }
}
static inline void umma_commit(uint64_t* bar) { mbar_arrive(bar); }
static inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  Tmem* t = tmem_of_cta();
  const uint32_t lane = (taddr >> 16) + (uint32_t)emu::t_lane, col = taddr & 0xFFFF;
  if (lane >= 128 || col + 16 > 512) abort();
  for (int j = 0; j < 16; ++j) r[j] = t->cell[lane][col + j];
}
static inline void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  Tmem* t = tmem_of_cta();
  const uint32_t lane = (taddr >> 16) + (uint32_t)emu::t_lane, col = taddr & 0xFFFF;
  if (lane >= 128 || col + 8 > 512) abort();
  for (int j = 0; j < 8; ++j) r[j] = t->cell[lane][col + j];
}
static inline void tmem_ld_wait() {}

}  // namespace sm100
}  // namespace tfasr

namespace emu {
// bar.sync id, nthreads: a named barrier over a subset of the CTA's threads
inline void named_barrier_sync(int id, int nthreads) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, std::unique_ptr<std::barrier<>>> table;
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto& slot = table[{t_dyn_smem, id}];
    if (!slot) slot = std::make_unique<std::barrier<>>(nthreads);
    b = slot.get();
  }
  b->arrive_and_wait();
}
}  // namespace emu
