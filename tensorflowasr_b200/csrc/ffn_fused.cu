// Macaron feed-forward module of the Conformer block as ONE kernel per direction (tensorflow_asr/models/encoders/conformer.py:101-109
// FFModule.call:  out = x + factor * Drop(W2 . Drop(swish(W1 . LN(x) + b1)) + b2)).
//
// Why: at d_model = 144 a Dense layer per launch is latency, not work (round 1: LayerNorm + 2 GEMM launches = 33 us forward for 2.65 GFLOP,
// i.e. ~2 us of tensor time).  Here one CTA owns a 128-row tile for the whole module:
//   TMA   x tile [128 x D] -> SMEM in the UMMA K-major 128B-swizzle layout
//   LN    128 threads (one row each) normalise the tile IN PLACE in SMEM (two-pass, fp32) -> the A operand; xn also leaves through a TMA
//         store (the weight-gradient GEMM of the backward wants it)
//   loop over the hidden dimension in chunks of 64:
//     G1  tcgen05.mma 128 x 64 x 16 (x9):   acc1[c & 1] (TMEM, 64 columns) = xn . W1[:, chunk]
//     E1  8 epilogue warps: tcgen05.ld -> + b1 -> z (saved pre-activation) -> swish -> dropout -> bf16 h chunk, written into SMEM as the next
//         A operand — the [128 x 576] hidden activation never makes the HBM round trip before GEMM 2 — and, from the same SMEM tiles, to
//         global by TMA stores (z, h: saved for the backward)
//     G2  tcgen05.mma 128 x D x 16 (x4):    acc2 (TMEM, D columns) += h_chunk . W2[chunk, :]
//   E2    acc2 -> + b2 -> dropout -> x + factor * (.) -> bf16 out -> SMEM -> TMA store
// Weight chunks (W1^T rows [64 x D], W2^T columns [D x 64]: 42 KB per chunk) stream through a 2-stage TMA ring from L2, where the 332 KB of
// module weights stay resident.  G1 of chunk c+1 overlaps E1 of chunk c (double-buffered acc1, h and z tiles).
// Every global access of the hot loop is a TMA transfer: thread-per-row 16-byte accesses to [M, 576] rows (32 cache lines per warp
// instruction) made the first version of this kernel slower than the three launches it replaces.
//
// Backward (MODE 1) is the same two-GEMM skeleton run on the gradient chain of FFModule's backward:
//   TMA   dout tile -> SMEM;  prologue: dy = factor * Drop2'(dout) in place (the A operand; dy also leaves through the TMA store for dW2 / db2)
//   G1    acc1 = dy . W2[chunk, :]^T            E1: dz = acc1 * Drop1' * swish'(z) (z chunk by TMA, with the weights) -> SMEM -> TMA store
//   G2    acc2 += dz_chunk . W1[:, chunk]^T     E2: da = acc2 -> bf16 (LayerNorm's input gradient; the LayerNorm backward kernel follows)
// i.e. scale_mask_colsum + the swish' GEMM + the dgrad GEMM of round 1 in one launch; weight / bias gradients run beside it on the side stream.
//
// Roofline: tensor pipe; per tile 2 * 128 * D * H * 2 flop = 42.5 MFLOP at D = 144, H = 576 -> 2.7 us at the per-SM bf16 rate; M = 8000 gives
// 63 tiles on 148 SMs, so the module's floor is one tile time.  See DESIGN.md section 4.
#include "gemm_kernel.cuh"

namespace tfasr {

// Tile = 64 rows (of a 128-row MMA: TMEM lanes 64..127 compute on stale SMEM and are never read).  The epilogue — sigmoid, dropout hash,
// packing: ~19 instructions per hidden activation — is what a tile costs, not the tensor work (9 % of the tensor pipe at 128 rows), so the
// module spreads over 125 SMs instead of 63.  (Measured: a tile's 64 rows sit in TMEM lane quarters 0 and 1, so its epilogue warps are the
// warps with id % 4 in {0, 1} — two of the SM's four schedulers; 125 SMs x 2 schedulers is the issue capacity of 63 SMs x 4, which is why this
// layout gained 3 us per module, not 2x, and why 16 epilogue warps of 8 columns each (1024 threads) measured the same 0.85 us per chunk.  The
// next step is a 2-CTA cluster per 128-row tile that splits the HIDDEN dimension and reduces the output accumulator over DSMEM.)  Warps: 2 = TMA producer, 3 = MMA issuer (+ the chunk stores); 0,1 / 4,5 / 8,9 / 12,13 =
// epilogue + LayerNorm (TMEM lane quarters 0 and 1 = the 64 rows; `sub` = which 16 of a chunk's 64 columns); 6,7,10,11 idle.
constexpr int kFfnEpiWarps = 8;
constexpr int kFfnThreads = 448;
constexpr int kFfnRows = 64;
constexpr int kHC = 64;           // hidden-dimension chunk = one 128B-swizzle atom of K for GEMM 2

template <int D>
struct FfnSmem {
  static_assert(D % 16 == 0 && D <= 192, "model dimension: multiple of 16, at most three swizzle atoms");
  static constexpr int KA = (D + 63) / 64;                 // K atoms of the model dimension
  static constexpr int A_BYTES = KA * 128 * 128;           // [128 rows x KA atoms x 128 B]
  static constexpr int H_BYTES = 128 * 128;                // one atom [128 x 64]
  static constexpr int W1_BYTES = KA * kHC * 128;          // W1^T chunk: [64 hidden rows x D k]
  static constexpr int W2_BYTES = ((D * 128 + 1023) / 1024) * 1024;  // W2^T chunk: [D rows x 64 k]
  static constexpr int W_BYTES = W1_BYTES + W2_BYTES;
  static constexpr int STAGES = 2;
  // forward: two z staging tiles + weight stages; backward: every stage also carries the z chunk (same total)
  static constexpr int REST_BYTES = STAGES * (W_BYTES + H_BYTES);
  static constexpr int BAR_BYTES = 32 * 8;
  static constexpr int H_MAX = 1024;                       // hidden units (b1 staged in SMEM)
  static constexpr int PARAM_BYTES = (H_MAX + 3 * D + 2 * kFfnRows * 4) * 4;  // b1 | b2 | ln gamma | ln beta (read once per CTA) | LN partial sums
  static constexpr int TOTAL = A_BYTES + 2 * H_BYTES + REST_BYTES + BAR_BYTES + PARAM_BYTES + 1024;
  static_assert(TOTAL <= 227 * 1024, "smem budget");
  static constexpr uint32_t TMEM_COLS = 512;               // acc1 x2 (64 columns each) + acc2 (D columns)
  static constexpr uint32_t ACC2_COL = 2 * kHC;
};

struct FfnArgs {
  const __nv_bfloat16* x;    // fwd: [M, D] module input (the residual is re-read from here)
  const float *ln_g, *ln_b;  // fwd: [D]
  const float *b1, *b2;      // fwd: [H], [D]
  float *mean, *rstd;        // fwd: [M]
  int M, H;
  float eps, factor, drop_p;
  const unsigned long long* seed_ptr;
  unsigned long long salt1, salt2;
  unsigned long long* trace;  // debug: phase timestamps of CTA 0 (tools/ffn_trace.py); null in production
  int pdl_trigger;            // 1: programmatic dependents may be scheduled as soon as this grid is resident (common.cuh)
};

// phase timestamps (globaltimer, ns) of CTA 0 -> trace[slot]; compiled in, costs one predictable branch per event when trace is null
__device__ __forceinline__ void ffn_trace(const FfnArgs& a, int slot) {
  if (a.trace && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[slot] = t;
  }
}

struct FfnMaps {
  CUtensorMap x;     // load : fwd x, bwd dout            [M, D], box [128 x 64]
  CUtensorMap w1;    // load : fwd W1^T [H, D], bwd W2 [H, D]          box [64 x 64]
  CUtensorMap w2;    // load : fwd W2^T [D, H], bwd W1 [D, H]          box [D x 64]
  CUtensorMap xn;    // store: fwd LN(x), bwd dy         [M, D], box [128 x 64]
  CUtensorMap h;     // store: fwd h, bwd dz             [M, H], box [128 x 64]
  CUtensorMap z;     // fwd store / bwd load: z          [M, H], box [128 x 64]
  CUtensorMap out;   // store: fwd out, bwd da           [M, D], box [128 x 64]
};

template <int D, int MODE>
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_kernel(const __grid_constant__ FfnMaps tm, const FfnArgs a) {
  using S = FfnSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = sm100::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;                 // x tile -> LN(x) (bwd: dout -> dy), K-major swizzled; at the end the output tile
  uint8_t* sH = sA + S::A_BYTES;      // [2][128 x 64] hidden chunk (bwd: dz chunk)
  uint8_t* sRest = sH + 2 * S::H_BYTES;
  // fwd: [2] z staging tiles, then [STAGES] weight stages;  bwd: [STAGES] (weights | z chunk)
  uint8_t* sZ = sRest;
  uint8_t* sW = (MODE == 0) ? sRest + 2 * S::H_BYTES : sRest;
  constexpr int STAGE_BYTES = (MODE == 0) ? S::W_BYTES : S::W_BYTES + S::H_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRest + S::REST_BYTES);
  uint64_t* x_full = bars + 0;
  uint64_t* a_ready = bars + 1;
  uint64_t* a_free = bars + 2;
  uint64_t* acc2_full = bars + 3;
  uint64_t* acc2_empty = bars + 4;
  uint64_t* w_full = bars + 5;             // [STAGES]
  uint64_t* w_empty = w_full + S::STAGES;  // [STAGES]
  uint64_t* acc1_full = w_empty + S::STAGES;  // [2]
  uint64_t* acc1_empty = acc1_full + 2;       // [2]
  uint64_t* h_full = acc1_empty + 2;          // [2]
  uint64_t* h_empty = h_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_empty + 2);
  float* sB1 = reinterpret_cast<float*>(bars + S::BAR_BYTES / 8);  // [H_MAX]
  float* sB2 = sB1 + S::H_MAX;                                    // [D]
  float* sG = sB2 + D;                                            // [D]
  float* sBeta = sG + D;                                          // [D]
  float* sRed = sBeta + D;                                        // [2][64 rows][4 row quarters]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = a.M, H = a.H;
  const int num_tiles = (M + kFfnRows - 1) / kFfnRows, nch = H / kHC;

  if (warp == 2 && lane == 0) {
    sm100::prefetch_tensormap(&tm.x);
    sm100::prefetch_tensormap(&tm.w1);
    sm100::prefetch_tensormap(&tm.w2);
    sm100::prefetch_tensormap(&tm.xn);
    sm100::prefetch_tensormap(&tm.h);
    sm100::prefetch_tensormap(&tm.z);
    sm100::prefetch_tensormap(&tm.out);
    sm100::mbar_init(x_full, 1);
    sm100::mbar_init(a_ready, 1);
    sm100::mbar_init(a_free, 3);  // the MMA warp (all G1 done) + the xn / dy TMA store + the output TMA store (sources read)
    sm100::mbar_init(acc2_full, 1);
    sm100::mbar_init(acc2_empty, kFfnEpiWarps);
    for (int s = 0; s < S::STAGES; ++s) {
      sm100::mbar_init(&w_full[s], 1);
      sm100::mbar_init(&w_empty[s], MODE == 0 ? 1 : 1 + kFfnEpiWarps);  // GEMM 2 has read the weights (+ bwd: the epilogue warps have read the z chunk)
    }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&acc1_full[i], 1);
      sm100::mbar_init(&acc1_empty[i], kFfnEpiWarps);
      sm100::mbar_init(&h_full[i], kFfnEpiWarps);
      sm100::mbar_init(&h_empty[i], 2);  // GEMM 2 has read the tile + its TMA stores have read it
    }
    sm100::fence_barrier_init();
  }
  if (warp == 3) sm100::tmem_alloc(tmem_slot, S::TMEM_COLS);
  // launched through launch_pdl: barrier initialisation and the TMEM allocation above overlap the tail of the previous kernel; nothing global
  // (not even the bias / LayerNorm parameters) is read before this line
  if (a.pdl_trigger) grid_launch_dependents();
  grid_dependency_wait();
  if constexpr (MODE == 0) {
    for (int i = threadIdx.x; i < a.H; i += kFfnThreads) sB1[i] = a.b1[i];
    for (int i = threadIdx.x; i < D; i += kFfnThreads) {
      sB2[i] = a.b2[i];
      sG[i] = a.ln_g[i];
      sBeta[i] = a.ln_b[i];
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) ffn_trace(a, 0);  // setup done

  if (warp == 2) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        sm100::mbar_wait(a_free, (ti & 1) ^ 1);
        sm100::mbar_arrive_expect_tx(x_full, S::KA * kFfnRows * 128);
#pragma unroll
        for (int ka = 0; ka < S::KA; ++ka) sm100::tma_load_2d(sA + ka * 16384, &tm.x, x_full, ka * 64, tile * kFfnRows);  // rows 0..63 of each atom
        for (int c = 0; c < nch; ++c, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (it / S::STAGES) & 1;
          sm100::mbar_wait(&w_empty[s], ph ^ 1);
          sm100::mbar_arrive_expect_tx(&w_full[s], S::W1_BYTES + D * 128 + (MODE == 1 ? kFfnRows * 128 : 0));
          uint8_t* w1 = sW + s * STAGE_BYTES;
#pragma unroll
          for (int ka = 0; ka < S::KA; ++ka) sm100::tma_load_2d(w1 + ka * (kHC * 128), &tm.w1, &w_full[s], ka * 64, c * kHC);  // [64 rows x 64 k]
          sm100::tma_load_2d(w1 + S::W1_BYTES, &tm.w2, &w_full[s], c * kHC, 0);                                                   // [D rows x 64 k]
          if constexpr (MODE == 1) sm100::tma_load_2d(w1 + S::W_BYTES, &tm.z, &w_full[s], c * kHC, tile * kFfnRows);               // z chunk [64 x 64]
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================================== MMA issuer (+ TMA stores of the chunk tiles)
    if (lane == 0) {
      constexpr uint32_t idesc1 = sm100::make_idesc_bf16_f32(128, kHC, 0, 0);
      constexpr uint32_t idesc2 = sm100::make_idesc_bf16_f32(128, D, 0, 0);
      const uint32_t a0 = sm100::smem_u32(sA);
      uint32_t it = 0, ti = 0;  // `it` counts hidden chunks across tiles: weight stage it % STAGES, acc1 / h buffer it & 1
      int tile_cur = 0;
      bool store_pending = false;  // a chunk-store group whose SMEM reads have not been waited for
      uint32_t pending_b = 0;
      auto gemm2 = [&](uint32_t itc, int c) {  // acc2 += h[chunk c] . W2[chunk c, :]
        const uint32_t b = itc & 1, bph = (itc >> 1) & 1, s = itc % S::STAGES;
        sm100::mbar_wait(&h_full[b], bph);
        if (c == 0) sm100::mbar_wait(acc2_empty, (ti & 1) ^ 1);  // the previous tile's output epilogue has drained acc2
        sm100::tc_fence_after();
        if (c < 12) ffn_trace(a, 40 + c);  // MMA thread: h chunk c published
        // the tiles the epilogue just published leave for global memory (saved for the backward / operands of the weight gradients)
        sm100::tma_store_2d(&tm.h, sH + b * S::H_BYTES, c * kHC, tile_cur * kFfnRows);
        if constexpr (MODE == 0) sm100::tma_store_2d(&tm.z, sZ + b * S::H_BYTES, c * kHC, tile_cur * kFfnRows);
        sm100::tma_store_commit();
        const uint32_t h0 = sm100::smem_u32(sH + b * S::H_BYTES);
        const uint32_t w2 = sm100::smem_u32(sW + s * STAGE_BYTES + S::W1_BYTES);
#pragma unroll
        for (int kk = 0; kk < kHC / 16; ++kk)
          sm100::umma_bf16(tmem_base + S::ACC2_COL, sm100::make_smem_desc(h0 + kk * 32, 16, 1024), sm100::make_smem_desc(w2 + kk * 32, 16, 1024), idesc2,
                           (c | kk) != 0 ? 1u : 0u);
        sm100::umma_commit(&h_empty[b]);
        sm100::umma_commit(&w_empty[s]);
        if (store_pending) {  // the previous chunk's stores have read their tiles: second arrival on that buffer's h_empty
          sm100::tma_store_wait_read<1>();
          sm100::mbar_arrive(&h_empty[pending_b]);
        }
        store_pending = true;
        pending_b = b;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        tile_cur = tile;
        sm100::mbar_wait(a_ready, ti & 1);
        sm100::tc_fence_after();
        for (int c = 0; c < nch; ++c, ++it) {
          const uint32_t s = it % S::STAGES, ph = (it / S::STAGES) & 1, b = it & 1, bph = (it >> 1) & 1;
          sm100::mbar_wait(&w_full[s], ph);
          if (c < 12) ffn_trace(a, 16 + c);  // MMA thread: weights of chunk c landed
          sm100::mbar_wait(&acc1_empty[b], bph ^ 1);
          sm100::tc_fence_after();
          if (c < 12) ffn_trace(a, 28 + c);  // MMA thread: issuing G1(c)
          const uint32_t w1 = sm100::smem_u32(sW + s * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const int ka = ks >> 2, kk = ks & 3;
            sm100::umma_bf16(tmem_base + b * kHC, sm100::make_smem_desc(a0 + ka * 16384 + kk * 32, 16, 1024),
                             sm100::make_smem_desc(w1 + ka * (kHC * 128) + kk * 32, 16, 1024), idesc1, ks != 0 ? 1u : 0u);
          }
          sm100::umma_commit(&acc1_full[b]);
          if (c > 0) gemm2(it - 1, c - 1);
        }
        gemm2(it - 1, nch - 1);
        ffn_trace(a, 2);  // MMA thread: last GEMM 2 issued
        sm100::umma_commit(acc2_full);
        sm100::umma_commit(a_free);  // every G1 of this tile has read sA
        sm100::tma_store_wait_read<0>();
        sm100::mbar_arrive(&h_empty[pending_b]);
        store_pending = false;
      }
    }
  } else if ((warp & 3) < 2) {
    // ===================================================================== LayerNorm + epilogues (warps 0,1,4,5,8,9,12,13)
    const int q = warp & 3;     // TMEM lane quarter of this warp: 0 or 1 = rows 0..63
    const int sub = warp >> 2;  // which 16 of a chunk's 64 hidden columns this warp takes; which quarter of a row in the prologue
    const int ew = sub * 2 + q;  // 0..7
    const int jb = sub < 2 ? sub * 5 : 10 + (sub - 2) * 4, je = jb + (sub < 2 ? 5 : 4);  // this thread's 16-byte chunks of its row (D = 144: 18)
    static_assert(D / 8 == 18, "row split of the prologue is written for 18 chunks");
    const bool has_drop = a.drop_p > 0.f;
    const float keep_scale = has_drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    const uint32_t drop_thr = dropout_threshold(a.drop_p);
    const unsigned long long seed0 = (has_drop && a.seed_ptr) ? *a.seed_ptr : 0ull;
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const int row_in_tile = q * 32 + lane;
      const int row = tile * kFfnRows + row_in_tile;
      const bool row_ok = row < M;
      if constexpr (MODE == 1) {
        // ---- backward prologue: dy = factor * Drop2'(dout) on this thread's quarter of its row, in place (scale_mask_colsum's job in round 1)
        sm100::mbar_wait(x_full, ti & 1);
        const uint32_t rowkey2b = has_drop ? dropout_rowkey(seed0 + a.salt2, (unsigned long long)row) : 0u;
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          if (j < jb || j >= je) continue;
          uint8_t* p = sA + swz_off(row_in_tile, j);
          const uint4 u = *reinterpret_cast<const uint4*>(p);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v0 = __uint_as_float(w[e] << 16) * a.factor, v1 = __uint_as_float(w[e] & 0xFFFF0000u) * a.factor;
            if (has_drop) {
              const uint32_t hsh = dropout_hash(rowkey2b, (uint32_t)(8 * j + 2 * e));
              v0 = ((hsh & 0xFFFFu) >= drop_thr) ? v0 * keep_scale : 0.f;
              v1 = ((hsh >> 16) >= drop_thr) ? v1 * keep_scale : 0.f;
            }
            o[e] = pack2(v0, v1);
          }
          *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      if constexpr (MODE == 0) {
        // ---- LayerNorm in place in the swizzled tile (keras LayerNormalization: biased variance, eps inside the sqrt): four threads per row
        // (one per lane-quarter warp group), partial sums through SMEM, two-pass statistics in fp32
        sm100::mbar_wait(x_full, ti & 1);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          if (j < jb || j >= je) continue;
          const uint4 u = *reinterpret_cast<const uint4*>(sA + swz_off(row_in_tile, j));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) sum += __uint_as_float(w[e] << 16) + __uint_as_float(w[e] & 0xFFFF0000u);
        }
        sRed[row_in_tile * 4 + sub] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float4 s4 = *reinterpret_cast<const float4*>(sRed + row_in_tile * 4);
        const float mean = (s4.x + s4.y + s4.z + s4.w) * (1.0f / D);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          if (j < jb || j >= je) continue;
          const uint4 u = *reinterpret_cast<const uint4*>(sA + swz_off(row_in_tile, j));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d0 = __uint_as_float(w[e] << 16) - mean, d1 = __uint_as_float(w[e] & 0xFFFF0000u) - mean;
            sq = fmaf(d0, d0, sq);
            sq = fmaf(d1, d1, sq);
          }
        }
        sRed[kFfnRows * 4 + row_in_tile * 4 + sub] = sq;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float4 q4 = *reinterpret_cast<const float4*>(sRed + kFfnRows * 4 + row_in_tile * 4);
        const float rstd = rsqrtf((q4.x + q4.y + q4.z + q4.w) * (1.0f / D) + a.eps);
        if (row_ok && sub == 0) {
          a.mean[row] = mean;
          a.rstd[row] = rstd;
        }
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          if (j < jb || j >= je) continue;
          const float4 g0 = *reinterpret_cast<const float4*>(sG + 8 * j), g1 = *reinterpret_cast<const float4*>(sG + 8 * j + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sBeta + 8 * j), b1v = *reinterpret_cast<const float4*>(sBeta + 8 * j + 4);
          const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
          uint8_t* p = sA + swz_off(row_in_tile, j);
          const uint4 u = *reinterpret_cast<const uint4*>(p);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = (__uint_as_float(w[e] << 16) - mean) * rstd * g[2 * e] + bb[2 * e];
            const float v1 = (__uint_as_float(w[e] & 0xFFFF0000u) - mean) * rstd * g[2 * e + 1] + bb[2 * e + 1];
            o[e] = pack2(v0, v1);
          }
          *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      {
        sm100::fence_proxy_async();  // generic-proxy writes of the A operand -> visible to tcgen05.mma / TMA (async proxy)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
          ffn_trace(a, 1);  // LayerNorm / prologue done
          sm100::mbar_arrive(a_ready);
          // fwd: xn, bwd: dy leave for the weight-gradient GEMMs (rows / columns beyond the tensor are clipped by the hardware)
#pragma unroll
          for (int ka = 0; ka < S::KA; ++ka) sm100::tma_store_2d(&tm.xn, sA + ka * 16384, ka * 64, tile * kFfnRows);
          sm100::tma_store_commit();
          sm100::tma_store_wait_read<0>();
          sm100::mbar_arrive(a_free);
        }
        __syncwarp();
      }
      // ---- hidden chunks: acc1 -> z, h  (bwd: acc1, z -> dz)
      [[maybe_unused]] const uint32_t rowkey1 = has_drop ? dropout_rowkey(seed0 + a.salt1, (unsigned long long)row) : 0u;
      const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < nch; ++c, ++it) {
        const uint32_t b = it & 1, bph = (it >> 1) & 1;
        const int col0 = c * kHC + sub * 16;  // first hidden column of this thread's 16
        uint32_t zp[8], hp[8];
        if constexpr (MODE == 1) {  // the saved pre-activation of this thread's 32 hidden units: arrived by TMA with the weights of the chunk
          const uint32_t s = it % S::STAGES, ph = (it / S::STAGES) & 1;
          sm100::mbar_wait(&w_full[s], ph);
          const uint8_t* zt = sW + s * STAGE_BYTES + S::W_BYTES;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint4 u = *reinterpret_cast<const uint4*>(zt + swz_off(row_in_tile, sub * 2 + j));
            zp[4 * j] = u.x, zp[4 * j + 1] = u.y, zp[4 * j + 2] = u.z, zp[4 * j + 3] = u.w;
          }
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(&w_empty[s]);
        }
        sm100::mbar_wait(&acc1_full[b], bph);
        sm100::tc_fence_after();
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 52 + c);  // epilogue warp 0: accumulator of chunk c ready
        uint32_t r[1][16];
        sm100::tmem_ld_32x32b_x16(tmem_row + b * kHC + sub * 16, r[0]);
        sm100::tmem_ld_wait();
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 88 + c);  // epilogue warp 0: accumulator in registers
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&acc1_empty[b]);  // acc1[b] may take chunk c + 2
#pragma unroll
        for (int ci = 0; ci < 1; ++ci) {
          float v[16];
          if constexpr (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(sB1 + col0 + ci * 16 + 4 * j);
              v[4 * j] = __uint_as_float(r[ci][4 * j]) + b4.x;
              v[4 * j + 1] = __uint_as_float(r[ci][4 * j + 1]) + b4.y;
              v[4 * j + 2] = __uint_as_float(r[ci][4 * j + 2]) + b4.z;
              v[4 * j + 3] = __uint_as_float(r[ci][4 * j + 3]) + b4.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) zp[ci * 8 + j] = pack2(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float s0, s1;
              sigmoid2_(v[j], v[j + 1], s0, s1);
              v[j] *= s0;
              v[j + 1] *= s1;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {  // * d swish / dz at the saved pre-activation
              const uint32_t zw = zp[ci * 8 + (j >> 1)];
              const float z0 = __uint_as_float(zw << 16), z1 = __uint_as_float(zw & 0xFFFF0000u);
              float s0, s1;
              sigmoid2_(z0, z1, s0, s1);
              v[j] = __uint_as_float(r[ci][j]) * (s0 * (1.0f + z0 * (1.0f - s0)));
              v[j + 1] = __uint_as_float(r[ci][j + 1]) * (s1 * (1.0f + z1 * (1.0f - s1)));
            }
          }
          if (has_drop) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const uint32_t hsh = dropout_hash(rowkey1, (uint32_t)(col0 + ci * 16 + j));
              v[j] = ((hsh & 0xFFFFu) >= drop_thr) ? v[j] * keep_scale : 0.f;
              v[j + 1] = ((hsh >> 16) >= drop_thr) ? v[j + 1] * keep_scale : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) hp[ci * 8 + j] = pack2(v[2 * j], v[2 * j + 1]);
        }
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 64 + c);  // epilogue warp 0: chunk c computed
        sm100::mbar_wait(&h_empty[b], bph ^ 1);  // GEMM 2 of chunk c - 2 and its tile stores have read these buffers
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 76 + c);  // epilogue warp 0: buffer free
        uint8_t* hb = sH + b * S::H_BYTES;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(hb + swz_off(row_in_tile, sub * 2 + j)) = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
        if constexpr (MODE == 0) {
          uint8_t* zb = sZ + b * S::H_BYTES;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(zb + swz_off(row_in_tile, sub * 2 + j)) = make_uint4(zp[4 * j], zp[4 * j + 1], zp[4 * j + 2], zp[4 * j + 3]);
        }
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 100 + c);  // epilogue warp 0: tiles written
        sm100::fence_proxy_async();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&h_full[b]);
        if (ew == 0 && lane == 0 && c < 12) ffn_trace(a, 112 + c);  // epilogue warp 0: published
      }
      // ---- output: acc2 -> + b2 -> dropout -> residual
      [[maybe_unused]] const uint32_t rowkey2 = has_drop ? dropout_rowkey(seed0 + a.salt2, (unsigned long long)row) : 0u;
      constexpr int NG = D / 16, G_MID = NG - (NG * 3) / 4;  // the 16-column groups of the output, split over the 4 warps of a lane quarter
      const int g0 = (NG * sub) / 4, ng = (NG * (sub + 1)) / 4 - g0;
      // the residual x of this thread's columns is requested before the wait for the accumulator (its latency hides behind GEMM 2)
      [[maybe_unused]] uint4 xres[G_MID][2];
      if constexpr (MODE == 0) {
#pragma unroll
        for (int gi = 0; gi < G_MID; ++gi)
          if (gi < ng && row_ok) {
            const uint4* xp = reinterpret_cast<const uint4*>(a.x + (size_t)row * D + (g0 + gi) * 16);
            xres[gi][0] = ld_stream_u4(xp);
            xres[gi][1] = ld_stream_u4(xp + 1);
          }
      }
      sm100::mbar_wait(acc2_full, ti & 1);
      sm100::tc_fence_after();
      if (ew == 0 && lane == 0) ffn_trace(a, 3);  // output accumulator ready
#pragma unroll
      for (int gi = 0; gi < G_MID; ++gi) {
        if (gi >= ng) break;
        const int g = g0 + gi;
        uint32_t racc[16];  // one 16-column group at a time: the 576-thread CTA has no registers for the whole row slice
        sm100::tmem_ld_32x32b_x16(tmem_row + S::ACC2_COL + g * 16, racc);
        sm100::tmem_ld_wait();
        float v[16];
        if constexpr (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(racc[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(sB2 + g * 16 + 4 * j);
            v[4 * j] = __uint_as_float(racc[4 * j]) + b4.x;
            v[4 * j + 1] = __uint_as_float(racc[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(racc[4 * j + 2]) + b4.z;
            v[4 * j + 3] = __uint_as_float(racc[4 * j + 3]) + b4.w;
          }
          if (has_drop) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const uint32_t hsh = dropout_hash(rowkey2, (uint32_t)(g * 16 + j));
              v[j] = ((hsh & 0xFFFFu) >= drop_thr) ? v[j] * keep_scale : 0.f;
              v[j + 1] = ((hsh >> 16) >= drop_thr) ? v[j + 1] * keep_scale : 0.f;
            }
          }
          if (row_ok) {
            const uint32_t w[8] = {xres[gi][0].x, xres[gi][0].y, xres[gi][0].z, xres[gi][0].w, xres[gi][1].x, xres[gi][1].y, xres[gi][1].z, xres[gi][1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[2 * j] = __uint_as_float(w[j] << 16) + a.factor * v[2 * j];
              v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u) + a.factor * v[2 * j + 1];
            }
          }
        }
        // the output tile takes the place of the A operand (every GEMM 1 of the tile has completed: acc2_full follows them)
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = pack2(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(sA + swz_off(row_in_tile, 2 * g)) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(sA + swz_off(row_in_tile, 2 * g + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(acc2_empty);  // this warp's slice of acc2 has been read
      sm100::fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ew == 0 && lane == 0) {
#pragma unroll
        for (int ka = 0; ka < S::KA; ++ka) sm100::tma_store_2d(&tm.out, sA + ka * 16384, ka * 64, tile * kFfnRows);
        sm100::tma_store_commit();
        sm100::tma_store_wait_read<0>();
        sm100::mbar_arrive(a_free);
        ffn_trace(a, 4);  // output tile handed to the TMA
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 3) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int D, int MODE>
static int launch_ffn(const FfnMaps& tm, const FfnArgs& a, cudaStream_t stream) {
  using S = FfnSmem<D>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(ffn_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  const int tiles = (a.M + kFfnRows - 1) / kFfnRows;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  FfnArgs a2 = a;
  a2.pdl_trigger = pdl_trigger_enabled() ? 1 : 0;
  launch_pdl(ffn_kernel<D, MODE>, dim3(grid), dim3(kFfnThreads), S::TOTAL, stream, tm, a2);
  count_launch();
  return check_launch();
}

// x / xn / out share [M, D]; z / h share [M, H]
static bool ffn_maps(FfnMaps* tm, const void* x, const void* w1, const void* w2, const void* xn, const void* h, const void* z, const void* out, int M, int D,
                     int H) {
  bool ok = make_tmap_bf16_2d(&tm->x, x, (uint64_t)M, (uint64_t)D, (uint64_t)D, kFfnRows, 64);
  ok = ok && make_tmap_bf16_2d(&tm->w1, w1, (uint64_t)H, (uint64_t)D, (uint64_t)D, kHC, 64);          // [H, D]: rows = hidden units, K = D
  ok = ok && make_tmap_bf16_2d(&tm->w2, w2, (uint64_t)D, (uint64_t)H, (uint64_t)H, (uint32_t)D, 64);  // [D, H]: rows = model dim, K = hidden
  ok = ok && make_tmap_bf16_2d(&tm->xn, xn, (uint64_t)M, (uint64_t)D, (uint64_t)D, kFfnRows, 64);
  ok = ok && make_tmap_bf16_2d(&tm->h, h, (uint64_t)M, (uint64_t)H, (uint64_t)H, kFfnRows, 64);
  ok = ok && make_tmap_bf16_2d(&tm->z, z, (uint64_t)M, (uint64_t)H, (uint64_t)H, kFfnRows, 64);
  ok = ok && make_tmap_bf16_2d(&tm->out, out, (uint64_t)M, (uint64_t)D, (uint64_t)D, kFfnRows, 64);
  return ok;
}

}  // namespace tfasr

using namespace tfasr;

static unsigned long long* g_ffn_trace = nullptr;
// debug: device buffer of 128 u64 that the next launches' CTA 0 fills with phase timestamps (null = off)
extern "C" void tfasr_ffn_set_trace(unsigned long long* dev_buffer) { g_ffn_trace = dev_buffer; }

extern "C" int tfasr_ffn_supported(int32_t D, int32_t H) { return (D == 144 && H >= 64 && H % 64 == 0 && H <= FfnSmem<144>::H_MAX) ? 1 : 0; }

extern "C" int tfasr_ffn_fwd_bf16(const void* x, const float* ln_g, const float* ln_b, const void* w1t, const float* b1, const void* w2t,
                                  const float* b2, void* out, void* xn, void* z, void* h, float* mean, float* rstd, int32_t M, int32_t D, int32_t H,
                                  float eps, float factor, float drop_p, const unsigned long long* seed_ptr, unsigned long long salt1,
                                  unsigned long long salt2, tfasr_stream_t stream) {
  if (!x || !ln_g || !ln_b || !w1t || !b1 || !w2t || !b2 || !out || !xn || !z || !h || !mean || !rstd || M < 0) return TFASR_INVALID_VALUE;
  if (!tfasr_ffn_supported(D, H) || !(drop_p >= 0.f && drop_p < 1.f)) return TFASR_INVALID_VALUE;
  if (M == 0) return TFASR_OK;
  FfnMaps tm;
  if (!ffn_maps(&tm, x, w1t, w2t, xn, h, z, out, M, D, H)) return TFASR_INVALID_VALUE;
  FfnArgs a{};
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.ln_g = ln_g;
  a.ln_b = ln_b;
  a.b1 = b1;
  a.b2 = b2;
  a.mean = mean;
  a.rstd = rstd;
  a.M = M;
  a.H = H;
  a.eps = eps;
  a.factor = factor;
  a.drop_p = drop_p;
  a.seed_ptr = seed_ptr;
  a.salt1 = salt1;
  a.salt2 = salt2;
  a.trace = g_ffn_trace;
  return launch_ffn<144, 0>(tm, a, static_cast<cudaStream_t>(stream));
}

extern "C" int tfasr_ffn_bwd_bf16(const void* dout, const void* z, const void* w2, const void* w1, void* dy, void* dz, void* da, int32_t M, int32_t D,
                                  int32_t H, float factor, float drop_p, const unsigned long long* seed_ptr, unsigned long long salt1,
                                  unsigned long long salt2, tfasr_stream_t stream) {
  if (!dout || !z || !w2 || !w1 || !dy || !dz || !da || M < 0) return TFASR_INVALID_VALUE;
  if (!tfasr_ffn_supported(D, H) || !(drop_p >= 0.f && drop_p < 1.f)) return TFASR_INVALID_VALUE;
  if (M == 0) return TFASR_OK;
  FfnMaps tm;
  // W2 [H, D]: ds = dy . W2^T (rows = hidden units);  W1 [D, H]: da = dz . W1^T (rows = model dim)
  if (!ffn_maps(&tm, dout, w2, w1, dy, dz, z, da, M, D, H)) return TFASR_INVALID_VALUE;
  FfnArgs a{};
  a.M = M;
  a.H = H;
  a.factor = factor;
  a.drop_p = drop_p;
  a.seed_ptr = seed_ptr;
  a.salt1 = salt1;
  a.salt2 = salt2;
  a.trace = g_ffn_trace;
  return launch_ffn<144, 1>(tm, a, static_cast<cudaStream_t>(stream));
}
