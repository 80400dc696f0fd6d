// tcgen05 GEMM for sm_100a: C[M,N] = epilogue(A[M,K] . B[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every dense contraction the reference hands to cuBLAS through Keras Dense / EinsumDense / 1x1 Conv1D
// (tensorflow_asr/models/encoders/conformer.py:72-87 FFN, :293-352 pointwise convs, :591 linear;
//  models/layers/multihead_attention.py:628-654 projections; models/transducer/base_transducer.py:285-292 joint) and their
// backward (dgrad: dX = dY . W, a TN GEMM on the transposed weight copy; wgrad: dW = X^T . dY with MN-major operands).
//
// One CTA per 128 x BN output tile, 192 threads, warp-specialised:
//   warp 0    TMA producer   cp.async.bulk.tensor (128B swizzle) into a STAGES-deep ring, mbarrier full/empty pairs
//   warp 1    MMA issuer     one elected lane issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16) — accumulator in TMEM
//   warps 2-5 epilogue       tcgen05.ld 32x32b (one TMEM lane = one output row per thread) -> bias / activation / GLU /
//                            residual / dropout -> 128-bit global stores
// Operands are K-major ("TN") by default; A_MN / B_MN select MN-major smem descriptors so that X^T . dY needs no transposes.
//
// Roofline: tensor-pipe bound for the joint (M = 520 000, K = 320, N = 1000); the encoder GEMMs (M = 8000, K,N <= 576) are
// launch/latency bound — see DESIGN.md.
#pragma once
#include "common.cuh"
#include "sm100.cuh"
#include "dropout.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

namespace tfasr {

bool make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);
bool make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

#ifdef TFASR_GEMM_TRACE
// debug build only: per-CTA phase timestamps (globaltimer ns) -> trace[blockIdx.x * 16 + slot]
extern unsigned long long* g_gemm_trace_host;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot)                                                      \
  do {                                                                   \
    if (epi.trace) epi.trace[blockIdx.x * 16 + (slot)] = gtime();        \
  } while (0)
#else
#define TRACE(slot) \
  do {              \
  } while (0)
#endif

// ------------------------------------------------------------------------------------------------ epilogue description
struct GemmEpi {
  void* c;              // output [M, ldc]  (bf16 or f32); GLU: [M, N/2]
  const float* bias;    // [N] or null
  void* aux;            // optional pre-activation copy (bf16 [M, ld_aux]) — what backward needs of swish / GLU
  const void* residual; // optional bf16 [M, ld_res]: out = residual + res_scale * value
  const void* mul;      // optional bf16 [M, ld_mul]: value *= f'(mul) per `mul_mode` (backward epilogues)
  float* colsum;        // optional f32 [N]: atomically accumulates the column sums of the final value (bias gradients)
  int ldc, ld_aux, ld_res, ld_mul;
  int c_f32;            // 1: f32 output
  int act;              // 0 none, 1 swish, 2 relu, 3 tanh
  int glu;              // 1: interleaved (a, b) column pairs -> a * sigmoid(b)
  int mul_mode;         // 0 none, 1: * swish'(mul), 2: * (1 - mul^2)  [tanh'], 3: * mul
  float res_scale;      // residual factor (0.5 for the macaron FFN)
  float alpha;          // scales the accumulator before bias (1 by default)
  float drop_p;         // dropout probability (0 = off); mask = hash(*seed_ptr + salt, row * N + col)
  const unsigned long long* seed_ptr;  // device counter, bumped once per step (CUDA-graph replay safe)
  unsigned long long salt;
  int atomic_out;       // 1: atomicAdd into f32 C (split-K partial sums / gradient accumulation)
  int tma_store;        // 1: C leaves through SMEM staging + cp.async.bulk.tensor stores (coalesced, clipped by hardware)
  int tma_side;         // 1: residual, 2: mul arrives through a TMA load into the staging buffer ahead of the accumulator
  int tma_aux;          // 1: the pre-activation copy leaves through its own staging buffer + TMA store
  int pdl_trigger;      // 1: let programmatic dependents be scheduled as soon as this (persistent) grid is resident
  float* rowstats;      // feature class 8 only: [M, rs_parts, 2] f32 = per row and column slice (max, sum exp(v - max)) of the final values
  int rs_parts;         // 2 * ceil(N / BN): every tile contributes two slices per row (the two epilogue warps of a TMEM lane quarter)
#ifdef TFASR_GEMM_TRACE
  unsigned long long* trace;
#endif
};

// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5 : one MUFU op (tanh.approx) instead of ex2 + rcp
__device__ __forceinline__ float sigmoidf_(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// two sigmoids per MUFU op: tanh.approx.f16x2 (11-bit mantissa: error 2^-12 on the sigmoid, far below the bf16 rounding of the output);
// the epilogue of the activation GEMMs is MUFU-bound otherwise (16 tanh per 16 columns, 8 warps)
__device__ __forceinline__ void sigmoid2_(float a, float b, float& sa, float& sb) {
  const __half2 h = __floats2half2_rn(0.5f * a, 0.5f * b);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&h)));
  const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&t));
  sa = fmaf(0.5f, tf.x, 0.5f);
  sb = fmaf(0.5f, tf.y, 0.5f);
}

constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kBM = 128;
constexpr int kBK = 64;

// Persistent, warp-specialised kernel: one CTA per SM walks output tiles (n fastest, so the N-tiles of one M-tile run
// back to back and A is read from HBM once); the smem ring and the two TMEM accumulator buffers run continuously across
// tiles, so the epilogue of tile i overlaps the TMA + MMA of tile i+1.
template <int BN>
struct GemmSmem {
  static constexpr int STAGES = (BN >= 256) ? 3 : ((BN >= 160) ? 4 : ((BN >= 128) ? 5 : 6));
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_OUT_BYTES = 8 /*epilogue warps*/ * 2 /*buffers*/ * 32 /*rows*/ * 128 /*bytes*/;
  static constexpr int BIAS_BYTES = 2 * BN * 4;  // bias slice of the tile, double-buffered with the accumulators
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 32) * 8 + 16;
  static constexpr int TOTAL = STAGES * (A_BYTES + B_BYTES) + STAGE_OUT_BYTES + BIAS_BYTES + 1024 /*align slack*/ + BAR_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);  // two accumulators
  static_assert(TOTAL <= 227 * 1024, "smem budget");
};


__device__ __forceinline__ void load16(const __nv_bfloat16* p, float* f) {  // 16 bf16 (32 B, 16-byte aligned) -> f32
  const uint4 u0 = *reinterpret_cast<const uint4*>(p);
  const uint4 u1 = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float* f) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(p + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// byte offset of 16-byte chunk `chunk` (8 bf16) of row `row` inside a K-major SWIZZLE_128B operand tile whose 64-column atoms are
// [128 rows x 128 B] (what TMA writes for a [128 x 64] bf16 box, and what the UMMA descriptors of this library read)
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) { return (uint32_t)(chunk >> 3) * (128 * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// Epilogue feature classes (compile-time): what a class does not name is not in the instruction stream.
//   0 lean: alpha, bias, bf16 / f32 output (TMA store, TMA reduce-add, or direct)      -> projections, dgrad, wgrad, joint logits
//   1 fwd : + activation, pre-activation copy (aux), GLU                               -> FFN-1, pointwise conv 1
//   2 dr  : + dropout, residual (side input by TMA)                                    -> FFN-2, attention output, pointwise conv 2, linear
//   4 bwd : + multiply by f'(saved) (side input by TMA), column sums (bias gradients)  -> backward epilogues
//   3 = 1|2 (FFN-1: swish + aux + dropout),  6 = 2|4 (FFN backward: dropout mask * swish'),  7 all
//   8 rs  : lean + per-row softmax statistics of the output (online max / sum-exp per column slice) -> joint logits feeding the RNN-T
//           loss without a separate log-sum-exp pass over the 2 GB of logits (first step)
template <uint32_t FEAT>
struct EpiFeat {
  static constexpr bool fwd = (FEAT & 1u) != 0, dr = (FEAT & 2u) != 0, bwd = (FEAT & 4u) != 0, rs = (FEAT & 8u) != 0;
};

template <int BN, bool A_MN, bool B_MN, uint32_t FEAT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmX, const GemmEpi epi, int M, int N, int K,
                 int splits) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = sm100::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S::STAGES * S::A_BYTES;
  uint8_t* sOut = smem + S::STAGES * (S::A_BYTES + S::B_BYTES);  // [8 warps][8 KB: 2 f32 / 4 bf16 group buffers], 1024-aligned
  float* sBias = reinterpret_cast<float*>(sOut + S::STAGE_OUT_BYTES);  // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + S::STAGE_OUT_BYTES + S::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* side_bar = tmem_empty + 2;          // [8 warps][4 groups]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(side_bar + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int total_k = (K + kBK - 1) / kBK;
  const int per_split = (total_k + splits - 1) / splits;
  const int num_tiles = m_tiles * n_tiles * splits;

  if (threadIdx.x == 0) TRACE(0);
  if (warp == 0 && lane == 0) {
    sm100::prefetch_tensormap(&tmA);
    sm100::prefetch_tensormap(&tmB);
    if (epi.tma_store) sm100::prefetch_tensormap(&tmC);
    if (epi.tma_side) sm100::prefetch_tensormap(&tmR);
    if (epi.tma_aux) sm100::prefetch_tensormap(&tmX);
    for (int i = 0; i < 32; ++i) sm100::mbar_init(&side_bar[i], 1);
    for (int s = 0; s < S::STAGES; ++s) {
      sm100::mbar_init(&full_bar[s], 1);
      sm100::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&tmem_full[i], 1);
      sm100::mbar_init(&tmem_empty[i], 8);  // one arrival per epilogue warp
    }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc(tmem_slot, S::TMEM_COLS);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (epi.pdl_trigger) sm100::grid_launch_dependents();  // dependents may become resident now; they block until this grid has completed
  sm100::grid_dependency_wait();  // everything above touched only this CTA's shared memory and TMEM; global memory from here on
  if (threadIdx.x == 0) TRACE(1);

  // tile -> (split, m_blk, n_blk), n fastest
  auto decode = [&](int tile, int& m_blk, int& n_blk, int& k_begin, int& num_k) {
    n_blk = tile % n_tiles;
    const int t2 = tile / n_tiles;
    m_blk = t2 % m_tiles;
    const int sp = t2 / m_tiles;
    k_begin = sp * per_split;
    num_k = max(0, min(total_k, k_begin + per_split) - k_begin);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk, k_begin, num_k;
        decode(tile, m_blk, n_blk, k_begin, num_k);
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (it / S::STAGES) & 1;
          sm100::mbar_wait(&empty_bar[s], ph ^ 1);
          sm100::mbar_arrive_expect_tx(&full_bar[s], S::A_BYTES + S::B_BYTES);
          uint8_t* a = sA + s * S::A_BYTES;
          uint8_t* b = sB + s * S::B_BYTES;
          const int kc = (k_begin + kb) * kBK;
          if constexpr (!A_MN) {
            sm100::tma_load_2d(a, &tmA, &full_bar[s], kc, m_blk * kBM);  // [128 rows(m) x 64 k]
          } else {
            // source [K, M]: two boxes of [64 k-rows x 64 m-cols]; each box = 64 rows of 128 B
            sm100::tma_load_2d(a, &tmA, &full_bar[s], m_blk * kBM, kc);
            sm100::tma_load_2d(a + kBK * 128, &tmA, &full_bar[s], m_blk * kBM + 64, kc);
          }
          if constexpr (!B_MN) {
            sm100::tma_load_2d(b, &tmB, &full_bar[s], kc, n_blk * BN);  // [BN rows(n) x 64 k]
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) sm100::tma_load_2d(b + j * kBK * 128, &tmB, &full_bar[s], n_blk * BN + j * 64, kc);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = sm100::make_idesc_bf16_f32(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
        int m_blk, n_blk, k_begin, num_k;
        decode(tile, m_blk, n_blk, k_begin, num_k);
        const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        sm100::mbar_wait(&tmem_empty[buf], aph ^ 1);  // the epilogue has drained this accumulator
        sm100::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (it / S::STAGES) & 1;
          sm100::mbar_wait(&full_bar[s], ph);
          sm100::tc_fence_after();
          if (it == 0) TRACE(2);
          const uint32_t a0 = sm100::smem_u32(sA + s * S::A_BYTES);
          const uint32_t b0 = sm100::smem_u32(sB + s * S::B_BYTES);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // K-major: 16 k = 32 B inside the 128 B swizzle row.  MN-major: 16 k-rows of 128 B = 2048 B.
            const uint64_t ad = A_MN ? sm100::make_smem_desc(a0 + k * 2048, kBK * 128, 1024) : sm100::make_smem_desc(a0 + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? sm100::make_smem_desc(b0 + k * 2048, kBK * 128, 1024) : sm100::make_smem_desc(b0 + k * 32, 16, 1024);
            sm100::umma_bf16(tmem_d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          sm100::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        sm100::umma_commit(&tmem_full[buf]);  // accumulator complete (fires immediately when no MMA was issued)
        if (acc_it < 2) TRACE(3 + acc_it);
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    // Two warps per TMEM lane quarter split the tile's 32-column groups evenly.  The code is specialised at compile time on the
    // feature class FEAT (what is not compiled in costs nothing: the generic body is ~250 instructions per 16 columns and made the
    // epilogue, not the MMA, the critical path of the small encoder GEMMs).
    using F = EpiFeat<FEAT>;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int ew = warp - 2;  // epilogue warp index 0..7
    const int half = ew >> 2;
    const float alpha = epi.alpha;
    const bool c_f32 = epi.c_f32 != 0;
    const bool tma_out = epi.tma_store != 0;
    const int side = (F::dr || F::bwd) ? (tma_out ? epi.tma_side : 0) : 0;  // 1 residual / 2 mul staged by TMA ahead of the accumulator
    const int gbytes = c_f32 ? 4096 : 2048;  // one group = 32 rows x 32 columns: 128-byte (f32, 128B swizzle) or 64-byte (bf16, 64B swizzle) rows
    const int nbuf_mask = c_f32 ? 1 : 3;     // staging buffers per warp: 2 (f32) or 4 (bf16) -> every bf16 group of a tile has its own
    const int aux_tma = F::fwd ? (tma_out ? epi.tma_aux : 0) : 0;  // the aux copy takes every second buffer
    uint8_t* sbuf0 = sOut + ew * 8192;
    const int etid = threadIdx.x - 64;  // 0..255 over the epilogue warps
    [[maybe_unused]] const bool has_drop = F::dr && epi.drop_p > 0.f;
    [[maybe_unused]] const float keep_scale = has_drop ? 1.0f / (1.0f - epi.drop_p) : 1.0f;
    [[maybe_unused]] const uint32_t drop_thr = dropout_threshold(epi.drop_p);
    [[maybe_unused]] const unsigned long long seed = (has_drop && epi.seed_ptr) ? (*epi.seed_ptr + epi.salt) : epi.salt;
    [[maybe_unused]] const float res_scale = epi.res_scale;
    const bool vec16 = ((epi.ldc & 7) == 0) || (c_f32 && (epi.ldc & 3) == 0);
    uint32_t side_phase = 0;  // bit g: parity of side_bar[ew][g]
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      int m_blk, n_blk, k_begin, num_k;
      decode(tile, m_blk, n_blk, k_begin, num_k);
      const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
      const int row0 = m_blk * kBM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      [[maybe_unused]] const uint32_t rowkey = has_drop ? dropout_rowkey(seed, (unsigned long long)row) : 0u;
      const uint32_t tmem_row = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      const int n_cols = min(BN, N - n_blk * BN);  // valid columns of this tile
      const int n_groups = (n_cols + 31) >> 5;
      const int g_mid = (n_groups + 1) >> 1;
      const int g0 = half ? g_mid : 0, g1 = half ? n_groups : g_mid;
      const int my_groups = g1 - g0;  // <= 4
      const int c_begin = g0 * 32, c_end = min(n_cols, g1 * 32);
      const int valid_rows = max(0, min(32, M - row0));
      const bool has_work = num_k > 0 && my_groups > 0;
      // ---- while the accumulator is still being produced: bias slice -> SMEM, side input (residual / mul) -> staging buffers by TMA
      float* bias_s = sBias + buf * BN;
      if (etid < BN) bias_s[etid] = (epi.bias && n_blk * BN + etid < N) ? __ldg(epi.bias + n_blk * BN + etid) : 0.f;
      if (tma_out) {
        if (lane == 0) sm100::tma_store_wait_read<0>();  // all staging buffers are free again
        __syncwarp();
        if (side && has_work && lane == 0) {
          for (int gi = 0; gi < my_groups; ++gi) {
            uint64_t* sb = &side_bar[ew * 4 + gi];
            sm100::mbar_arrive_expect_tx(sb, 2048);
            sm100::tma_load_2d(sbuf0 + gi * 2048, &tmR, sb, n_blk * BN + c_begin + gi * 32, row0);
          }
        }
      }
      [[maybe_unused]] float rs_m = -CUDART_INF_F, rs_s = 0.f;  // F::rs: this thread's row, this warp's column slice of the tile
      sm100::mbar_wait(&tmem_full[buf], aph);
      sm100::tc_fence_after();
      if (warp == 2 && lane == 0 && acc_it < 2) TRACE(5 + 2 * acc_it);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // bias slice visible to the 8 epilogue warps
#pragma unroll 1
      for (int gi = 0; gi < my_groups && num_k > 0; ++gi) {
        const int cg0 = c_begin + gi * 32;
        uint8_t* sbuf = sbuf0 + ((aux_tma ? 2 * gi : gi) & nbuf_mask) * gbytes;
        [[maybe_unused]] uint8_t* abuf = sbuf0 + ((2 * gi + 1) & 3) * 2048;
        uint32_t r[2][16];
        const bool two = cg0 + 16 < c_end;
        sm100::tmem_ld_32x32b_x16(tmem_row + (uint32_t)cg0, r[0]);
        if (two) sm100::tmem_ld_32x32b_x16(tmem_row + (uint32_t)(cg0 + 16), r[1]);
        if (tma_out && gi >= 2 && (c_f32 || aux_tma)) {
          // buffers come round again: the stores issued two groups ago must have finished reading them
          if (lane == 0) {
            if (aux_tma) sm100::tma_store_wait_read<2>();
            else sm100::tma_store_wait_read<1>();
          }
          __syncwarp();
        }
        if (side) {
          sm100::mbar_wait(&side_bar[ew * 4 + gi], (side_phase >> gi) & 1u);
          side_phase ^= 1u << gi;
        }
        sm100::tmem_ld_wait();
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          if (ci == 1 && !two) break;
          const int c0 = cg0 + ci * 16;
          const int n0 = n_blk * BN + c0;
          const bool full = (c0 + 16 <= n_cols) && vec16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[ci][j]);
          if (alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= alpha;
          }
          {
            const float4* gb = reinterpret_cast<const float4*>(bias_s + c0);  // broadcast reads; zeros when there is no bias
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = gb[j];
              v[4 * j] += b4.x;
              v[4 * j + 1] += b4.y;
              v[4 * j + 2] += b4.z;
              v[4 * j + 3] += b4.w;
            }
          }
          if constexpr (F::rs) {
            if (!c_f32) {  // bf16 output: the statistics describe the values as stored, so that exp(stored - lse) sums to one exactly
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
            }
            // online log-sum-exp over the valid columns of this 16-column chunk; a term is FFMA + ex2 + FADD (base-2 exponent with the
            // running maximum folded into the FFMA), and full chunks carry no per-column predicates
            const bool full = c0 + 16 <= n_cols;
            float cm = -CUDART_INF_F;
            if (full) {
#pragma unroll
              for (int j = 0; j < 16; ++j) cm = fmaxf(cm, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < n_cols) cm = fmaxf(cm, v[j]);
            }
            if (cm > -CUDART_INF_F) {
              const float mn = fmaxf(rs_m, cm);
              const float nm = -mn * kLog2e;
              float acc = 0.f;
              if (full) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += fast_ex2(fmaf(v[j], kLog2e, nm));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < n_cols) acc += fast_ex2(fmaf(v[j], kLog2e, nm));
              }
              rs_s = fmaf(rs_s, fast_ex2((rs_m - mn) * kLog2e), acc);  // rs_m = -inf on the first chunk: 2^-inf = 0
              rs_m = mn;
            }
          }
          if constexpr (F::fwd) {
            if (aux_tma) {
              uint8_t* arow = abuf + lane * 64;
              const int swz = (lane >> 1) & 3;
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = pack2(v[2 * j], v[2 * j + 1]);
              *reinterpret_cast<uint4*>(arow + (((2 * ci) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(arow + (((2 * ci + 1) ^ swz) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else if (epi.aux && row_ok) {
              __nv_bfloat16* aux_row = reinterpret_cast<__nv_bfloat16*>(epi.aux) + (size_t)row * epi.ld_aux;
              if (full && (epi.ld_aux & 7) == 0) store16(aux_row + n0, v);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < n_cols) aux_row[n0 + j] = __float2bfloat16(v[j]);
              }
            }
            if (epi.glu) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                float s0, s1;
                sigmoid2_(v[2 * j + 1], v[2 * j + 3], s0, s1);
                o[j] = v[2 * j] * s0;
                o[j + 1] = v[2 * j + 2] * s1;
              }
              if (tma_out) {  // 16 output columns per group: 32-byte rows, no swizzle
                *reinterpret_cast<uint4*>(sbuf + lane * 32 + ci * 16) = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
              } else if (row_ok) {
                __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.c) + (size_t)row * epi.ldc + (n0 >> 1);
                if (full) {
                  uint32_t pk[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 t = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t*>(&t);
                  }
                  *reinterpret_cast<uint4*>(cp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (c0 + 2 * j + 1 < n_cols) cp[j] = __float2bfloat16(o[j]);
                }
              }
              continue;
            }
            const int act = epi.act;
            if (act == 1) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float s0, s1;
                sigmoid2_(v[j], v[j + 1], s0, s1);
                v[j] *= s0;
                v[j + 1] *= s1;
              }
            } else if (act == 2) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (act == 3) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = tanhf(v[j]);
            }
          }
          // side input staged by TMA (bf16 output only): this thread's 16 bf16 sit exactly where its 16 outputs will go
          [[maybe_unused]] float zs[16];
          if constexpr (F::dr || F::bwd) {
            if (side) {
              const uint8_t* srow = sbuf + lane * 64;
              const int swz = (lane >> 1) & 3;
              const uint4 u0 = *reinterpret_cast<const uint4*>(srow + (((2 * ci) ^ swz) << 4));
              const uint4 u1 = *reinterpret_cast<const uint4*>(srow + (((2 * ci + 1) ^ swz) << 4));
              const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                zs[2 * j] = __uint_as_float(w[j] << 16);
                zs[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
              }
            }
          }
          if constexpr (F::bwd) {
            if (epi.mul && (row_ok || side == 2)) {
              const __nv_bfloat16* mul_row = reinterpret_cast<const __nv_bfloat16*>(epi.mul) + (size_t)row * epi.ld_mul;
              float z[16];
              if (side == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = zs[j];
              } else if (full && (epi.ld_mul & 7) == 0) {
                load16(mul_row + n0, z);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = (c0 + j < n_cols) ? __bfloat162float(mul_row[n0 + j]) : 0.f;
              }
              const int mul_mode = epi.mul_mode;
              if (mul_mode == 1) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float s0, s1;
                  sigmoid2_(z[j], z[j + 1], s0, s1);
                  v[j] *= s0 * (1.0f + z[j] * (1.0f - s0));  // d swish / dz
                  v[j + 1] *= s1 * (1.0f + z[j + 1] * (1.0f - s1));
                }
              } else if (mul_mode == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= 1.0f - z[j] * z[j];  // d tanh / dz from the tanh value
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= z[j];
              }
            }
          }
          if constexpr (F::dr) {
            if (has_drop) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const uint32_t hsh = dropout_hash(rowkey, (uint32_t)(n0 + j));  // n0 % 16 == 0: one hash per column pair
                v[j] = ((hsh & 0xFFFFu) >= drop_thr) ? v[j] * keep_scale : 0.f;
                v[j + 1] = ((hsh >> 16) >= drop_thr) ? v[j + 1] * keep_scale : 0.f;
              }
            }
            if (epi.residual && (row_ok || side == 1)) {
              const __nv_bfloat16* res_row = reinterpret_cast<const __nv_bfloat16*>(epi.residual) + (size_t)row * epi.ld_res;
              float z[16];
              if (side == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = zs[j];
              } else if (full && (epi.ld_res & 7) == 0) {
                load16(res_row + n0, z);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = (c0 + j < n_cols) ? __bfloat162float(res_row[n0 + j]) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = z[j] + res_scale * v[j];
            }
          }
          if constexpr (F::bwd) {
            if (epi.colsum && !tma_out) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float sj = row_ok ? v[j] : 0.f;
                sj = warp_sum(sj);
                if (lane == 0 && c0 + j < n_cols) atomicAdd(epi.colsum + n0 + j, sj);
              }
            }
          }
          if (tma_out) {
            // stage this thread's 16 values (one row) into the swizzled group buffer
            if (c_f32) {
              uint8_t* srow = sbuf + lane * 128;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(srow + (((4 * ci + j) ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint8_t* srow = sbuf + lane * 64;
              const int swz = (lane >> 1) & 3;
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&t);
              }
              *reinterpret_cast<uint4*>(srow + (((2 * ci) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(srow + (((2 * ci + 1) ^ swz) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else if (row_ok) {
            if (c_f32) {
              float* cp = reinterpret_cast<float*>(epi.c) + (size_t)row * epi.ldc + n0;
              if (epi.atomic_out) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < n_cols) atomicAdd(cp + j, v[j]);
              } else if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(cp + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < n_cols) cp[j] = v[j];
              }
            } else {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.c) + (size_t)row * epi.ldc + n0;
              if (full) store16(cp, v);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < n_cols) cp[j] = __float2bfloat16(v[j]);
              }
            }
          }
        }
        if (tma_out) {
          sm100::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (aux_tma) {
              sm100::tma_store_2d(&tmX, abuf, n_blk * BN + cg0, row0);
              sm100::tma_store_commit();
            }
            if (epi.atomic_out) sm100::tma_reduce_add_2d(&tmC, sbuf, n_blk * BN + cg0, row0);
            else sm100::tma_store_2d(&tmC, sbuf, (F::fwd && epi.glu) ? ((n_blk * BN + cg0) >> 1) : (n_blk * BN + cg0), row0);
            sm100::tma_store_commit();
          }
          if constexpr (F::bwd) {
            if (epi.colsum) {
              // column sums straight from the staged group (32 columns)
              const int ncols_here = min(32, c_end - cg0);
              if (c_f32) {  // 128-byte rows: lane = column
                float s0 = 0.f;
                for (int rr = 0; rr < valid_rows; ++rr)
                  s0 += *reinterpret_cast<const float*>(sbuf + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
                if (lane < ncols_here) atomicAdd(epi.colsum + n_blk * BN + cg0 + lane, s0);
              } else {  // 64-byte rows: 16 four-byte slots (2 bf16 columns each); the two half-warps take alternate rows
                const int slot = lane & 15;
                float s0 = 0.f, s1 = 0.f;
                for (int rr = lane >> 4; rr < valid_rows; rr += 2) {
                  const uint32_t wv = *reinterpret_cast<const uint32_t*>(sbuf + rr * 64 + (((slot >> 2) ^ ((rr >> 1) & 3)) << 4) + ((slot & 3) << 2));
                  s0 += __uint_as_float(wv << 16);
                  s1 += __uint_as_float(wv & 0xFFFF0000u);
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                if (lane < 16) {
                  if (2 * slot < ncols_here) atomicAdd(epi.colsum + n_blk * BN + cg0 + 2 * slot, s0);
                  if (2 * slot + 1 < ncols_here) atomicAdd(epi.colsum + n_blk * BN + cg0 + 2 * slot + 1, s1);
                }
              }
            }
          }
        }
      }
      if constexpr (F::rs) {
        if (epi.rowstats && row_ok && num_k > 0) {  // (-inf, 0) when this warp had no columns in the tile
          float2* rp = reinterpret_cast<float2*>(epi.rowstats) + (size_t)row * epi.rs_parts + (n_blk * 2 + half);
          *rp = make_float2(rs_m, rs_s);
        }
      }
      // this thread's tcgen05.ld of the tile are complete: hand the accumulator buffer back to the MMA warp
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[buf]);
      if (warp == 2 && lane == 0 && acc_it < 2) TRACE(6 + 2 * acc_it);
    }
    if (tma_out && lane == 0) sm100::tma_store_wait_read<0>();  // smem must outlive the in-flight stores (completion: kernel end)
    if (warp == 2 && lane == 0) TRACE(9);
  }
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
  if (threadIdx.x == 0) TRACE(10);
}

template <int BN, bool A_MN, bool B_MN, uint32_t FEAT>
int launch_gemm(const void* A, const void* B, const GemmEpi& epi, int M, int N, int K, int lda, int ldb, int splits, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  CUtensorMap tmA, tmB;
  bool ok;
  if (!A_MN) ok = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, kBK);
  else ok = make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, 64);
  if (!B_MN) ok = ok && make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, kBK);
  else ok = ok && make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, 64);
  if (!ok) return TFASR_INVALID_VALUE;
  CUtensorMap tmC, tmR, tmX;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmR, 0, sizeof(tmR));
  memset(&tmX, 0, sizeof(tmX));
  GemmEpi e2 = epi;
  e2.tma_store = 0;
  e2.tma_side = 0;
  e2.tma_aux = 0;
#ifndef TFASR_EMU_BUILD
  e2.pdl_trigger = pdl_trigger_enabled() ? 1 : 0;
#else
  e2.pdl_trigger = 0;
#endif
#ifdef TFASR_GEMM_TRACE
  e2.trace = g_gemm_trace_host;
#endif
  if (epi.glu) {
    // GLU halves the width: 16 output columns per 32-column accumulator group
    if ((FEAT & 1u) && !epi.c_f32 && !epi.atomic_out && ((N / 2) % 8 == 0) &&
        make_tmap_2d(&tmC, epi.c, 2, (uint64_t)M, (uint64_t)(N / 2), (uint64_t)epi.ldc, 32, 16))
      e2.tma_store = 1;
  } else if (!epi.atomic_out || epi.c_f32) {
    const int eb = epi.c_f32 ? 4 : 2;
    if (make_tmap_2d(&tmC, epi.c, eb, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ldc, 32, 32) && (N % (16 / eb) == 0)) e2.tma_store = 1;
  }
  if (e2.tma_store && !epi.c_f32) {
    if ((FEAT & 1u) && epi.aux && (N % 8 == 0) && make_tmap_2d(&tmX, epi.aux, 2, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ld_aux, 32, 32)) e2.tma_aux = 1;
    // one bf16 side input of the output's shape rides a TMA load into the staging buffer (residual first, else the `mul` operand)
    if (!e2.tma_aux && !epi.glu) {
      if ((FEAT & 2u) && epi.residual && make_tmap_2d(&tmR, epi.residual, 2, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ld_res, 32, 32)) e2.tma_side = 1;
      else if ((FEAT & 4u) && epi.mul && make_tmap_2d(&tmR, epi.mul, 2, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ld_mul, 32, 32)) e2.tma_side = 2;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_MN, B_MN, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  const long long tiles = (long long)ceil_div(M, kBM) * ceil_div(N, BN) * splits;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
#ifndef TFASR_EMU_BUILD
  if (pdl_enabled()) {  // TFASR_PDL (default on): the launch may overlap its prologue with the tail of the kernel before it (sm100::grid_dependency_wait)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, A_MN, B_MN, FEAT>, tmA, tmB, tmC, tmR, tmX, e2, M, N, K, splits) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    count_launch();
    return check_launch();
  }
#endif
  gemm_bf16_kernel<BN, A_MN, B_MN, FEAT><<<grid, kGemmThreads, S::TOTAL, stream>>>(tmA, tmB, tmC, tmR, tmX, e2, M, N, K, splits);
  count_launch();
  return check_launch();
}


#define TFASR_GEMM_INSTANTIATE(BN)                                                                                                   \
  template int launch_gemm<BN, false, false, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 1u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 2u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 3u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 4u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 6u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, true, true, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);   \
  template int launch_gemm<BN, true, true, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);

// K-major operands only (MN-major B boxes are 64 columns wide: tile widths that are not a multiple of 64 are not offered there)
#define TFASR_GEMM_INSTANTIATE_KMAJOR(BN)                                                                                            \
  template int launch_gemm<BN, false, false, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 1u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 2u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 3u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 4u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 6u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  template int launch_gemm<BN, false, false, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);

}  // namespace tfasr
