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
#include "common.cuh"
#include "sm100.cuh"
#include "dropout.cuh"

#include <string.h>

#include <mutex>

namespace tfasr {

// ------------------------------------------------------------------------------------------------ host: tensor maps
PFN_tmapEncodeTiled get_tmap_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

bool make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                       uint32_t box_cols) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  uint32_t box_cols) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15)) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

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
};

// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5 : one MUFU op (tanh.approx) instead of ex2 + rcp
__device__ __forceinline__ float sigmoidf_(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kBM = 128;
constexpr int kBK = 64;

// Persistent, warp-specialised kernel: one CTA per SM walks output tiles (n fastest, so the N-tiles of one M-tile run
// back to back and A is read from HBM once); the smem ring and the two TMEM accumulator buffers run continuously across
// tiles, so the epilogue of tile i overlaps the TMA + MMA of tile i+1.
template <int BN>
struct GemmSmem {
  static constexpr int STAGES = (BN >= 192) ? 4 : 6;
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_OUT_BYTES = 8 /*epilogue warps*/ * 32 /*rows*/ * 128 /*bytes*/;
  static constexpr int TOTAL = STAGES * (A_BYTES + B_BYTES) + STAGE_OUT_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);  // two accumulators
  static_assert(TOTAL <= 227 * 1024, "smem budget");
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const GemmEpi epi, int M, int N, int K, int splits) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = sm100::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S::STAGES * S::A_BYTES;
  uint8_t* sOut = smem + S::STAGES * (S::A_BYTES + S::B_BYTES);  // [4 warps][2][32 rows x 128 B], 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + S::STAGE_OUT_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int total_k = (K + kBK - 1) / kBK;
  const int per_split = (total_k + splits - 1) / splits;
  const int num_tiles = m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    sm100::prefetch_tensormap(&tmA);
    sm100::prefetch_tensormap(&tmB);
    if (epi.tma_store) sm100::prefetch_tensormap(&tmC);
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
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const float alpha = epi.alpha;
    const int act = epi.act, mul_mode = epi.mul_mode;
    const bool has_drop = epi.drop_p > 0.f;
    const float drop_p = epi.drop_p;
    const float keep_scale = has_drop ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t drop_thr = dropout_threshold(drop_p);
    const unsigned long long seed = (has_drop && epi.seed_ptr) ? (*epi.seed_ptr + epi.salt) : epi.salt;
    const float res_scale = epi.res_scale;
    const bool vec16 = ((epi.ldc & 7) == 0) || (epi.c_f32 && (epi.ldc & 3) == 0);
    const bool bias_vec = epi.bias && ((reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0);

    auto load16 = [](const __nv_bfloat16* p, float* f) {  // 16 bf16 (32 B, 16-byte aligned) -> f32
      const uint4 u0 = *reinterpret_cast<const uint4*>(p);
      const uint4 u1 = *reinterpret_cast<const uint4*>(p + 8);
      const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[2 * j] = __uint_as_float(w[j] << 16);
        f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
    };
    auto store16 = [](__nv_bfloat16* p, const float* f) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(p + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };


    // C staging: per warp two 32-row x 128-byte buffers in the 128B-swizzled layout the store tensor map expects
    uint8_t* sbuf = sOut + (warp - 2) * (32 * 128);  // this warp's staging buffer (1024-byte aligned)
    const int half = (warp - 2) >> 2;               // which column range of the tile this warp drains
    const bool tma_out = epi.tma_store != 0;
    const int group_cols = epi.c_f32 ? 32 : 64;  // columns per 128-byte staging row
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      int m_blk, n_blk, k_begin, num_k;
      decode(tile, m_blk, n_blk, k_begin, num_k);
      const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
      const int row = m_blk * kBM + q * 32 + lane;
      const bool row_ok = row < M;
      const __nv_bfloat16* mul_row = epi.mul ? reinterpret_cast<const __nv_bfloat16*>(epi.mul) + (size_t)row * epi.ld_mul : nullptr;
      const __nv_bfloat16* res_row = epi.residual ? reinterpret_cast<const __nv_bfloat16*>(epi.residual) + (size_t)row * epi.ld_res : nullptr;
      __nv_bfloat16* aux_row = epi.aux ? reinterpret_cast<__nv_bfloat16*>(epi.aux) + (size_t)row * epi.ld_aux : nullptr;
      const uint32_t tmem_row = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      const int n_cols = min(BN, N - n_blk * BN);  // valid columns of this tile
      sm100::mbar_wait(&tmem_full[buf], aph);
      sm100::tc_fence_after();
      // column split between the two warps of a lane quarter, in whole staging groups (64 bf16 / 32 f32 columns)
      const int groups = (BN + group_cols - 1) / group_cols;
      const int g_split = (groups + 1) / 2;
      const int c_begin = half == 0 ? 0 : g_split * group_cols;
      const int c_end = min(n_cols, half == 0 ? g_split * group_cols : BN);
      const int valid_rows = max(0, min(32, M - (m_blk * kBM + q * 32)));
      uint32_t r[16];
      if (num_k > 0 && c_begin < c_end) sm100::tmem_ld_32x32b_x16(tmem_row + (uint32_t)c_begin, r);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end && num_k > 0; c0 += 16) {
      const int n0 = n_blk * BN + c0;
      sm100::tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      if (c0 + 16 < c_end) sm100::tmem_ld_32x32b_x16(tmem_row + (uint32_t)(c0 + 16), r);  // prefetch the next chunk
      const bool full = (c0 + 16 <= n_cols) && vec16;
      if (alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= alpha;
      }
      if (epi.bias) {
        if (bias_vec && c0 + 16 <= n_cols) {  // same address in every lane: one broadcast transaction per load
          const float4* gb = reinterpret_cast<const float4*>(epi.bias + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(gb + j);
            v[4 * j] += b4.x;
            v[4 * j + 1] += b4.y;
            v[4 * j + 2] += b4.z;
            v[4 * j + 3] += b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < n_cols) v[j] += __ldg(epi.bias + n0 + j);
        }
      }
      if (aux_row && row_ok) {
        if (full && (epi.ld_aux & 7) == 0) store16(aux_row + n0, v);
        else
        {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < n_cols) aux_row[n0 + j] = __float2bfloat16(v[j]);
        }
      }
      if (epi.glu) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[2 * j] * sigmoidf_(v[2 * j + 1]);
        if (row_ok) {
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
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] * sigmoidf_(v[j]);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (act == 3) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tanhf(v[j]);
      }
      if (mul_row && row_ok) {
        float z[16];
        if (full && (epi.ld_mul & 7) == 0) {
          load16(mul_row + n0, z);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = (c0 + j < n_cols) ? __bfloat162float(mul_row[n0 + j]) : 0.f;
        }
        if (mul_mode == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float sg = sigmoidf_(z[j]);
            v[j] *= sg * (1.0f + z[j] * (1.0f - sg));  // d swish / dz
          }
        } else if (mul_mode == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= 1.0f - z[j] * z[j];  // d tanh / dz from the tanh value
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= z[j];
        }
      }
      if (has_drop) {
        const unsigned long long base = (unsigned long long)row * (unsigned long long)N + (unsigned long long)n0;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const uint32_t hsh = dropout_hash(seed, base + j);  // base is even (n0 % 16 == 0, N even)
          v[j] = ((hsh & 0xFFFFu) >= drop_thr) ? v[j] * keep_scale : 0.f;
          v[j + 1] = ((hsh >> 16) >= drop_thr) ? v[j + 1] * keep_scale : 0.f;
        }
      }
      if (res_row && row_ok) {
        float z[16];
        if (full && (epi.ld_res & 7) == 0) {
          load16(res_row + n0, z);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = (c0 + j < n_cols) ? __bfloat162float(res_row[n0 + j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = z[j] + res_scale * v[j];
      }
      if (epi.colsum && !tma_out) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float sj = row_ok ? v[j] : 0.f;
          sj = warp_sum(sj);
          if (lane == 0 && c0 + j < n_cols) atomicAdd(epi.colsum + n0 + j, sj);
        }
      }
      if (tma_out) {
        // stage this thread's 16 values (one row) into the swizzled buffer; a full 128-byte group leaves with one TMA store
        const int col_in_group = (c0 - c_begin) % group_cols;
        if (col_in_group == 0) {
          // buffer reuse: the previous store must have finished reading it
          if (lane == 0) sm100::tma_store_wait_read<0>();
          __syncwarp();
        }
        uint8_t* srow = sbuf + lane * 128;
        if (epi.c_f32) {
          const int cg = col_in_group >> 2;  // 16-byte chunk index of the first of 4 chunks
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(srow + (((cg + j) ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          const int cg = col_in_group >> 3;
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(srow + ((cg ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(srow + (((cg + 1) ^ (lane & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        const bool group_done = (col_in_group + 16 == group_cols) || (c0 + 16 >= c_end);
        if (group_done) {
          sm100::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            sm100::tma_store_2d(&tmC, sbuf, n_blk * BN + (c0 - col_in_group), m_blk * kBM + q * 32);
            sm100::tma_store_commit();
          }
          if (epi.colsum) {
            // column sums straight from the staged tile: each lane owns one 4-byte column slot (2 bf16 / 1 f32 columns)
            const int gcol0 = n_blk * BN + (c0 - col_in_group);
            const int ncols_here = min(group_cols, c_end - (c0 - col_in_group));
            float s0 = 0.f, s1 = 0.f;
            for (int rr = 0; rr < valid_rows; ++rr) {
              const uint32_t wv = *reinterpret_cast<const uint32_t*>(sbuf + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
              if (epi.c_f32) {
                s0 += __uint_as_float(wv);
              } else {
                s0 += __uint_as_float(wv << 16);
                s1 += __uint_as_float(wv & 0xFFFF0000u);
              }
            }
            if (epi.c_f32) {
              if (lane < ncols_here) atomicAdd(epi.colsum + gcol0 + lane, s0);
            } else {
              if (2 * lane < ncols_here) atomicAdd(epi.colsum + gcol0 + 2 * lane, s0);
              if (2 * lane + 1 < ncols_here) atomicAdd(epi.colsum + gcol0 + 2 * lane + 1, s1);
            }
          }
        }
      } else if (row_ok) {
        if (epi.c_f32) {
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
      // this thread's tcgen05.ld of the tile are complete: hand the accumulator buffer back to the MMA warp
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[buf]);
    }
    if (tma_out && lane == 0) sm100::tma_store_wait_all();  // smem must outlive the in-flight stores
  }
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const void* A, const void* B, const GemmEpi& epi, int M, int N, int K, int lda, int ldb, int splits, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  CUtensorMap tmA, tmB;
  bool ok;
  if (!A_MN) ok = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, kBK);
  else ok = make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, 64);
  if (!B_MN) ok = ok && make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, kBK);
  else ok = ok && make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, 64);
  if (!ok) return TFASR_INVALID_VALUE;
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  GemmEpi e2 = epi;
  e2.tma_store = 0;
  if (!epi.glu && !epi.atomic_out) {
    const int eb = epi.c_f32 ? 4 : 2;
    if (make_tmap_2d(&tmC, epi.c, eb, (uint64_t)M, (uint64_t)N, (uint64_t)epi.ldc, 32, 128 / eb) && (N % (16 / eb) == 0)) e2.tma_store = 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  const long long tiles = (long long)ceil_div(M, kBM) * ceil_div(N, BN) * splits;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  gemm_bf16_kernel<BN, A_MN, B_MN><<<grid, kGemmThreads, S::TOTAL, stream>>>(tmA, tmB, tmC, e2, M, N, K, splits);
  count_launch();
  return check_launch();
}

static int choose_bn(int N) {
  const int cands[4] = {64, 128, 192, 256};
  int best = 256, best_cost = 1 << 30;
  for (int c : cands) {
    const int cost = (int)ceil_div(N, c) * c;
    if (cost < best_cost || (cost == best_cost && c > best)) {
      best = c;
      best_cost = cost;
    }
  }
  return best;
}

template <bool A_MN, bool B_MN>
static int dispatch_bn(const void* A, const void* B, const GemmEpi& epi, int M, int N, int K, int lda, int ldb, int splits, cudaStream_t stream) {
  switch (choose_bn(N)) {
    case 64: return launch_gemm<64, A_MN, B_MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    case 128: return launch_gemm<128, A_MN, B_MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    case 192: return launch_gemm<192, A_MN, B_MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    default: return launch_gemm<256, A_MN, B_MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_gemm_bf16(const void* A, const void* B, void* C, int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb, int32_t ldc,
                               int32_t a_mn_major, int32_t b_mn_major, const tfasr_gemm_epilogue* e, tfasr_stream_t stream_) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1) return TFASR_INVALID_VALUE;
  GemmEpi epi;
  memset(&epi, 0, sizeof(epi));
  epi.c = C;
  epi.ldc = ldc;
  epi.alpha = 1.0f;
  epi.res_scale = 1.0f;
  if (e) {
    epi.bias = e->bias;
    epi.aux = e->aux;
    epi.ld_aux = e->ld_aux;
    epi.residual = e->residual;
    epi.ld_res = e->ld_residual;
    epi.mul = e->mul;
    epi.ld_mul = e->ld_mul;
    epi.colsum = e->colsum;
    epi.c_f32 = e->c_f32;
    epi.act = e->activation;
    epi.glu = e->glu;
    epi.mul_mode = e->mul_mode;
    epi.res_scale = e->residual_scale;
    epi.alpha = e->alpha;
    epi.drop_p = e->dropout_p;
    epi.seed_ptr = e->dropout_seed_ptr;
    epi.salt = e->dropout_salt;
    epi.atomic_out = e->atomic_out;
  }
  if (epi.glu && ((N & 1) || epi.c_f32)) return TFASR_INVALID_VALUE;
  if (epi.atomic_out && !epi.c_f32) return TFASR_INVALID_VALUE;
  int splits = (e && e->k_splits > 1) ? e->k_splits : 1;
  if (splits > 1 && !epi.atomic_out) return TFASR_INVALID_VALUE;
  if (splits > (K + kBK - 1) / kBK) splits = (K + kBK - 1) / kBK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a_mn_major && !b_mn_major) return dispatch_bn<false, false>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  if (a_mn_major && b_mn_major) return dispatch_bn<true, true>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  return TFASR_INVALID_VALUE;
}
