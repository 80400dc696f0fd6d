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
};

__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

constexpr int kGemmThreads = 192;
constexpr int kBM = 128;
constexpr int kBK = 64;

template <int BN>
struct GemmSmem {
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : 6);
  static_assert(STAGES * (kBM * kBK * 2 + BN * kBK * 2) + 1280 <= 227 * 1024, "smem budget");
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int TOTAL = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

// A_MN / B_MN: operand is MN-major in global memory (the contraction index is the ROW index of the source matrix):
//   A_MN: A source is [K, M] row-major (e.g. X^T . dY with X stored [rows=K_contraction.., ...]) — box [kBK rows x 64 cols] x (128/64)
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpi epi, int M, int N, int K) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = sm100::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S::STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * (S::A_BYTES + S::B_BYTES));
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int total_k = (K + kBK - 1) / kBK;
  const int per_split = (total_k + (int)gridDim.z - 1) / (int)gridDim.z;
  const int k_begin = (int)blockIdx.z * per_split;
  const int num_k = max(0, min(total_k, k_begin + per_split) - k_begin);

  if (warp == 0 && lane == 0) {
    sm100::prefetch_tensormap(&tmA);
    sm100::prefetch_tensormap(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      sm100::mbar_init(&full_bar[s], 1);
      sm100::mbar_init(&empty_bar[s], 1);
    }
    sm100::mbar_init(tmem_full, 1);
    sm100::fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  if (warp == 1) sm100::tmem_alloc(tmem_slot, kTmemCols);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        sm100::mbar_wait(&empty_bar[s], ph ^ 1);
        sm100::mbar_arrive_expect_tx(&full_bar[s], S::A_BYTES + S::B_BYTES);
        uint8_t* a = sA + s * S::A_BYTES;
        uint8_t* b = sB + s * S::B_BYTES;
        if constexpr (!A_MN) {
          sm100::tma_load_2d(a, &tmA, &full_bar[s], (k_begin + kb) * kBK, m_blk * kBM);  // [128 rows(m) x 64 k]
        } else {
          // source [K, M]: two boxes of [64 k-rows x 64 m-cols]; each box = 64 rows of 128 B
          sm100::tma_load_2d(a, &tmA, &full_bar[s], m_blk * kBM, (k_begin + kb) * kBK);
          sm100::tma_load_2d(a + kBK * 128, &tmA, &full_bar[s], m_blk * kBM + 64, (k_begin + kb) * kBK);
        }
        if constexpr (!B_MN) {
          sm100::tma_load_2d(b, &tmB, &full_bar[s], (k_begin + kb) * kBK, n_blk * BN);  // [BN rows(n) x 64 k]
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) sm100::tma_load_2d(b + j * kBK * 128, &tmB, &full_bar[s], n_blk * BN + j * 64, (k_begin + kb) * kBK);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = sm100::make_idesc_bf16_f32(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        sm100::mbar_wait(&full_bar[s], ph);
        sm100::tc_fence_after();
        const uint32_t a0 = sm100::smem_u32(sA + s * S::A_BYTES);
        const uint32_t b0 = sm100::smem_u32(sB + s * S::B_BYTES);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // K-major: 16 k = 32 B inside the 128 B swizzle row.  MN-major: 16 k-rows of 128 B = 2048 B.
          const uint64_t ad = A_MN ? sm100::make_smem_desc(a0 + k * 2048, kBK * 128, 1024) : sm100::make_smem_desc(a0 + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? sm100::make_smem_desc(b0 + k * 2048, kBK * 128, 1024) : sm100::make_smem_desc(b0 + k * 32, 16, 1024);
          sm100::umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        sm100::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      sm100::umma_commit(tmem_full);  // accumulator complete (fires immediately when no MMA was issued)
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m_blk * kBM + q * 32 + lane;
    sm100::mbar_wait(tmem_full, 0);
    sm100::tc_fence_after();
    const bool row_ok = row < M;
    const float keep_scale = epi.drop_p > 0.f ? 1.0f / (1.0f - epi.drop_p) : 1.0f;
    const unsigned long long seed = (epi.drop_p > 0.f && epi.seed_ptr) ? (*epi.seed_ptr + epi.salt) : epi.salt;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int n0 = n_blk * BN + c0;
      if (n0 >= N || num_k == 0) break;  // warp-uniform (num_k == 0: empty split, nothing to add)
      uint32_t r[16];
      sm100::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      sm100::tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + j;
        float x = __uint_as_float(r[j]) * epi.alpha;
        if (epi.bias && n < N) x += __ldg(epi.bias + n);
        v[j] = x;
      }
      if (epi.aux && row_ok) {  // pre-activation copy
        __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(epi.aux) + (size_t)row * epi.ld_aux + n0;
        if (n0 + 16 <= N) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(ap) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(ap + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
          for (int j = 0; j < 16 && n0 + j < N; ++j) ap[j] = __float2bfloat16(v[j]);
        }
      }
      if (epi.glu) {
        // interleaved pairs: out[n/2] = a * sigmoid(b)
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[2 * j] * sigmoidf_(v[2 * j + 1]);
        if (row_ok) {
          const int on0 = n0 >> 1;
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.c) + (size_t)row * epi.ldc + on0;
          if (n0 + 16 <= N) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 t = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(cp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          } else {
            for (int j = 0; j < 8 && n0 + 2 * j + 1 < N; ++j) cp[j] = __float2bfloat16(o[j]);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = v[j];
        if (epi.act == 1) x = x * sigmoidf_(x);
        else if (epi.act == 2) x = fmaxf(x, 0.f);
        else if (epi.act == 3) x = tanhf(x);
        v[j] = x;
      }
      if (epi.mul && row_ok) {
        const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(epi.mul) + (size_t)row * epi.ld_mul + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (n0 + j < N) {
            const float z = __bfloat162float(mp[j]);
            float f;
            if (epi.mul_mode == 1) {
              const float sg = sigmoidf_(z);
              f = sg * (1.0f + z * (1.0f - sg));  // d swish / dz
            } else if (epi.mul_mode == 2) {
              f = 1.0f - z * z;  // d tanh / dz given tanh value
            } else {
              f = z;
            }
            v[j] *= f;
          }
        }
      }
      if (epi.drop_p > 0.f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float u = hash_uniform(seed, (unsigned long long)row * (unsigned long long)N + (unsigned long long)(n0 + j));
          v[j] = (u >= epi.drop_p) ? v[j] * keep_scale : 0.f;
        }
      }
      if (epi.residual && row_ok) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(epi.residual) + (size_t)row * epi.ld_res + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + j < N) v[j] = __bfloat162float(rp[j]) + epi.res_scale * v[j];
      }
      if (epi.colsum) {
        // column sums over this warp's 32 rows, one atomic per column per warp
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float s = row_ok ? v[j] : 0.f;
          s = warp_sum(s);
          if (lane == 0 && n0 + j < N) atomicAdd(epi.colsum + n0 + j, s);
        }
      }
      if (row_ok) {
        if (epi.c_f32) {
          float* cp = reinterpret_cast<float*>(epi.c) + (size_t)row * epi.ldc + n0;
          if (epi.atomic_out) {
            for (int j = 0; j < 16 && n0 + j < N; ++j) atomicAdd(cp + j, v[j]);
          } else if (n0 + 16 <= N && ((epi.ldc & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(cp + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            for (int j = 0; j < 16 && n0 + j < N; ++j) cp[j] = v[j];
          }
        } else {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(epi.c) + (size_t)row * epi.ldc + n0;
          if (n0 + 16 <= N && ((epi.ldc & 7) == 0)) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(cp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(cp + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            for (int j = 0; j < 16 && n0 + j < N; ++j) cp[j] = __float2bfloat16(v[j]);
          }
        }
      }
    }
    sm100::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, kTmemCols);
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
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(M, kBM), (unsigned)ceil_div(N, BN), (unsigned)splits);
  gemm_bf16_kernel<BN, A_MN, B_MN><<<grid, kGemmThreads, S::TOTAL, stream>>>(tmA, tmB, epi, M, N, K);
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
