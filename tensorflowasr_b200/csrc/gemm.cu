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
#include "gemm_kernel.cuh"

#include <mutex>
#include <stdlib.h>

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
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols * elem_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_NONE  // GLU output groups: 32-byte rows
                                               : (box_cols * elem_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


// kernels are instantiated per tile width in gemm_inst_*.cu (parallel compilation)

#define TFASR_GEMM_EXTERN(BN)                                                                                                               \
  extern template int launch_gemm<BN, false, false, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 1u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 2u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 3u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 4u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 6u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, true, true, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);   \
  extern template int launch_gemm<BN, true, true, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);
#define TFASR_GEMM_EXTERN_KMAJOR(BN)                                                                                                        \
  extern template int launch_gemm<BN, false, false, 0u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 1u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 2u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 3u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 4u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 6u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t); \
  extern template int launch_gemm<BN, false, false, 7u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);
TFASR_GEMM_EXTERN_KMAJOR(160)
TFASR_GEMM_EXTERN(64)
TFASR_GEMM_EXTERN(128)
TFASR_GEMM_EXTERN(192)
TFASR_GEMM_EXTERN(256)
extern template int launch_gemm<256, false, false, 8u>(const void*, const void*, const GemmEpi&, int, int, int, int, int, int, cudaStream_t);

#ifdef TFASR_GEMM_TRACE
unsigned long long* g_gemm_trace_host = nullptr;
#endif

// tile width with the least padded work; ties go to the wider tile (A is re-read once per column tile)
static int choose_bn(int N, bool mn_major) {
  const int cands[5] = {64, 128, 160, 192, 256};
  int best = 256, best_cost = 1 << 30;
  for (int c : cands) {
    if (mn_major && (c % 64)) continue;
    const int cost = (int)ceil_div(N, c) * c;
    if (cost < best_cost || (cost == best_cost && c > best)) {
      best = c;
      best_cost = cost;
    }
  }
  return best;
}

// smallest compiled feature class that covers what this call asks for (gemm_kernel.cuh: EpiFeat)
static uint32_t feature_class(const GemmEpi& e, bool mn_major) {
  uint32_t need = 0;
  if (e.act || e.aux || e.glu) need |= 1u;
  if (e.drop_p > 0.f || e.residual) need |= 2u;
  if (e.mul || e.colsum) need |= 4u;
  if (mn_major) return need ? 7u : 0u;
  return need == 5u ? 7u : need;
}

template <int BN, bool MN>
static int dispatch_feat(const void* A, const void* B, const GemmEpi& epi, int M, int N, int K, int lda, int ldb, int splits, cudaStream_t stream) {
  const uint32_t f = feature_class(epi, MN);
  if constexpr (MN) {
    if (f == 0u) return launch_gemm<BN, true, true, 0u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    return launch_gemm<BN, true, true, 7u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  } else {
    switch (f) {
      case 0u: return launch_gemm<BN, false, false, 0u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      case 1u: return launch_gemm<BN, false, false, 1u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      case 2u: return launch_gemm<BN, false, false, 2u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      case 3u: return launch_gemm<BN, false, false, 3u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      case 6u: return launch_gemm<BN, false, false, 6u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      case 4u: return launch_gemm<BN, false, false, 4u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
      default: return launch_gemm<BN, false, false, 7u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    }
  }
}

template <bool MN>
static int dispatch_bn(const void* A, const void* B, const GemmEpi& epi, int M, int N, int K, int lda, int ldb, int splits, cudaStream_t stream) {
  if (epi.rowstats) {
    // row statistics ride on the lean epilogue of the 256-column tile only (K-major operands, f32 or bf16 output, no split-K)
    if constexpr (MN) return TFASR_INVALID_VALUE;
    else {
      if (feature_class(epi, false) != 0u || splits != 1 || epi.atomic_out) return TFASR_INVALID_VALUE;
      return launch_gemm<256, false, false, 8u>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    }
  }
  const int bn = choose_bn(N, MN);
  if constexpr (!MN) {
    if (bn == 160) return dispatch_feat<160, false>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  }
  switch (bn) {
    case 64: return dispatch_feat<64, MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    case 128: return dispatch_feat<128, MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    case 192: return dispatch_feat<192, MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
    default: return dispatch_feat<256, MN>(A, B, epi, M, N, K, lda, ldb, splits, stream);
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
    epi.rowstats = e->rowstats;
    epi.rs_parts = 2 * ((N + 255) / 256);
  }
  if (epi.glu && ((N & 1) || epi.c_f32)) return TFASR_INVALID_VALUE;
  if (epi.atomic_out && !epi.c_f32) return TFASR_INVALID_VALUE;
  int splits = (e && e->k_splits > 1) ? e->k_splits : 1;
  if (splits > 1 && !epi.atomic_out) return TFASR_INVALID_VALUE;
  if (splits > (K + kBK - 1) / kBK) splits = (K + kBK - 1) / kBK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a_mn_major && !b_mn_major) return dispatch_bn<false>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  if (a_mn_major && b_mn_major) return dispatch_bn<true>(A, B, epi, M, N, K, lda, ldb, splits, stream);
  return TFASR_INVALID_VALUE;
}

#ifdef TFASR_GEMM_TRACE
extern "C" int tfasr_debug_set_gemm_trace(void* p) {
  tfasr::g_gemm_trace_host = static_cast<unsigned long long*>(p);
  return 0;
}
#endif
