// Log-softmax normaliser of the joint logits from the per-slice softmax statistics the GEMM epilogue leaves behind (gemm_kernel.cuh,
// feature class 8; SURVEY.md 8f N1, first step): replaces K1 of rnnt_loss.cu — a full 2 GB read of the logits — by a pass over
// [rows, P, 2] floats (P = 8 at V = 1000) plus two gathered logits per row (blank, truth).
//   lse[row]   = M + log sum_s sum_s * exp(max_s - M),  M = max_s max_s        (= log sum_v exp(logits[row, v]), losses/impl/rnnt.py:211-214)
//   blank[row] = logits[row, 0] - lse;   truth[row] = logits[row, labels[b, u]] - lse  (u < U_b; -inf on the last column, :216-231)
// Plain SIMT: builds twice (nvcc -> libtfasr_b200.so; g++ -DTFASR_EMU_BUILD -> tests/emu, host pointers).
#include "simt.cuh"
#include "lattice_lens.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

__device__ __forceinline__ float logit_f(const float* x, int j) { return x[j]; }
__device__ __forceinline__ float logit_f(const __nv_bfloat16* x, int j) { return __bfloat162float(x[j]); }

// LT = float, or __nv_bfloat16 (the joint GEMM writes bf16 logits and takes its statistics over the rounded values: ops.bf16_logits)
template <typename LT>
__global__ void __launch_bounds__(256) rnnt_lse_combine_kernel(const LT* __restrict__ logits, int64_t ld, const float2* __restrict__ rowstats, int P,
                                                               const int32_t* __restrict__ labels, const int32_t* __restrict__ logit_len,
                                                               const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                               float* __restrict__ lse, float* __restrict__ blank, float* __restrict__ truth) {
  const int64_t nrows = (int64_t)B * T * U1;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r32 = (uint32_t)row;  // B T U1 < 2^31 (checked by the launcher): 32-bit divisions
    const int u = (int)(r32 % (uint32_t)U1);
    const int t = (int)((r32 / (uint32_t)U1) % (uint32_t)T);
    const int b = (int)(r32 / ((uint32_t)U1 * (uint32_t)T));
    int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
    if (t >= Tb || u > Ub) continue;  // rows outside the lattice are never read
    const float2* rp = rowstats + (size_t)row * P;
    float m = -CUDART_INF_F;
    for (int s = 0; s < P; ++s) m = fmaxf(m, rp[s].x);
    float acc = 0.f;
    for (int s = 0; s < P; ++s) {
      const float2 ps = rp[s];
      if (ps.y > 0.f) acc = fmaf(ps.y, __expf(ps.x - m), acc);
    }
    const float l = m + logf(acc);
    const LT* x = logits + (size_t)row * ld;
    lse[row] = l;
    blank[row] = logit_f(x, 0) - l;
    truth[row] = (u < Ub) ? logit_f(x, labels[(size_t)b * (U1 - 1) + u]) - l : -CUDART_INF_F;
  }
}

}  // namespace tfasr

using namespace tfasr;

template <typename LT>
static int lse_from_rowstats(const LT* logits, int64_t ld, const float* rowstats, int32_t parts, const int32_t* labels, const int32_t* logit_len,
                             const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V, float* lse, float* blank, float* truth,
                             tfasr_stream_t stream_) {
  if ((int64_t)B * T * U1 > 0x7fffffff) return TFASR_INVALID_VALUE;
  if (!logits || ld < V || !rowstats || parts < 1 || !logit_len || !label_len || (U1 > 1 && !labels) || B < 0 || T < 1 || U1 < 1 || V < 2 || !lse || !blank ||
      !truth || (reinterpret_cast<uintptr_t>(rowstats) & 7))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int64_t nrows = (int64_t)B * T * U1;
  int64_t g = ceil_div(nrows, 256);
  if (g > (int64_t)32 * kNumSMs) g = (int64_t)32 * kNumSMs;
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(dim3((unsigned)g), dim3(256), 0, [&] {
    rnnt_lse_combine_kernel<LT>(logits, ld, reinterpret_cast<const float2*>(rowstats), parts, labels, logit_len, label_len, B, T, U1, V, lse, blank, truth);
  });
  return TFASR_OK;
#else
  rnnt_lse_combine_kernel<LT><<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream_)>>>(logits, ld, reinterpret_cast<const float2*>(rowstats), parts,
                                                                                          labels, logit_len, label_len, B, T, U1, V, lse, blank, truth);
  count_launch();
  return check_launch();
#endif
}

extern "C" int tfasr_rnnt_lse_from_rowstats_f32(const float* logits, const float* rowstats, int32_t parts, const int32_t* labels,
                                                const int32_t* logit_len, const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V,
                                                float* lse, float* blank, float* truth, tfasr_stream_t stream) {
  return lse_from_rowstats<float>(logits, V, rowstats, parts, labels, logit_len, label_len, B, T, U1, V, lse, blank, truth, stream);
}

extern "C" int tfasr_rnnt_lse_from_rowstats_bf16(const void* logits, int64_t ld_logits, const float* rowstats, int32_t parts, const int32_t* labels,
                                                 const int32_t* logit_len, const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V,
                                                 float* lse, float* blank, float* truth, tfasr_stream_t stream) {
  return lse_from_rowstats<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(logits), ld_logits, rowstats, parts, labels, logit_len, label_len, B, T, U1, V, lse,
                                          blank, truth, stream);
}
