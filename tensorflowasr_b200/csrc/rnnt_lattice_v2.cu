// RNN-T alpha / beta wavefronts, second generation (ops.set_lattice_v2): same recurrences and outputs as rnnt_lattice_kernel (rnnt_loss.cu;
// losses/impl/rnnt.py:108-178 forward_dp / backward_dp, :230,277 loss = -beta[0,0]), restructured around what bounds it.
//
// The lattice is T_b + U_b - 1 DEPENDENT steps (314 at T' = 250, U = 64); one step is one log-add-exp per cell of an anti-diagonal, one
// warp per 32 label positions, so its duration is the per-step instruction chain of a single warp times the step count.  The first
// generation spends most of each step on 64-bit address arithmetic for three global arrays and on global-load latency behind a 4-deep
// register ring (150 us, ~480 ns per step).  Here
//   * the utterance's blank / truth log-probability planes ([T_b, U+1] f32 each, 65 KB at the bench shape) are staged into shared memory
//     once, with every thread's loads in flight together; a step then reads them with one 32-bit index and ~30 cycles of latency,
//     prefetched a step ahead so that no load sits on the dependent chain,
//   * the direction (alpha / beta) is a template parameter: no per-step branches on it,
//   * the only 64-bit arithmetic left is the running output pointer.
// Needs 8 T (U+1) bytes of shared memory (tfasr_rnnt_lattice_smem_supported); larger lattices stay on the first-generation kernel.
// Latency-bound by construction: no roofline applies; its bytes (3 planes read or written once) are counted in the RNN-T group.
// Plain SIMT: builds twice (nvcc -> libtfasr_b200.so; g++ -DTFASR_EMU_BUILD -> tests/emu, host pointers).
#include "simt.cuh"
#include "lattice_lens.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

__device__ __forceinline__ float lattice_lae(float a, float b) {  // log(exp(a) + exp(b)); -inf when both are
  const float m = fmaxf(a, b);
  const float d = -fabsf(a - b);
  const float r = m + __logf(1.0f + __expf(d));
  return (m == -CUDART_INF_F) ? -CUDART_INF_F : r;
}

struct LatticeArgs {
  const float *blank, *truth;  // [B, T, U1] log-probabilities of the two transitions leaving (t, u)
  float *alpha, *beta;         // [B, T, U1]
  float* costs;                // [B]
  const int32_t *logit_len, *label_len;
  int T, U1;
};

// Thread i owns label position u (alpha: u = i; beta: u = U_b - i) and walks time; at step n it is at own-time tt = n - i:
//   alpha: t = tt          a[t,u]  = lae(a[t-1,u] + blank[t-1,u], a[t,u-1] + truth[t,u-1])
//   beta : t = T_b-1-tt    be[t,u] = lae(be[t+1,u] + blank[t,u],  be[t,u+1] + truth[t,u])
// `stay` is the thread's own running term, the other term arrives from thread i-1 (its value of the previous step): warp shuffle inside a
// warp, s_edge between warps.
template <bool BETA>
__device__ __forceinline__ void lattice_walk(const LatticeArgs& a, int b, int Tb, int Ub, float* sb, float* st, float* s_edge) {
  const int i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const int U1 = a.U1;
  const size_t base = (size_t)b * a.T * U1;
  const int cells = Tb * U1;
  {  // stage the two planes: all loads of a batch of 8 are in flight before the first store
    const float* __restrict__ gb = a.blank + base;
    const float* __restrict__ gt = a.truth + base;
    const int nth = blockDim.x;
    int j = i;
    for (; j + 7 * nth < cells; j += 8 * nth) {
      float vb[8], vt[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        vb[k] = gb[j + k * nth];
        vt[k] = gt[j + k * nth];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        sb[j + k * nth] = vb[k];
        st[j + k * nth] = vt[k];
      }
    }
    for (; j < cells; j += nth) {
      sb[j] = gb[j];
      st[j] = gt[j];
    }
  }
  __syncthreads();
  const bool active = i <= Ub;
  const int u = BETA ? Ub - i : i;
  const int stride = BETA ? -U1 : U1;
  int idx = active ? (BETA ? (Tb - 1) * U1 + u : u) : 0;  // cell of own-time 0
  float* op = (BETA ? a.beta : a.alpha) + base + idx;
  float bl = active ? sb[idx] : 0.f, tr = active ? st[idx] : 0.f;  // own-time 0, refilled one step ahead below
  float stay = (i == 0) ? 0.f : -CUDART_INF_F;  // alpha: virtual a[-1,0] + blank := 0; beta: virtual be[T_b, U_b] := 0
  float emit_out = -CUDART_INF_F;
  const bool multi_warp = blockDim.x > 32;
  const int nsteps = Tb + Ub;
  const int first = i, last = i + Tb - 1;  // the steps at which this thread holds a cell
  for (int n = 0; n < nsteps; ++n) {
    float emit_in = __shfl_up_sync(0xffffffffu, emit_out, 1);
    if (lane == 0) emit_in = (warp == 0) ? -CUDART_INF_F : s_edge[((n + 1) & 1) * 32 + warp - 1];
    float new_emit = -CUDART_INF_F;
    if (active && n >= first && n <= last) {
      float val;
      if (!BETA) {
        val = lattice_lae(stay, emit_in);
        stay = val + bl;      // feeds a[t+1,u]
        new_emit = val + tr;  // feeds a[t,u+1]  (tr = -inf at u == U_b)
      } else {
        val = lattice_lae(stay + bl, emit_in + tr);  // emit_in = be[t,u+1]; tr = -inf at u == U_b
        stay = val;
        new_emit = val;
      }
      *op = val;
      if (BETA && n == last && i == Ub) a.costs[b] = -val;  // loss = -beta[0,0]
      if (n < last) {
        idx += stride;
        op += stride;
        bl = sb[idx];
        tr = st[idx];
      }
    }
    emit_out = new_emit;
    if (multi_warp) {
      if (lane == 31) s_edge[(n & 1) * 32 + warp] = emit_out;
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024) rnnt_lattice_smem_kernel(LatticeArgs a) {
  TFASR_DYN_SMEM(smem_raw);
  const int b = blockIdx.x;
  int Tb, Ub;
  lattice_lens(a.logit_len, a.label_len, b, a.T, a.U1 - 1, Tb, Ub);
  if (Tb <= 0) {  // empty utterance: no alignment exists
    if (blockIdx.y == 1 && threadIdx.x == 0) a.costs[b] = CUDART_INF_F;
    return;
  }
  float* s_edge = reinterpret_cast<float*>(smem_raw);  // [2][32]
  float* sb = s_edge + 64;
  float* st = sb + (size_t)a.T * a.U1;
  if (blockIdx.y == 1) lattice_walk<true>(a, b, Tb, Ub, sb, st, s_edge);
  else lattice_walk<false>(a, b, Tb, Ub, sb, st, s_edge);
}

static size_t lattice_smem_bytes(int T, int U1) { return sizeof(float) * (64 + 2 * (size_t)T * U1); }

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_rnnt_lattice_smem_supported(int32_t T, int32_t U1) {
  return T >= 1 && U1 >= 1 && U1 <= 1024 && lattice_smem_bytes(T, U1) <= 227u * 1024u;
}

extern "C" int tfasr_rnnt_lattice_smem_f32(const float* blank, const float* truth, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                           int32_t T, int32_t U1, float* alpha, float* beta, float* costs, tfasr_stream_t stream_) {
  if (!blank || !truth || !logit_len || !label_len || !alpha || !beta || !costs || B < 0) return TFASR_INVALID_VALUE;
  if (!tfasr_rnnt_lattice_smem_supported(T, U1)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  LatticeArgs a{blank, truth, alpha, beta, costs, logit_len, label_len, T, U1};
  const int threads = 32 * (int)ceil_div(U1, 32);
  const size_t smem = lattice_smem_bytes(T, U1);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(dim3((unsigned)B, 2), dim3(threads), smem, [&] { rnnt_lattice_smem_kernel(a); });
  return TFASR_OK;
#else
  if (smem > 48 * 1024 && cudaFuncSetAttribute(rnnt_lattice_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return TFASR_EXECUTION_FAILED;
  rnnt_lattice_smem_kernel<<<dim3((unsigned)B, 2), threads, smem, static_cast<cudaStream_t>(stream_)>>>(a);
  count_launch();
  return check_launch();
#endif
}
