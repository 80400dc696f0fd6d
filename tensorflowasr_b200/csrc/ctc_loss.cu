// CTC loss + gradient for sm_100a.
//
// Replaces tf.nn.ctc_loss as called by tensorflow_asr/losses/ctc_loss.py:57-66 (dense labels, batch-major logits,
// blank_index = 0) — and equivalently the in-tree TPU restatement tensorflow_asr/losses/impl/ctc_tpu.py:1295-1314.
//
//   K1 ctc_lse_kernel      one warp per (b,t) row: log-sum-exp over V.
//   K2 ctc_lattice_kernel  one CTA per (utterance, direction), one thread per extended-label state s in [0, 2L];
//                          time loop with double-buffered SMEM rows (one barrier per frame), log-prob gathers prefetched.
//   K3 ctc_grad_kernel     one warp per (b,t) row: grad = softmax - occupancy, occupancy scattered per label in SMEM.
//
// Roofline: HBM-bound, algorithmic bytes = 8*B*T*V (read logits + write grads).
#include "common.cuh"
#include "lattice_lens.cuh"

namespace tfasr {

struct CtcWs {
  float *lse, *alpha, *beta;  // lse [B,T]; alpha/beta [B,T,S] with S = 2L+1
};

static size_t ctc_ws_lse_bytes(int B, int T) { return align_up(sizeof(float) * (size_t)B * T, 256); }
static size_t ctc_ws_plane_bytes(int B, int T, int L) { return align_up(sizeof(float) * (size_t)B * T * (2 * L + 1), 256); }

static CtcWs ctc_ws_carve(void* ws, int B, int T, int L) {
  char* p = static_cast<char*>(ws);
  CtcWs w;
  w.lse = reinterpret_cast<float*>(p);
  w.alpha = reinterpret_cast<float*>(p + ctc_ws_lse_bytes(B, T));
  w.beta = reinterpret_cast<float*>(p + ctc_ws_lse_bytes(B, T) + ctc_ws_plane_bytes(B, T, L));
  return w;
}

constexpr int kCtcRowWarps = 4;

__global__ void __launch_bounds__(kCtcRowWarps * 32) ctc_lse_kernel(const float* __restrict__ logits, const int32_t* __restrict__ logit_len,
                                                                    int B, int T, int V, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kCtcRowWarps + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const int t = (int)(row % T), b = (int)(row / T);
  if (t >= min(logit_len[b], T)) return;  // a length beyond the tensor: the utterance is infeasible (lattice_lens.cuh), rows stay in bounds
  const float* x = logits + row * V;
  float m = -CUDART_INF_F;
  for (int j = lane; j < V; j += 32) m = fmaxf(m, x[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < V; j += 32) s += __expf(x[j] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + logf(s);
}

__device__ __forceinline__ float lae3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// grid (B, 2); blockDim = 32 * ceil(S / 32), S = 2L + 1 <= 1024.
__global__ void __launch_bounds__(1024) ctc_lattice_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                           const int32_t* __restrict__ logit_len, const int32_t* __restrict__ label_len,
                                                           int T, int L, int V, int blank, CtcWs ws, float* __restrict__ costs) {
  extern __shared__ float s_row[];  // [2][S_pad + 4], data starts at offset 2 (two -inf guards on each side)
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int s = threadIdx.x;
  const int Smax = 2 * L + 1;
  int Tb, Lb;
  lattice_lens(logit_len, label_len, b, T, L, Tb, Lb);
  const int S = 2 * Lb + 1;
  const int stride = blockDim.x + 4;
  float* buf0 = s_row + 2;
  float* buf1 = s_row + stride + 2;
  const bool active = s < S;
  int lab = blank;
  bool skip = false;  // alpha: may come from s-2 ; beta: may go to s+2
  if (active && (s & 1)) {
    lab = labels[(int64_t)b * L + (s >> 1)];
    if (!is_beta) {
      if (s >= 3) skip = lab != labels[(int64_t)b * L + (s >> 1) - 1];
    } else {
      if (s + 2 < S) skip = lab != labels[(int64_t)b * L + (s >> 1) + 1];
    }
  }
  // guards
  if (s < 2) {
    s_row[s] = -CUDART_INF_F;
    s_row[stride + s] = -CUDART_INF_F;
  }
  for (int j = s; j < stride - 2; j += blockDim.x) {
    buf0[j] = -CUDART_INF_F;
    buf1[j] = -CUDART_INF_F;
  }
  __syncthreads();
  if (Tb <= 0) {
    if (s == 0 && !is_beta) costs[b] = CUDART_INF_F;
    return;
  }
  const float* __restrict__ x = logits + (size_t)b * T * V;
  const float* __restrict__ lse = ws.lse + (size_t)b * T;
  float* __restrict__ out = (is_beta ? ws.beta : ws.alpha) + (size_t)b * T * Smax;
  const int t_first = is_beta ? Tb - 1 : 0;
  const int dt = is_beta ? -1 : 1;
  float lp_next = active ? x[(size_t)t_first * V + lab] - lse[t_first] : 0.f;
  float* cur = buf0;
  float* nxt = buf1;
  for (int k = 0; k < Tb; ++k) {
    const int t = t_first + k * dt;
    const float lp = lp_next;
    if (k + 1 < Tb && active) lp_next = x[(size_t)(t + dt) * V + lab] - lse[t + dt];  // prefetch
    float val = -CUDART_INF_F;
    if (active) {
      if (k == 0) {
        if (!is_beta) val = (s <= 1) ? lp : -CUDART_INF_F;
        else val = (s >= S - 2) ? lp : -CUDART_INF_F;
      } else {
        const float a0 = cur[s];
        const float a1 = is_beta ? cur[s + 1] : cur[s - 1];
        const float a2 = skip ? (is_beta ? cur[s + 2] : cur[s - 2]) : -CUDART_INF_F;
        val = lae3(a0, a1, a2) + lp;
      }
      out[(size_t)t * Smax + s] = val;
    }
    nxt[s] = active ? val : -CUDART_INF_F;
    __syncthreads();
    float* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (!is_beta && s == 0) {
    const float a_last = cur[S - 1];
    const float a_prev = (S > 1) ? cur[S - 2] : -CUDART_INF_F;
    const float logz = log_add_exp(a_last, a_prev);
    costs[b] = -logz;  // +inf when infeasible
  }
}

__global__ void __launch_bounds__(kCtcRowWarps * 32) ctc_grad_kernel(const float* logits, const int32_t* __restrict__ labels,
                                                                     const int32_t* __restrict__ logit_len,
                                                                     const int32_t* __restrict__ label_len, int B, int T, int L, int V,
                                                                     int blank, CtcWs ws, const float* __restrict__ costs,
                                                                     const float* __restrict__ grad_scale, float* grads) {
  extern __shared__ float s_occ[];  // [kCtcRowWarps][V]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kCtcRowWarps + w;
  if (row >= (int64_t)B * T) return;
  const int t = (int)(row % T), b = (int)(row / T);
  int Tb, Lb;
  lattice_lens(logit_len, label_len, b, T, L, Tb, Lb);
  const float* x = logits + row * V;
  float* g = grads + row * V;
  const float cost = costs[b];
  if (t >= Tb || !(cost < CUDART_INF_F)) {  // beyond the utterance, or infeasible alignment: zero gradient
    for (int j = lane; j < V; j += 32) g[j] = 0.f;
    return;
  }
  float* occ = s_occ + (size_t)w * V;
  for (int j = lane; j < V; j += 32) occ[j] = 0.f;
  __syncwarp();
  const int S = 2 * Lb + 1, Smax = 2 * L + 1;
  const float lse = ws.lse[row];
  const float logz = -cost;
  const float* al = ws.alpha + (size_t)row * Smax;
  const float* be = ws.beta + (size_t)row * Smax;
  for (int s = lane; s < S; s += 32) {
    const int lab = (s & 1) ? labels[(int64_t)b * L + (s >> 1)] : blank;
    const float ab = al[s] + be[s];
    if (ab > -CUDART_INF_F) atomicAdd(&occ[lab], expf(ab - (x[lab] - lse) - logz));
  }
  __syncwarp();
  const float sc = grad_scale ? grad_scale[b] : 1.0f;
  for (int j = lane; j < V; j += 32) g[j] = sc * (__expf(x[j] - lse) - occ[j]);
}

}  // namespace tfasr

using namespace tfasr;

extern "C" size_t tfasr_ctc_workspace_bytes(int32_t B, int32_t T, int32_t L, int32_t V) {
  (void)V;
  if (B < 0 || T < 1 || L < 0) return 0;
  return ctc_ws_lse_bytes(B, T) + 2 * ctc_ws_plane_bytes(B, T, L) + 256;
}

extern "C" int tfasr_ctc_loss_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                                  int32_t B, int32_t T, int32_t L, int32_t V, int32_t blank, float* costs, float* grads,
                                  const float* grad_scale, void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  if (blank != 0) return TFASR_INVALID_VALUE;  // losses/base_loss.py:24
  if (B < 0 || T < 1 || L < 0 || V < 2 || 2 * L + 1 > 1024) return TFASR_INVALID_VALUE;
  if (!logits || !logit_len || !label_len || !costs || !workspace || (L > 0 && !labels)) return TFASR_INVALID_VALUE;
  if (workspace_bytes < tfasr_ctc_workspace_bytes(B, T, L, V)) return TFASR_WORKSPACE_TOO_SMALL;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const CtcWs ws = ctc_ws_carve(workspace, B, T, L);
  const unsigned rgrid = (unsigned)ceil_div((int64_t)B * T, kCtcRowWarps);
  ctc_lse_kernel<<<rgrid, kCtcRowWarps * 32, 0, stream>>>(logits, logit_len, B, T, V, ws.lse);
  count_launch();
  const int threads = 32 * (int)ceil_div(2 * L + 1, 32);
  const size_t lat_smem = sizeof(float) * 2 * (threads + 4);
  ctc_lattice_kernel<<<dim3(B, 2), threads, lat_smem, stream>>>(logits, labels, logit_len, label_len, T, L, V, blank, ws, costs);
  count_launch();
  if (grads) {
    const size_t smem = sizeof(float) * (size_t)kCtcRowWarps * V;
    if (smem > 200 * 1024) return TFASR_INVALID_VALUE;
    if (smem > 48 * 1024) {
      if (cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return TFASR_EXECUTION_FAILED;
    }
    ctc_grad_kernel<<<rgrid, kCtcRowWarps * 32, smem, stream>>>(logits, labels, logit_len, label_len, B, T, L, V, blank, ws, costs,
                                                                grad_scale, grads);
    count_launch();
  }
  return check_launch();
}
