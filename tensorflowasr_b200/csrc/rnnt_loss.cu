// RNN-T loss + gradient for sm_100a.
//
// Replaces the warp-transducer FFI the reference sits behind (tensorflow_asr/losses/impl/rnnt.py:8,55) and its pure-TF
// twin compute_rnnt_loss_and_grad_helper (impl/rnnt.py:181-278: ~60 TF ops, 2 x (T+U-2) tf.scan iterations, a one-hot
// [B,T,U,V] tensor and 4-6 passes over the logits).  Three kernels:
//
//   K1 rnnt_lse_kernel     one warp per (b,t,u) row: log-sum-exp over V, blank / truth log-probs.      HBM: read logits once
//   K2 rnnt_lattice_kernel alpha and beta as skewed wavefronts, one CTA per (utterance, direction): lane u walks t = n - u,
//                          the emit term moves between lanes with a warp shuffle, between warps through SMEM.  Latency-bound
//                          (T + U - 1 dependent steps), touches only the small [B,T,U+1] side arrays (L2-resident).
//   K3 rnnt_grad_kernel    one warp per row: grads = g - softmax * sum(g)  (impl/rnnt.py:233-275), optionally scaled by the
//                          upstream gradient (:321-324).  HBM: read logits a second time + write grads once (may be in place).
//
// Roofline: HBM-bound.  Algorithmic bytes = 8*B*T*U1*V (read logits once + write grads once); practical floor with the
// unavoidable second read = 12*B*T*U1*V.  No tensor cores: the lattice is a log-space scan, not a contraction.
#include "common.cuh"
#include "lattice_lens.cuh"

namespace tfasr {

struct RnntWs {
  float *lse, *blank, *truth, *alpha, *beta;
};

static size_t rnnt_ws_plane(int B, int T, int U1) { return align_up(sizeof(float) * (size_t)B * T * U1, 256); }

static RnntWs rnnt_ws_carve(void* ws, int B, int T, int U1) {
  const size_t plane = rnnt_ws_plane(B, T, U1);
  char* p = static_cast<char*>(ws);
  RnntWs w;
  w.lse = reinterpret_cast<float*>(p);
  w.blank = reinterpret_cast<float*>(p + plane);
  w.truth = reinterpret_cast<float*>(p + 2 * plane);
  w.alpha = reinterpret_cast<float*>(p + 3 * plane);
  w.beta = reinterpret_cast<float*>(p + 4 * plane);
  return w;
}

constexpr int kRowWarps = 8;  // warps per CTA in the row kernels

// ---------------------------------------------------------------------------------------------- K1
// NV4 = float4 per lane held in registers (V <= 128 * NV4).  NV4 == 0: generic scalar path (any V / alignment).
template <int NV4>
__global__ void __launch_bounds__(kRowWarps * 32) rnnt_lse_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                                  const int32_t* __restrict__ logit_len,
                                                                  const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                  RnntWs ws) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int64_t nrows = (int64_t)B * T * U1;
  if (row >= nrows) return;
  const int u = (int)(row % U1);
  const int t = (int)((row / U1) % T);
  const int b = (int)(row / ((int64_t)U1 * T));
  int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
  if (t >= Tb || u > Ub) return;  // rows outside the lattice are never read (ragged batches cost nothing)
  const float* x = logits + row * V;
  float m = -CUDART_INF_F, s = 0.f;
  if constexpr (NV4 > 0) {
    float4 v[NV4];
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int n4 = V >> 2;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      v[i] = (j < n4) ? ld_stream_f4(x4 + j) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    }
#pragma unroll
    for (int i = 0; i < NV4; ++i) m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    m = warp_max(m);
#pragma unroll
    for (int i = 0; i < NV4; ++i) s += __expf(v[i].x - m) + __expf(v[i].y - m) + __expf(v[i].z - m) + __expf(v[i].w - m);
  } else {
    for (int j = lane; j < V; j += 32) m = fmaxf(m, x[j]);
    m = warp_max(m);
    for (int j = lane; j < V; j += 32) s += __expf(x[j] - m);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float lse = m + logf(s);
    ws.lse[row] = lse;
    ws.blank[row] = x[0] - lse;  // just read: L1/L2 hit
    ws.truth[row] = (u < Ub) ? x[labels[(int64_t)b * (U1 - 1) + u]] - lse : -CUDART_INF_F;
  }
}

// ---------------------------------------------------------------------------------------------- K2
// grid (B, 2): y = 0 alpha, y = 1 beta.  blockDim = 32 * ceil(U1 / 32)  (U1 <= 1024).
// Skewed wavefront: at step n, thread i handles lattice cell (tt, i) with tt = n - i in "own" coordinates:
//   alpha: (t, u) = (tt, i)                     a[t,u]  = lae(a[t-1,u] + blank[t-1,u], a[t,u-1] + truth[t,u-1])
//   beta : (t, u) = (Tb-1-tt, Ub-i)             be[t,u] = lae(be[t+1,u] + blank[t,u],  be[t,u+1] + truth[t,u])
// `stay` is the thread's own running value, `emit` arrives from thread i-1 (previous step) by shuffle / SMEM.
constexpr int kPrefetch = 4;

__global__ void __launch_bounds__(1024) rnnt_lattice_kernel(const int32_t* __restrict__ logit_len, const int32_t* __restrict__ label_len,
                                                            int T, int U1, RnntWs ws, float* __restrict__ costs) {
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int i = threadIdx.x;
  const int lane = i & 31, warp = i >> 5;
  int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
  __shared__ float s_edge[2][32];  // [step parity][warp]: emit value leaving the last lane of each warp
  if (Tb <= 0) {                    // empty utterance: no alignment exists
    if (is_beta && i == 0) costs[b] = CUDART_INF_F;
    return;
  }
  const size_t base = (size_t)b * T * U1;
  const float* __restrict__ blank = ws.blank + base;
  const float* __restrict__ truth = ws.truth + base;
  float* __restrict__ out = (is_beta ? ws.beta : ws.alpha) + base;
  const bool active = i <= Ub;
  const int nsteps = Tb + Ub;  // n = 0 .. Tb + Ub - 1
  // address of this thread's cell at own-time tt
  const int u = is_beta ? Ub - i : i;
  auto cell = [&](int tt) -> size_t { return (size_t)(is_beta ? (Tb - 1 - tt) : tt) * U1 + u; };

  // Prefetch ring: slot k serves the steps n with n % kPrefetch == k.  The first such step for this thread
  // is n = i + ((k - i) mod kPrefetch), i.e. own-time tt0 = (k - i) mod kPrefetch.
  float pb[kPrefetch], pt[kPrefetch];
#pragma unroll
  for (int k = 0; k < kPrefetch; ++k) {
    const int tt0 = (((k - i) % kPrefetch) + kPrefetch) % kPrefetch;
    const bool ok = active && tt0 < Tb;
    pb[k] = ok ? blank[cell(tt0)] : 0.f;
    pt[k] = ok ? truth[cell(tt0)] : 0.f;
  }
  float stay = (i == 0) ? 0.f : -CUDART_INF_F;  // alpha: virtual a[-1,0] + blank := 0 ; beta: virtual be[Tb,Ub] := 0
  float emit_out = -CUDART_INF_F;               // value offered to thread i+1 at the next step
  const bool multi_warp = blockDim.x > 32;
  for (int n0 = 0; n0 < nsteps; n0 += kPrefetch) {
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) {
      const int n = n0 + k;
      // emit term from the left neighbour (its value of the previous step)
      float emit_in = __shfl_up_sync(0xffffffffu, emit_out, 1);
      if (lane == 0) emit_in = (warp == 0) ? -CUDART_INF_F : s_edge[(n + 1) & 1][warp - 1];
      const int tt = n - i;
      const bool on = active && tt >= 0 && tt < Tb;
      float new_emit = -CUDART_INF_F;
      if (on) {
        const float bl = pb[k], tr = pt[k];
        float val;
        if (!is_beta) {
          val = log_add_exp_fast(stay, emit_in);
          stay = val + bl;      // feeds a[t+1,u]
          new_emit = val + tr;  // feeds a[t,u+1]   (tr = -inf at u == Ub)
        } else {
          val = log_add_exp_fast(stay + bl, emit_in + tr);  // emit_in = be[t,u+1]; tr = -inf at u == Ub
          stay = val;                                  // be[t,u] feeds be[t-1,u]
          new_emit = val;                              // and be[t,u-1]
        }
        out[cell(tt)] = val;
        if (is_beta && tt == Tb - 1 && i == Ub) costs[b] = -val;  // loss = -beta[0,0]  (impl/rnnt.py:230,277)
        const int tn = tt + kPrefetch;                             // refill this slot for kPrefetch steps ahead
        if (tn < Tb) {
          pb[k] = blank[cell(tn)];
          pt[k] = truth[cell(tn)];
        }
      }
      emit_out = new_emit;
      if (multi_warp) {
        if (lane == 31) s_edge[n & 1][warp] = emit_out;
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- K3
template <int NV4>
__global__ void __launch_bounds__(kRowWarps * 32) rnnt_grad_kernel(const float* logits, const int32_t* __restrict__ labels,
                                                                   const int32_t* __restrict__ logit_len,
                                                                   const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                   RnntWs ws, const float* __restrict__ costs,
                                                                   const float* __restrict__ grad_scale, float* grads) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int64_t nrows = (int64_t)B * T * U1;
  if (row >= nrows) return;
  const int u = (int)(row % U1);
  const int t = (int)((row / U1) % T);
  const int b = (int)(row / ((int64_t)U1 * T));
  int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
  const float* x = logits + row * V;
  float* g = grads + row * V;
  const bool valid = (t < Tb) && (u <= Ub);
  float gb = 0.f, gt = 0.f, lse = 0.f;
  int label = -1;
  if (valid) {
    // d loss / d log-prob of the two transitions leaving (t,u)   (impl/rnnt.py:233-253)
    const float logz = -costs[b];
    const float a = ws.alpha[row];
    lse = ws.lse[row];
    if (t < Tb - 1) {
      gb = -expf(a + ws.beta[row + U1] + ws.blank[row] - logz);
    } else if (u == Ub) {
      gb = -1.0f;  // the reference scatters exactly -1 at (T_b-1, U_b)  (:240-247)
    }
    if (u < Ub) {
      gt = -expf(a + ws.beta[row + 1] + ws.truth[row] - logz);
      label = labels[(int64_t)b * (U1 - 1) + u];
    }
    const float sc = grad_scale ? grad_scale[b] : 1.0f;
    gb *= sc;
    gt *= sc;
  }
  const float gsum = gb + gt;
  if constexpr (NV4 > 0) {
    const int n4 = V >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    if (!valid) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int j = lane + 32 * i;
        if (j < n4) st_stream_f4(g4 + j, z);
      }
      return;
    }
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4 v[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) v[i] = ld_stream_f4(x4 + j);
    }
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        float4 o;
        // grads_logits = g - softmax * sum(g)   (impl/rnnt.py:275)
        o.x = -__expf(v[i].x - lse) * gsum;
        o.y = -__expf(v[i].y - lse) * gsum;
        o.z = -__expf(v[i].z - lse) * gsum;
        o.w = -__expf(v[i].w - lse) * gsum;
        const int c0 = 4 * j;
        if (c0 == 0) o.x += gb;
        if (label >= c0 && label < c0 + 4) {
          const int r = label - c0;
          if (r == 0) o.x += gt;
          else if (r == 1) o.y += gt;
          else if (r == 2) o.z += gt;
          else o.w += gt;
        }
        st_stream_f4(g4 + j, o);
      }
    }
  } else {
    if (!valid) {
      for (int j = lane; j < V; j += 32) g[j] = 0.f;
      return;
    }
    for (int j = lane; j < V; j += 32) {
      float o = -__expf(x[j] - lse) * gsum;
      if (j == 0) o += gb;
      if (j == label) o += gt;
      g[j] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------- K3' (training head)
// Same gradient as K3, written as bf16 with an arbitrary row pitch (ld_g elements) — what the joint's backward GEMMs consume —
// optionally IN PLACE over the fp32 logits (ld_g = 2V: row r's bf16 gradient occupies the first half of row r's fp32 logits,
// which this warp has already pulled into registers).  Persistent: each warp walks rows with a grid stride and keeps per-column
// partial sums in registers, so the vocab-bias gradient costs one atomic per column per CTA instead of one per row.
template <int NV4>
__global__ void __launch_bounds__(kRowWarps * 32) rnnt_grad_bf16_kernel(const float* logits, const int32_t* __restrict__ labels,
                                                                        const int32_t* __restrict__ logit_len,
                                                                        const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                        RnntWs ws, const float* __restrict__ costs,
                                                                        const float* __restrict__ grad_scale, __nv_bfloat16* grads, int64_t ld_g,
                                                                        float* __restrict__ dbias) {
  extern __shared__ float s_cs[];  // [V] column sums of this CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nrows = (int64_t)B * T * U1;
  const int n4 = V >> 2;
  if (dbias) {
    for (int j = threadIdx.x; j < V; j += blockDim.x) s_cs[j] = 0.f;
    __syncthreads();
  }
  float4 cs[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < nrows; row += (int64_t)gridDim.x * kRowWarps) {
    const int u = (int)(row % U1);
    const int t = (int)((row / U1) % T);
    const int b = (int)(row / ((int64_t)U1 * T));
    int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
    uint2* g2 = reinterpret_cast<uint2*>(grads + row * ld_g);
    if (!((t < Tb) && (u <= Ub))) {
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int j = lane + 32 * i;
        if (j < n4) g2[j] = make_uint2(0u, 0u);
      }
      continue;
    }
    const float logz = -costs[b];
    const float a = ws.alpha[row];
    const float lse = ws.lse[row];
    float gb = 0.f, gt = 0.f;
    int label = -1;
    if (t < Tb - 1) gb = -expf(a + ws.beta[row + U1] + ws.blank[row] - logz);
    else if (u == Ub) gb = -1.0f;
    if (u < Ub) {
      gt = -expf(a + ws.beta[row + 1] + ws.truth[row] - logz);
      label = labels[(int64_t)b * (U1 - 1) + u];
    }
    const float sc = grad_scale ? grad_scale[b] : 1.0f;
    gb *= sc;
    gt *= sc;
    const float gsum = gb + gt;
    const float4* x4 = reinterpret_cast<const float4*>(logits + row * V);
    float4 v[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) v[i] = ld_stream_f4(x4 + j);
    }
    __syncwarp();  // in-place mode: every lane has its part of the row before any lane overwrites it
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        float4 o;
        o.x = -__expf(v[i].x - lse) * gsum;
        o.y = -__expf(v[i].y - lse) * gsum;
        o.z = -__expf(v[i].z - lse) * gsum;
        o.w = -__expf(v[i].w - lse) * gsum;
        const int c0 = 4 * j;
        if (c0 == 0) o.x += gb;
        if (label >= c0 && label < c0 + 4) {
          const int r = label - c0;
          if (r == 0) o.x += gt;
          else if (r == 1) o.y += gt;
          else if (r == 2) o.z += gt;
          else o.w += gt;
        }
        cs[i].x += o.x;
        cs[i].y += o.y;
        cs[i].z += o.z;
        cs[i].w += o.w;
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        g2[j] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
  if (dbias) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        atomicAdd(&s_cs[4 * j], cs[i].x);
        atomicAdd(&s_cs[4 * j + 1], cs[i].y);
        atomicAdd(&s_cs[4 * j + 2], cs[i].z);
        atomicAdd(&s_cs[4 * j + 3], cs[i].w);
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < V; j += blockDim.x) atomicAdd(dbias + j, s_cs[j]);
  }
}

static int rnnt_check(const void* logits, const int32_t* labels, const int32_t* tl, const int32_t* ul, int B, int T, int U1, int V,
                      int blank, const void* ws, size_t ws_bytes) {
  if (blank != 0) return TFASR_INVALID_VALUE;  // losses/base_loss.py:24
  if (B < 0 || T < 1 || U1 < 1 || V < 2 || U1 > 1024) return TFASR_INVALID_VALUE;
  if (!logits || !tl || !ul || !ws || (U1 > 1 && !labels)) return TFASR_INVALID_VALUE;
  if (ws_bytes < tfasr_rnnt_workspace_bytes(B, T, U1, V)) return TFASR_WORKSPACE_TOO_SMALL;
  return TFASR_OK;
}

static int rnnt_nv4(const void* p, int V) {
  if ((V & 3) || (reinterpret_cast<uintptr_t>(p) & 15)) return 0;
  if (V <= 128 * 2) return 2;
  if (V <= 128 * 4) return 4;
  if (V <= 128 * 8) return 8;
  if (V <= 128 * 16) return 16;
  return 0;
}

static int rnnt_forward(const float* logits, const int32_t* labels, const int32_t* tl, const int32_t* ul, int B, int T, int U1, int V,
                        float* costs, const RnntWs& ws, cudaStream_t stream) {
  const int64_t nrows = (int64_t)B * T * U1;
  const unsigned grid = (unsigned)ceil_div(nrows, kRowWarps);
  switch (rnnt_nv4(logits, V)) {
    case 2: rnnt_lse_kernel<2><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws); break;
    case 4: rnnt_lse_kernel<4><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws); break;
    case 8: rnnt_lse_kernel<8><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws); break;
    case 16: rnnt_lse_kernel<16><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws); break;
    default: rnnt_lse_kernel<0><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws); break;
  }
  count_launch();
  const int threads = 32 * (int)ceil_div(U1, 32);
  rnnt_lattice_kernel<<<dim3(B, 2), threads, 0, stream>>>(tl, ul, T, U1, ws, costs);
  count_launch();
  return check_launch();
}

static int rnnt_backward(const float* logits, const int32_t* labels, const int32_t* tl, const int32_t* ul, int B, int T, int U1, int V,
                         const float* costs, const float* grad_scale, float* grads, const RnntWs& ws, cudaStream_t stream) {
  const int64_t nrows = (int64_t)B * T * U1;
  const unsigned grid = (unsigned)ceil_div(nrows, kRowWarps);
  int nv4 = rnnt_nv4(logits, V);
  if (reinterpret_cast<uintptr_t>(grads) & 15) nv4 = 0;
  switch (nv4) {
    case 2: rnnt_grad_kernel<2><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs, grad_scale, grads); break;
    case 4: rnnt_grad_kernel<4><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs, grad_scale, grads); break;
    case 8: rnnt_grad_kernel<8><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs, grad_scale, grads); break;
    case 16: rnnt_grad_kernel<16><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs, grad_scale, grads); break;
    default: rnnt_grad_kernel<0><<<grid, kRowWarps * 32, 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs, grad_scale, grads); break;
  }
  count_launch();
  return check_launch();
}

template <int NV4>
static void launch_grad_bf16(const float* logits, const int32_t* labels, const int32_t* tl, const int32_t* ul, int B, int T, int U1, int V,
                             const RnntWs& ws, const float* costs, const float* grad_scale, __nv_bfloat16* grads, int64_t ld_g, float* dbias,
                             cudaStream_t stream) {
  const int64_t nrows = (int64_t)B * T * U1;
  int64_t ctas = ceil_div(nrows, kRowWarps);
  if (ctas > kNumSMs * 8) ctas = kNumSMs * 8;
  rnnt_grad_bf16_kernel<NV4><<<(unsigned)ctas, kRowWarps * 32, dbias ? sizeof(float) * V : 0, stream>>>(logits, labels, tl, ul, B, T, U1, V, ws, costs,
                                                                                                        grad_scale, grads, ld_g, dbias);
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_rnnt_loss_bwd_bf16(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                        int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                                        int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  int st = rnnt_check(logits, labels, logit_len, label_len, B, T, U1, V, blank, workspace, workspace_bytes);
  if (st != TFASR_OK) return st;
  if (!costs || !grads || ld_grads < V || (ld_grads & 3)) return TFASR_INVALID_VALUE;
  if ((V & 3) || V > 2048 || (reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(grads) & 7)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const RnntWs ws = rnnt_ws_carve(const_cast<void*>(workspace), B, T, U1);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  __nv_bfloat16* g = static_cast<__nv_bfloat16*>(grads);
  if (V <= 256) launch_grad_bf16<2>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias, stream);
  else if (V <= 512) launch_grad_bf16<4>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias, stream);
  else if (V <= 1024) launch_grad_bf16<8>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias, stream);
  else launch_grad_bf16<16>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias, stream);
  count_launch();
  return check_launch();
}

extern "C" size_t tfasr_rnnt_workspace_bytes(int32_t B, int32_t T, int32_t U1, int32_t V) {
  (void)V;
  if (B < 0 || T < 1 || U1 < 1) return 0;
  return 5 * rnnt_ws_plane(B, T, U1) + 256;
}

extern "C" int tfasr_rnnt_loss_fwd_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                                       int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs, void* workspace,
                                       size_t workspace_bytes, tfasr_stream_t stream) {
  int st = rnnt_check(logits, labels, logit_len, label_len, B, T, U1, V, blank, workspace, workspace_bytes);
  if (st != TFASR_OK) return st;
  if (!costs) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const RnntWs ws = rnnt_ws_carve(workspace, B, T, U1);
  return rnnt_forward(logits, labels, logit_len, label_len, B, T, U1, V, costs, ws, static_cast<cudaStream_t>(stream));
}

// Forward from the GEMM epilogue's row statistics (rnnt_combine.cu) instead of K1: the logits are not read again.
extern "C" int tfasr_rnnt_loss_fwd_rowstats_f32(const float* logits, const float* rowstats, int32_t parts, const int32_t* labels,
                                                const int32_t* logit_len, const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V,
                                                int32_t blank, float* costs, void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  int st = rnnt_check(logits, labels, logit_len, label_len, B, T, U1, V, blank, workspace, workspace_bytes);
  if (st != TFASR_OK) return st;
  if (!costs || !rowstats || parts < 1) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const RnntWs ws = rnnt_ws_carve(workspace, B, T, U1);
  st = tfasr_rnnt_lse_from_rowstats_f32(logits, rowstats, parts, labels, logit_len, label_len, B, T, U1, V, ws.lse, ws.blank, ws.truth, stream_);
  if (st != TFASR_OK) return st;
  const int threads = 32 * (int)ceil_div(U1, 32);
  rnnt_lattice_kernel<<<dim3(B, 2), threads, 0, static_cast<cudaStream_t>(stream_)>>>(logit_len, label_len, T, U1, ws, costs);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_rnnt_loss_fwd_rowstats_bf16(const void* logits, int64_t ld_logits, const float* rowstats, int32_t parts, const int32_t* labels,
                                                 const int32_t* logit_len, const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V,
                                                 int32_t blank, float* costs, void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  int st = rnnt_check(logits, labels, logit_len, label_len, B, T, U1, V, blank, workspace, workspace_bytes);
  if (st != TFASR_OK) return st;
  if (!costs || !rowstats || parts < 1) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const RnntWs ws = rnnt_ws_carve(workspace, B, T, U1);
  st = tfasr_rnnt_lse_from_rowstats_bf16(logits, ld_logits, rowstats, parts, labels, logit_len, label_len, B, T, U1, V, ws.lse, ws.blank, ws.truth, stream_);
  if (st != TFASR_OK) return st;
  const int threads = 32 * (int)ceil_div(U1, 32);
  rnnt_lattice_kernel<<<dim3(B, 2), threads, 0, static_cast<cudaStream_t>(stream_)>>>(logit_len, label_len, T, U1, ws, costs);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_rnnt_loss_bwd_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                                       int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs,
                                       const float* grad_scale, float* grads, const void* workspace, size_t workspace_bytes,
                                       tfasr_stream_t stream) {
  int st = rnnt_check(logits, labels, logit_len, label_len, B, T, U1, V, blank, workspace, workspace_bytes);
  if (st != TFASR_OK) return st;
  if (!costs || !grads) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const RnntWs ws = rnnt_ws_carve(const_cast<void*>(workspace), B, T, U1);
  return rnnt_backward(logits, labels, logit_len, label_len, B, T, U1, V, costs, grad_scale, grads, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tfasr_rnnt_loss_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                                   int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs, float* grads, void* workspace,
                                   size_t workspace_bytes, tfasr_stream_t stream) {
  int st = tfasr_rnnt_loss_fwd_f32(logits, labels, logit_len, label_len, B, T, U1, V, blank, costs, workspace, workspace_bytes, stream);
  if (st != TFASR_OK || !grads) return st;
  return tfasr_rnnt_loss_bwd_f32(logits, labels, logit_len, label_len, B, T, U1, V, blank, costs, nullptr, grads, workspace,
                                 workspace_bytes, stream);
}
