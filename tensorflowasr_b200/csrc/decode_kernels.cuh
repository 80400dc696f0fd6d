// Persistent greedy transducer decode: the reference's `tf.while_loop` of tiny kernels — one LSTM step + joint + argmax
// per iteration, tensorflow_asr/models/transducer/base_transducer.py:496-575 (recognize_batch) and :577-712 (recognize_single) — as ONE
// kernel launch: one CTA per utterance walks its whole lattice without host round trips.
//
// Per iteration and utterance (call_next, base_transducer.py:437-464):
//   x = embedding[prev_token]                                       (models/layers/embedding.py:41-44)
//   z = x W_ih + h W_hh + b;  i,f,g,o = split(z);  c' = sig(f) c + sig(i) tanh(g);  h' = sig(o) tanh(c')      (keras LSTM, gate order i,f,c,o)
//   y = LayerNorm(h') (eps 1e-3, optional)                          (base_transducer.py:89-96)
//   logits = tanh(e_proj[b, t] + y Wp + bp) Wv + bv;  token = argmax (lowest index on ties; argmax of log_softmax == argmax of logits)
// The encoder-side projection e_proj = enc We + be is computed once for all frames by a GEMM before the launch.
// All arithmetic fp32 with accurate expf / tanhf: greedy tokens are gated bit-exact against the oracle.
//
// Work per iteration is a chain of matrix-vector products whose weights stream from L2 (HBM-/L2-bound byte work, no tensor cores): a CTA
// moves 4*((E+R)*4R + R*J + J*V) bytes per step; `block_matvec` keeps 16-byte loads from every thread in flight and splits the
// contraction over idle threads.  Plain SIMT: compiled a second time for the CPU emulator (tests/emu).
#pragma once
#include "simt.cuh"

namespace tfasr {

struct GreedyParams {
  const float* e_proj;     // [B, T, J]
  const int32_t* nframes;  // [B] encoder lengths
  const float* emb;        // [V, E]
  const float* w_ih;       // [E, 4R]
  const float* w_hh;       // [R, 4R]
  const float* b;          // [4R]
  const float* ln_g;       // [R] or NULL (no LayerNorm)
  const float* ln_b;
  const float* wp;         // [R, J]
  const float* bp;         // [J]
  const float* wv;         // [J, V]
  const float* bv;         // [V]
  int32_t* prev_tok;       // [B]      in/out
  float* states;           // [B, 2, R] in/out (h, c)
  int32_t* frame_idx;      // [B]      in/out
  int32_t* tok_idx;        // [B]      in/out (batch modes: next write column; single mode: last written index)
  int32_t* tokens;         // [B, max_tokens] in/out
  int32_t* iters;          // [B]  mode 1: in = iteration budget;  modes 0 / 2: out = iterations executed (mode 0: -1 if the cap was hit first)
  int32_t* cap_iter;       // [B]  mode 0: out = first iteration at which tok_idx >= max_tokens - 1 held at the loop test, or -1
  int T, E, R, J, V, max_tokens, blank;
  int mode;                // 0: recognize_batch until this utterance's own frame condition; 1: recognize_batch for exactly iters[b] more
                           // iterations; 2: recognize_single (max_tokens_per_frame = per_frame_cap)
  int max_iters;           // mode 0 cap
  int per_frame_cap;       // mode 2
  int part_cap;            // floats available in the split-K scratch
};

constexpr int kMvMaxSplit = 8;

// block-wide sum of one float per thread (all threads get the result); scratch: >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += scratch[w];
  return t;
}

// out[n] (+)= bias[n] + sum_k x[k] W[k*N + n]   for n < N.  x, out, part in shared memory; W row-major in global memory.
// N % 4 == 0 and W 16-byte aligned: thread = 4 adjacent columns x a slice of the contraction (ks slices, partials reduced through `part`);
// otherwise one column per thread.  Ends with a barrier: `out` is ready for every thread.
__device__ __forceinline__ void block_matvec(const float* x, int K, const float* __restrict__ W, int N, const float* __restrict__ bias, float* out,
                                             float* part, int part_cap, bool accumulate) {
  const int nth = blockDim.x, tid = threadIdx.x;
  if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    const int N4 = N >> 2;
    int ks = nth / N4;
    if (ks > kMvMaxSplit) ks = kMvMaxSplit;
    if (ks > part_cap / N) ks = part_cap / N;
    if (ks > K) ks = K;
    if (ks < 1) ks = 1;
    const int kc = (K + ks - 1) / ks;
    if (ks == 1) {  // no scratch needed: accumulate straight into out
      for (int cg = tid; cg < N4; cg += nth) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* Wp = reinterpret_cast<const float4*>(W) + cg;
        int k = 0;
        for (; k + 4 <= K; k += 4) {
          const float4 w0 = Wp[(size_t)k * N4], w1 = Wp[(size_t)(k + 1) * N4], w2 = Wp[(size_t)(k + 2) * N4], w3 = Wp[(size_t)(k + 3) * N4];
          const float x0 = x[k], x1 = x[k + 1], x2 = x[k + 2], x3 = x[k + 3];
          a.x = fmaf(x0, w0.x, a.x); a.y = fmaf(x0, w0.y, a.y); a.z = fmaf(x0, w0.z, a.z); a.w = fmaf(x0, w0.w, a.w);
          a.x = fmaf(x1, w1.x, a.x); a.y = fmaf(x1, w1.y, a.y); a.z = fmaf(x1, w1.z, a.z); a.w = fmaf(x1, w1.w, a.w);
          a.x = fmaf(x2, w2.x, a.x); a.y = fmaf(x2, w2.y, a.y); a.z = fmaf(x2, w2.z, a.z); a.w = fmaf(x2, w2.w, a.w);
          a.x = fmaf(x3, w3.x, a.x); a.y = fmaf(x3, w3.y, a.y); a.z = fmaf(x3, w3.z, a.z); a.w = fmaf(x3, w3.w, a.w);
        }
        for (; k < K; ++k) {
          const float4 w0 = Wp[(size_t)k * N4];
          const float x0 = x[k];
          a.x = fmaf(x0, w0.x, a.x); a.y = fmaf(x0, w0.y, a.y); a.z = fmaf(x0, w0.z, a.z); a.w = fmaf(x0, w0.w, a.w);
        }
        const int n = cg * 4;
        const float r[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = r[e] + (bias ? bias[n + e] : 0.f);
          if (accumulate) v += out[n + e];
          out[n + e] = v;
        }
      }
      __syncthreads();
      return;
    }
    for (int it = tid; it < N4 * ks; it += nth) {
      const int s = it / N4, cg = it - s * N4;
      const int k0 = s * kc, k1 = (k0 + kc < K) ? k0 + kc : K;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* Wp = reinterpret_cast<const float4*>(W) + cg;
      int k = k0;
      for (; k + 4 <= k1; k += 4) {
        const float4 w0 = Wp[(size_t)k * N4], w1 = Wp[(size_t)(k + 1) * N4], w2 = Wp[(size_t)(k + 2) * N4], w3 = Wp[(size_t)(k + 3) * N4];
        const float x0 = x[k], x1 = x[k + 1], x2 = x[k + 2], x3 = x[k + 3];
        a.x = fmaf(x0, w0.x, a.x); a.y = fmaf(x0, w0.y, a.y); a.z = fmaf(x0, w0.z, a.z); a.w = fmaf(x0, w0.w, a.w);
        a.x = fmaf(x1, w1.x, a.x); a.y = fmaf(x1, w1.y, a.y); a.z = fmaf(x1, w1.z, a.z); a.w = fmaf(x1, w1.w, a.w);
        a.x = fmaf(x2, w2.x, a.x); a.y = fmaf(x2, w2.y, a.y); a.z = fmaf(x2, w2.z, a.z); a.w = fmaf(x2, w2.w, a.w);
        a.x = fmaf(x3, w3.x, a.x); a.y = fmaf(x3, w3.y, a.y); a.z = fmaf(x3, w3.z, a.z); a.w = fmaf(x3, w3.w, a.w);
      }
      for (; k < k1; ++k) {
        const float4 w0 = Wp[(size_t)k * N4];
        const float x0 = x[k];
        a.x = fmaf(x0, w0.x, a.x); a.y = fmaf(x0, w0.y, a.y); a.z = fmaf(x0, w0.z, a.z); a.w = fmaf(x0, w0.w, a.w);
      }
      float* pp = part + (size_t)s * N + cg * 4;
      pp[0] = a.x; pp[1] = a.y; pp[2] = a.z; pp[3] = a.w;
    }
    __syncthreads();
    for (int n = tid; n < N; n += nth) {
      float v = bias ? bias[n] : 0.f;
      for (int s = 0; s < ks; ++s) v += part[(size_t)s * N + n];
      if (accumulate) v += out[n];
      out[n] = v;
    }
    __syncthreads();
    return;
  }
  for (int n = tid; n < N; n += nth) {
    float v = bias ? bias[n] : 0.f;
    for (int k = 0; k < K; ++k) v = fmaf(x[k], W[(size_t)k * N + n], v);
    if (accumulate) v += out[n];
    out[n] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// One call_next for this CTA's utterance at encoder frame t.  Reads s_h / s_c (current state) and prev_tok; leaves the candidate state in
// s_hn / s_cn and returns the arg-max token (identical in every thread).
__device__ __forceinline__ int greedy_step(const GreedyParams& p, int b, int t, int prev_tok, float* s_x, const float* s_h, const float* s_c, float* s_hn,
                                           float* s_cn, float* s_y, float* s_g, float* s_j, float* s_part, float* s_red) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int E = p.E, R = p.R, J = p.J, V = p.V;
  for (int k = tid; k < E; k += nth) s_x[k] = p.emb[(size_t)prev_tok * E + k];
  __syncthreads();
  block_matvec(s_x, E, p.w_ih, 4 * R, p.b, s_g, s_part, p.part_cap, false);
  block_matvec(s_h, R, p.w_hh, 4 * R, nullptr, s_g, s_part, p.part_cap, true);
  for (int r = tid; r < R; r += nth) {
    const float ig = sigmoid_acc(s_g[r]), fg = sigmoid_acc(s_g[R + r]), gg = tanhf(s_g[2 * R + r]), og = sigmoid_acc(s_g[3 * R + r]);
    const float c = fg * s_c[r] + ig * gg;
    s_cn[r] = c;
    s_hn[r] = og * tanhf(c);
  }
  __syncthreads();
  const float* y = s_hn;
  if (p.ln_g) {
    float part = 0.f;
    for (int r = tid; r < R; r += nth) part += s_hn[r];
    const float mean = block_sum(part, s_red) / (float)R;
    part = 0.f;
    for (int r = tid; r < R; r += nth) {
      const float d = s_hn[r] - mean;
      part = fmaf(d, d, part);
    }
    const float rstd = rsqrtf(block_sum(part, s_red) / (float)R + 1e-3f);
    for (int r = tid; r < R; r += nth) s_y[r] = (s_hn[r] - mean) * rstd * p.ln_g[r] + p.ln_b[r];
    __syncthreads();
    y = s_y;
  }
  block_matvec(y, R, p.wp, J, p.bp, s_j, s_part, p.part_cap, false);
  const float* e = p.e_proj + ((size_t)b * p.T + t) * J;
  for (int j = tid; j < J; j += nth) s_j[j] = tanhf(e[j] + s_j[j]);
  __syncthreads();
  // logits land in s_g when they fit (V <= 4R), else in the tail of the split scratch reserved by the launcher
  float* s_logit = s_g;
  block_matvec(s_j, J, p.wv, V, p.bv, s_logit, s_part, p.part_cap, false);
  float best = -CUDART_INF_F;
  int arg = 0x7fffffff;
  for (int v = tid; v < V; v += nth) {
    const float l = s_logit[v];
    if (l > best) {  // ascending v per thread: the first maximum is kept
      best = l;
      arg = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
  }
  const int lane = tid & 31, warp = tid >> 5, nw = (nth + 31) >> 5;
  int* s_arg = reinterpret_cast<int*>(s_red + 32);
  __syncthreads();
  if (lane == 0) {
    s_red[warp] = best;
    s_arg[warp] = arg;
  }
  __syncthreads();
  best = s_red[0];
  arg = s_arg[0];
  for (int w = 1; w < nw; ++w) {
    const float ob = s_red[w];
    const int oa = s_arg[w];
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
  }
  __syncthreads();
  return arg;
}

// grid B, block `threads` (multiple of 32).  Dynamic shared memory: E + 5R + max(4R, V) + J + part_cap + 64 floats.
__global__ void __launch_bounds__(1024) greedy_decode_kernel(const GreedyParams p) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_x = reinterpret_cast<float*>(smem_raw);
  float* s_h = s_x + p.E;
  float* s_c = s_h + p.R;
  float* s_hn = s_c + p.R;
  float* s_cn = s_hn + p.R;
  float* s_y = s_cn + p.R;
  float* s_g = s_y + p.R;  // [max(4R, V)]
  const int G = 4 * p.R > p.V ? 4 * p.R : p.V;
  float* s_j = s_g + G;
  float* s_red = s_j + p.J;  // [64]
  float* s_part = s_red + 64;
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int R = p.R;
  for (int r = tid; r < R; r += nth) {
    s_h[r] = p.states[((size_t)b * 2 + 0) * R + r];
    s_c[r] = p.states[((size_t)b * 2 + 1) * R + r];
  }
  __syncthreads();
  const int nframes = p.nframes[b];
  int prev_tok = p.prev_tok[b];
  int frame = p.frame_idx[b], tok_idx = p.tok_idx[b];
  int32_t* toks = p.tokens + (size_t)b * p.max_tokens;
  int it = 0, cap_it = -1, result_it = 0;
  if (p.mode == 2) {
    // recognize_single (base_transducer.py:577-712): a frame is left on blank or after per_frame_cap symbols; tok_idx = last written index
    int per_frame = 0;
    while (frame < nframes) {
      const int cur = greedy_step(p, b, frame, prev_tok, s_x, s_h, s_c, s_hn, s_cn, s_y, s_g, s_j, s_part, s_red);
      if (cur == p.blank) {
        ++frame;
        per_frame = 0;
      } else {
        ++per_frame;
        ++tok_idx;
        if (tid == 0 && tok_idx < p.max_tokens) toks[tok_idx] = cur;
        prev_tok = cur;
        for (int r = tid; r < R; r += nth) {
          s_h[r] = s_hn[r];
          s_c[r] = s_cn[r];
        }
        __syncthreads();
        if (per_frame >= p.per_frame_cap) {
          ++frame;
          per_frame = 0;
        }
      }
      ++it;
    }
    result_it = it;
  } else {
    // recognize_batch (base_transducer.py:496-575), this utterance's trajectory; quirks kept: the clamped last frame is re-read until
    // frame_idx > nframes, emitted symbols are stored from column tok_idx + 1 on, blanks overwrite column 0
    const int budget = (p.mode == 1) ? p.iters[b] : 0;
    for (;;) {
      if (p.mode == 0) {
        if (cap_it < 0 && tok_idx >= p.max_tokens - 1) cap_it = it;
        if (frame >= nframes - 1) {
          result_it = it;
          break;
        }
        if (it >= p.max_iters) {
          result_it = -1;
          break;
        }
      } else if (it >= budget) {
        result_it = it;
        break;
      }
      int t = frame < nframes - 1 ? frame : nframes - 1;
      if (t < 0) t = 0;
      const int cur = greedy_step(p, b, t, prev_tok, s_x, s_h, s_c, s_hn, s_cn, s_y, s_g, s_j, s_part, s_red);
      const bool eq_blank = (cur == p.blank) || (tok_idx >= p.max_tokens) || (frame > nframes);
      if (eq_blank) {
        if (tid == 0) toks[0] = p.blank;
        ++frame;
      } else {
        const int nxt = tok_idx + 1 < p.max_tokens - 1 ? tok_idx + 1 : p.max_tokens - 1;
        if (tid == 0) toks[nxt] = cur;
        tok_idx = nxt;
        prev_tok = cur;
        for (int r = tid; r < R; r += nth) {
          s_h[r] = s_hn[r];
          s_c[r] = s_cn[r];
        }
        __syncthreads();
      }
      ++it;
    }
  }
  __syncthreads();  // every thread has read this utterance's entry state (a loop of zero iterations has no barrier of its own)
  for (int r = tid; r < R; r += nth) {
    p.states[((size_t)b * 2 + 0) * R + r] = s_h[r];
    p.states[((size_t)b * 2 + 1) * R + r] = s_c[r];
  }
  if (tid == 0) {
    p.prev_tok[b] = prev_tok;
    p.frame_idx[b] = frame;
    p.tok_idx[b] = tok_idx;
    if (p.mode != 1) p.iters[b] = result_it;
    if (p.mode == 0) p.cap_iter[b] = cap_it;
  }
}

// floats of dynamic shared memory for given sizes and split scratch capacity
inline size_t greedy_smem_floats(int E, int R, int J, int V, int part_cap) {
  const int G = 4 * R > V ? 4 * R : V;
  return (size_t)E + 5 * (size_t)R + G + J + 64 + part_cap;
}

}  // namespace tfasr
