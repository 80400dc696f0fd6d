// Plain-SIMT kernels of the ContextNet / DeepSpeech2 encoder bodies, channels-last bf16 activations,
// fp32 parameters and statistics.  HBM-bound byte work: no tensor cores here (the pointwise convs / conv-as-GEMM go to gemm.cu).
//
//   seq_dwconv_*      the depthwise half of SeparableConv1D(padding="causal", strides=s)  — tensorflow_asr/models/encoders/contextnet.py:53-63,
//                     causal pad K-1 frames on the left then VALID (models/layers/convolution.py:25-37)
//   seq_masked_mean   GlobalAveragePooling1D over the valid frames (contextnet.py:137, :156)
//   se_fc_*           fc1 -> activation -> fc2 -> sigmoid of SEModule.call (contextnet.py:157-160), one CTA per utterance
//   se_apply_*        outputs * se (+ residual) -> activation: SEModule.call :162-163 + ConvBlock.call :258-261
//   conv_im2col / conv_col2im   lowering of Conv2D(kernel (kh,kw), strides (sh,sw), "same" | causal) around the tcgen05 GEMM
//                     (models/encoders/deepspeech2.py:56-60,105-117)
//
// This header is also compiled for the CPU by tests/emu (see simt.cuh): keep it to the SIMT vocabulary the emulator provides.
#pragma once
#include "simt.cuh"

namespace tfasr {

constexpr int kSeqTT = 32;       // frames per CTA tile of the depthwise kernels
constexpr int kSeqThreads = 256;

// ------------------------------------------------------------------------------------------------ depthwise conv, causal, strided
// y[b,j,c] = sum_k w[k,c] * x[b, j*s - (K-1) + k, c]                  x [B,T,C], y [B,To,C], To = ceil(T / s)
// REVERSE: dx[b,t,c] = sum_k w[k,c] * dy[b, (t + K-1 - k) / s, c]     over the taps with (t + K-1 - k) % s == 0 and a frame < To
// thread = 8 channels (16-byte accesses) x one frame; tap weights staged in SMEM.   grid (ceil(Tout/kSeqTT), B), block 256
template <bool REVERSE>
__global__ void __launch_bounds__(kSeqThreads) seq_dwconv_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ out, int Tin, int Tout, int C, int K, int stride) {
  TFASR_DYN_SMEM(smem_raw);
  float* sw = reinterpret_cast<float*>(smem_raw);  // [K][C]
  const int tid = threadIdx.x;
  for (int i = tid; i < K * C; i += kSeqThreads) sw[i] = w[i];
  __syncthreads();
  const int CV = C >> 3;
  const int b = blockIdx.y, t0 = blockIdx.x * kSeqTT;
  const __nv_bfloat16* inb = in + (size_t)b * Tin * C;
  __nv_bfloat16* outb = out + (size_t)b * Tout * C;
  for (int i = tid; i < kSeqTT * CV; i += kSeqThreads) {
    const int tt = i / CV, cv = i - tt * CV;
    const int t = t0 + tt;
    if (t >= Tout) break;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int k = 0; k < K; ++k) {
      int src;
      if (!REVERSE) {
        src = t * stride - (K - 1) + k;
        if (src < 0 || src >= Tin) continue;
      } else {
        const int u = t + (K - 1) - k;
        if (u % stride) continue;
        src = u / stride;
        if (src >= Tin) continue;
      }
      float f[8];
      bf8_unpack(ld16(inb + (size_t)src * C + cv * 8), f);
      const float* wk = sw + k * C + cv * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(f[e], wk[e], acc[e]);
    }
    st16(outb + (size_t)t * C + cv * 8, bf8_pack(acc));
  }
}

// dw[k,c] += sum_{b,j} dy[b,j,c] * x[b, j*s - (K-1) + k, c]
// thread (c2, ty): channel pair x rows r = b*To + j of this CTA's strip with stride blockDim.y; per-thread partials for all K taps, reduced
// through SMEM atomics, then one global atomic per (k, c) and CTA.          grid (ceil(C2/64), strips), block (64, 4)
template <int KMAX>
__global__ void __launch_bounds__(256) seq_dwconv_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                               float* __restrict__ dw, int B, int T, int To, int C, int K, int stride,
                                                               int rows_per_block) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_acc = reinterpret_cast<float*>(smem_raw);  // [K][128]
  const int C2 = C >> 1;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  for (int i = tid; i < K * 128; i += nth) s_acc[i] = 0.f;
  __syncthreads();
  const int c2 = blockIdx.x * 64 + threadIdx.x;
  const bool live = c2 < C2;
  float a0[KMAX], a1[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) a0[k] = a1[k] = 0.f;
  const int64_t R = (int64_t)B * To;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
  if (live) {
    const __nv_bfloat162* dy2 = reinterpret_cast<const __nv_bfloat162*>(dy);
    const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(x);
    for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const int b = (int)(r / To), j = (int)(r - (int64_t)b * To);
      const float2 d = __bfloat1622float2(dy2[(size_t)r * C2 + c2]);
      const int tb = j * stride - (K - 1);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int t = tb + k;
        if (k < K && t >= 0 && t < T) {
          const float2 v = __bfloat1622float2(x2[((size_t)b * T + t) * C2 + c2]);
          a0[k] = fmaf(d.x, v.x, a0[k]);
          a1[k] = fmaf(d.y, v.y, a1[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        atomicAdd(&s_acc[k * 128 + 2 * threadIdx.x], a0[k]);
        atomicAdd(&s_acc[k * 128 + 2 * threadIdx.x + 1], a1[k]);
      }
  }
  __syncthreads();
  for (int i = tid; i < K * 128; i += nth) {
    const int k = i >> 7, cc = blockIdx.x * 128 + (i & 127);
    if (cc < C) atomicAdd(dw + (size_t)k * C + cc, s_acc[i]);
  }
}

// ------------------------------------------------------------------------------------------------ masked mean over time
// m[b,c] += (1 / max(len_b, 1)) * sum_{t < len_b, t in this split} y[b,t,c]     (m zeroed by the launcher)
// grid (ceil(C2/64), splits, B), block (64, 4)
__global__ void __launch_bounds__(256) seq_masked_mean_kernel(const __nv_bfloat16* __restrict__ y, const int32_t* __restrict__ length,
                                                              float* __restrict__ m, int T, int C, int frames_per_split) {
  __shared__ float s_acc[128];
  const int C2 = C >> 1;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 128) s_acc[tid] = 0.f;
  __syncthreads();
  const int b = blockIdx.z;
  int len = length[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const int c2 = blockIdx.x * 64 + threadIdx.x;
  const int t0 = blockIdx.y * frames_per_split;
  const int t1 = min(len, t0 + frames_per_split);
  if (c2 < C2 && t0 < t1) {
    const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(y) + (size_t)b * T * C2;
    float s0 = 0.f, s1 = 0.f;
    for (int t = t0 + threadIdx.y; t < t1; t += blockDim.y) {
      const float2 v = __bfloat1622float2(y2[(size_t)t * C2 + c2]);
      s0 += v.x;
      s1 += v.y;
    }
    atomicAdd(&s_acc[2 * threadIdx.x], s0);
    atomicAdd(&s_acc[2 * threadIdx.x + 1], s1);
  }
  __syncthreads();
  if (tid < 128 && t0 < t1) {
    const int cc = blockIdx.x * 128 + tid;
    if (cc < C) atomicAdd(m + (size_t)b * C + cc, s_acc[tid] / (float)max(len, 1));
  }
}

// ------------------------------------------------------------------------------------------------ squeeze-excite fully-connected pair
// z1 = m W1 + b1 [H];  h = act(z1);  s = sigmoid(h W2 + b2) [C].       W1 [C,H], W2 [H,C] row-major fp32.  grid B, block 256
__global__ void __launch_bounds__(256) se_fc_fwd_kernel(const float* __restrict__ m, const float* __restrict__ W1, const float* __restrict__ b1,
                                                        const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ z1,
                                                        float* __restrict__ s, int C, int H, int act) {
  TFASR_DYN_SMEM(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw);  // [C]
  float* sh = sm + C;                              // [H]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) sm[c] = m[(size_t)b * C + c];
  __syncthreads();
  for (int j = tid; j < H; j += 256) {
    float z = b1[j];
    for (int c = 0; c < C; ++c) z = fmaf(sm[c], W1[(size_t)c * H + j], z);
    z1[(size_t)b * H + j] = z;
    sh[j] = simt_act(z, act);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float z = b2[c];
    for (int j = 0; j < H; ++j) z = fmaf(sh[j], W2[(size_t)j * C + c], z);
    s[(size_t)b * C + c] = simt_sigmoid(z);
  }
}

// backward of the pair for one utterance: ds [C] -> dz2 = ds s (1 - s); dh = W2 dz2; dz1 = dh act'(z1); dm = (W1 dz1) / max(len, 1)
// (dm already carries the 1/len of the masked mean).  dz2 / dz1 are kept for the parameter-gradient kernel.      grid B, block 256
__global__ void __launch_bounds__(256) se_fc_bwd_data_kernel(const float* __restrict__ ds, const float* __restrict__ s, const float* __restrict__ z1,
                                                             const float* __restrict__ W1, const float* __restrict__ W2,
                                                             const int32_t* __restrict__ length, float* __restrict__ dz2, float* __restrict__ dz1,
                                                             float* __restrict__ dm, int T, int C, int H, int act) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_dz2 = reinterpret_cast<float*>(smem_raw);  // [C]
  float* s_dz1 = s_dz2 + C;                           // [H]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) {
    const float sg = s[(size_t)b * C + c];
    const float d = ds[(size_t)b * C + c] * sg * (1.0f - sg);
    s_dz2[c] = d;
    dz2[(size_t)b * C + c] = d;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < H; j += 8) {  // warp per hidden unit, lanes stride the C axis (coalesced W2 rows)
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W2[(size_t)j * C + c], s_dz2[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float d = acc * simt_act_grad(z1[(size_t)b * H + j], act);
      s_dz1[j] = d;
      dz1[(size_t)b * H + j] = d;
    }
  }
  __syncthreads();
  int len = length[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const float inv = 1.0f / (float)max(len, 1);
  for (int c = tid; c < C; c += 256) {
    float acc = 0.f;
    for (int j = 0; j < H; ++j) acc = fmaf(W1[(size_t)c * H + j], s_dz1[j], acc);
    dm[(size_t)b * C + c] = acc * inv;
  }
}

// parameter gradients of the pair, one thread per parameter, summed over the batch (single writer: plain +=):
//   dW1[c,j] += sum_b m[b,c] dz1[b,j];  db1[j] += sum_b dz1[b,j];  dW2[j,c] += sum_b act(z1[b,j]) dz2[b,c];  db2[c] += sum_b dz2[b,c]
__global__ void __launch_bounds__(256) se_fc_bwd_params_kernel(const float* __restrict__ m, const float* __restrict__ z1, const float* __restrict__ dz1,
                                                               const float* __restrict__ dz2, float* __restrict__ dW1, float* __restrict__ db1,
                                                               float* __restrict__ dW2, float* __restrict__ db2, int B, int C, int H, int act) {
  const int64_t n1 = (int64_t)C * H, total = 2 * n1 + H + C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (i < n1) {
      const int c = (int)(i / H), j = (int)(i - (int64_t)c * H);
      for (int b = 0; b < B; ++b) acc = fmaf(m[(size_t)b * C + c], dz1[(size_t)b * H + j], acc);
      dW1[i] += acc;
    } else if (i < 2 * n1) {
      const int64_t q = i - n1;
      const int j = (int)(q / C), c = (int)(q - (int64_t)j * C);
      for (int b = 0; b < B; ++b) acc = fmaf(simt_act(z1[(size_t)b * H + j], act), dz2[(size_t)b * C + c], acc);
      dW2[q] += acc;
    } else if (i < 2 * n1 + H) {
      const int j = (int)(i - 2 * n1);
      for (int b = 0; b < B; ++b) acc += dz1[(size_t)b * H + j];
      db1[j] += acc;
    } else {
      const int c = (int)(i - 2 * n1 - H);
      for (int b = 0; b < B; ++b) acc += dz2[(size_t)b * C + c];
      db2[c] += acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ scale (+ residual) + activation
// out[b,t,c] = act(y[b,t,c] * s[b,c] + res[b,t,c])          res may be NULL.  thread = 8 channels, grid-stride over B*T*C/8
__global__ void __launch_bounds__(256) se_apply_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ s,
                                                           const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ out, int64_t rows, int T,
                                                           int C, int act) {
  const int CV = C >> 3;
  const int64_t total = rows * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / CV;
    const int cv = (int)(i - r * CV);
    const int b = (int)(r / T);
    float f[8], g[8];
    bf8_unpack(ld16(y + (size_t)i * 8), f);
    if (res) bf8_unpack(ld16(res + (size_t)i * 8), g);
    const float* sb = s + (size_t)b * C + cv * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = simt_act(fmaf(f[e], sb[e], res ? g[e] : 0.f), act);
    st16(out + (size_t)i * 8, bf8_pack(f));
  }
}

// dpre = dout * act'(y s + res)  -> dres (bf16, = dpre);   ds[b,c] += sum_t dpre * y      (ds zeroed by the launcher)
// grid (ceil(CV/32), splits, B), block (32, 8): thread (cv, ty) walks the frames of its split with stride 8
__global__ void __launch_bounds__(256) se_apply_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                                                           const float* __restrict__ s, const __nv_bfloat16* __restrict__ res,
                                                           __nv_bfloat16* __restrict__ dres, float* __restrict__ ds, int T, int C, int act,
                                                           int frames_per_split) {
  __shared__ float s_acc[256];
  const int CV = C >> 3;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  s_acc[tid] = 0.f;
  __syncthreads();
  const int b = blockIdx.z;
  const int cv = blockIdx.x * 32 + threadIdx.x;
  const int t0 = blockIdx.y * frames_per_split, t1 = min(T, t0 + frames_per_split);
  if (cv < CV) {
    float sc[8], acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = s[(size_t)b * C + cv * 8 + e];
      acc[e] = 0.f;
    }
    for (int t = t0 + threadIdx.y; t < t1; t += blockDim.y) {
      const size_t off = ((size_t)b * T + t) * C + cv * 8;
      float d[8], f[8], g[8];
      bf8_unpack(ld16(dout + off), d);
      bf8_unpack(ld16(y + off), f);
      if (res) bf8_unpack(ld16(res + off), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(f[e], sc[e], res ? g[e] : 0.f);
        d[e] *= simt_act_grad(z, act);
        acc[e] = fmaf(d[e], f[e], acc[e]);
      }
      st16(dres + off, bf8_pack(d));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[threadIdx.x * 8 + e], acc[e]);
  }
  __syncthreads();
  const int cc = blockIdx.x * 256 + tid;
  if (cc < C) atomicAdd(ds + (size_t)b * C + cc, s_acc[tid]);
}

// dy[b,t,c] = dres[b,t,c] * s[b,c] + (t < len_b ? dm[b,c] : 0)           (dm carries the 1/len of the masked mean)
__global__ void __launch_bounds__(256) se_dy_kernel(const __nv_bfloat16* __restrict__ dres, const float* __restrict__ s, const float* __restrict__ dm,
                                                    const int32_t* __restrict__ length, __nv_bfloat16* __restrict__ dy, int64_t rows, int T, int C) {
  const int CV = C >> 3;
  const int64_t total = rows * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / CV;
    const int cv = (int)(i - r * CV);
    const int b = (int)(r / T), t = (int)(r - (int64_t)b * T);
    const bool valid = t < length[b];
    float f[8];
    bf8_unpack(ld16(dres + (size_t)i * 8), f);
    const float* sb = s + (size_t)b * C + cv * 8;
    const float* db = dm + (size_t)b * C + cv * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sb[e], valid ? db[e] : 0.f);
    st16(dy + (size_t)i * 8, bf8_pack(f));
  }
}

// ------------------------------------------------------------------------------------------------ Conv2D lowering (im2col / col2im)
struct ConvGeom {
  int B, T, F, Cin;   // input  [B, T, F, Cin]
  int To, Fo;         // output [B, To, Fo, Cout]
  int kh, kw, sh, sw; // kernel / strides over (T, F)
  int pt, pl;         // zero rows / columns in front ("same": total // 2; causal: k - 1)
  int Kp;             // padded row length of col: round_up(kh * kw * Cin, 8)
};

// col[(b,to,fo), (i*kw + j)*Cin + ci] = x[b, to*sh - pt + i, fo*sw - pl + j, ci]  (0 outside; columns >= kh*kw*Cin are 0)
// thread = 8 consecutive columns of one row; X_F32: the input is the fp32 feature tensor (first layer), else bf16.
template <bool X_F32>
__global__ void __launch_bounds__(256) conv_im2col_kernel(const void* __restrict__ x_, __nv_bfloat16* __restrict__ col, const ConvGeom g) {
  const int KV = g.Kp >> 3, Kreal = g.kh * g.kw * g.Cin;
  const int64_t rows = (int64_t)g.B * g.To * g.Fo, total = rows * KV;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / KV;
    const int kv = (int)(idx - r * KV);
    const int fo = (int)(r % g.Fo);
    const int64_t q = r / g.Fo;
    const int to = (int)(q % g.To), b = (int)(q / g.To);
    float v[8];
    const int k0 = kv * 8;
    const int tap0 = k0 / g.Cin, ci0 = k0 - tap0 * g.Cin;
    bool done = false;
    if (!X_F32 && (g.Cin & 7) == 0 && k0 < Kreal) {  // the 8 columns are 8 consecutive channels of one tap
      const int i = tap0 / g.kw, j = tap0 - i * g.kw;
      const int t = to * g.sh - g.pt + i, f = fo * g.sw - g.pl + j;
      if (t >= 0 && t < g.T && f >= 0 && f < g.F) {
        st16(col + (size_t)idx * 8, ld16(static_cast<const __nv_bfloat16*>(x_) + (((size_t)b * g.T + t) * g.F + f) * g.Cin + ci0));
        done = true;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        float val = 0.f;
        if (k < Kreal) {
          const int tap = k / g.Cin, ci = k - tap * g.Cin;
          const int i = tap / g.kw, j = tap - i * g.kw;
          const int t = to * g.sh - g.pt + i, f = fo * g.sw - g.pl + j;
          if (t >= 0 && t < g.T && f >= 0 && f < g.F) {
            const size_t off = (((size_t)b * g.T + t) * g.F + f) * g.Cin + ci;
            val = X_F32 ? static_cast<const float*>(x_)[off] : __bfloat162float(static_cast<const __nv_bfloat16*>(x_)[off]);
          }
        }
        v[e] = val;
      }
    }
    if (!done) st16(col + (size_t)idx * 8, bf8_pack(v));
  }
}

// da[b,t,f,ci] = sum over taps (i,j) with (t + pt - i) % sh == 0, (f + pl - j) % sw == 0 of dcol[(b,to,fo), (i*kw+j)*Cin + ci]
// thread = 8 channels of one input position (Cin % 8 == 0)
__global__ void __launch_bounds__(256) conv_col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ da, const ConvGeom g) {
  const int CV = g.Cin >> 3;
  const int64_t total = (int64_t)g.B * g.T * g.F * CV;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    int64_t q = idx / CV;
    const int f = (int)(q % g.F);
    q /= g.F;
    const int t = (int)(q % g.T), b = (int)(q / g.T);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int i = 0; i < g.kh; ++i) {
      const int ut = t + g.pt - i;
      if (ut < 0 || ut % g.sh) continue;
      const int to = ut / g.sh;
      if (to >= g.To) continue;
      for (int j = 0; j < g.kw; ++j) {
        const int uf = f + g.pl - j;
        if (uf < 0 || uf % g.sw) continue;
        const int fo = uf / g.sw;
        if (fo >= g.Fo) continue;
        float v[8];
        bf8_unpack(ld16(dcol + (((size_t)b * g.To + to) * g.Fo + fo) * g.Kp + (size_t)(i * g.kw + j) * g.Cin + cv * 8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    }
    st16(da + (size_t)idx * 8, bf8_pack(acc));
  }
}

// ------------------------------------------------------------------------------------------------ launch geometry (shared with tests/emu)
struct SeqLaunch {
  dim3 grid, block;
  size_t smem;
};
inline int seq_split_frames(int T, int B, int xblocks) {  // frames per split so that the grid has ~4 CTAs per SM
  int64_t want = (int64_t)4 * kNumSMs / ((int64_t)(B > 0 ? B : 1) * (xblocks > 0 ? xblocks : 1));
  if (want < 1) want = 1;
  int fps = (int)ceil_div(T, want);
  if (fps < 32) fps = 32;
  return fps;
}
inline SeqLaunch seq_dwconv_launch(int B, int Tout, int C, int K) {
  return {dim3((unsigned)ceil_div(Tout, kSeqTT), (unsigned)B), dim3(kSeqThreads), sizeof(float) * (size_t)K * C};
}
inline SeqLaunch seq_dwconv_wgrad_launch(int B, int To, int C, int K, int* rows_per_block) {
  const int gx = (int)ceil_div(C / 2, 64);
  const int64_t R = (int64_t)B * To;
  int64_t strips = (int64_t)4 * kNumSMs / gx;
  if (strips < 1) strips = 1;
  int64_t rpb = ceil_div(R, strips);
  if (rpb < 64) rpb = 64;
  *rows_per_block = (int)rpb;
  return {dim3((unsigned)gx, (unsigned)ceil_div(R, rpb)), dim3(64, 4), sizeof(float) * (size_t)K * 128};
}
inline unsigned seq_ew_grid(int64_t items) {
  int64_t g = ceil_div(items, 256);
  const int64_t cap = (int64_t)16 * kNumSMs;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace tfasr
