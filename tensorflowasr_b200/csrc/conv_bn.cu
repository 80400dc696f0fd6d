// Conv2dSubsampling and BatchNorm kernels for sm_100a (channels-last bf16 activations, fp32 statistics).
//
// Replaces, around the tcgen05 GEMM that does the dense part of conv 2:
//   Conv2D(3x3, stride 2, causal) #1, Cin = 1         tensorflow_asr/models/layers/subsampling.py:218-230 + convolution.py:132-144
//   im2col / col2im of Conv2D #2 (Cin = 144)          (the reference hands this to cuDNN)
//   BatchNormalization(sync) + swish, fwd and bwd      subsampling.py:199-206, conformer.py:326-333,371-372 (keras BN: eps 1e-3,
//                                                      batch statistics over every leading axis, biased variance)
// All of these are HBM-bound streaming kernels; per-channel reductions use CTA-level partial sums + one atomic per channel per CTA.
#include "common.cuh"

namespace tfasr {

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------- conv 1 (Cin = 1), forward
// x [B, T, F] f32 -> y [B, T1, F1, C] bf16 with T1 = ceil(T/2), F1 = ceil(F/2); causal pad (2 top, 2 left), VALID, stride 2.
// w [3, 3, 1, C] f32, bias [C].  block: (C/2 threads, PY positions); each thread owns a channel pair.
constexpr int kConv1PY = 8;
__global__ void __launch_bounds__(1024) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ y, int B, int T, int F, int T1, int F1, int C,
                                                         int pos_per_block) {
  const int c2 = threadIdx.x;  // channel pair
  float w0[9], w1[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w0[k] = w[k * C + 2 * c2];
    w1[k] = w[k * C + 2 * c2 + 1];
  }
  const float b0 = bias[2 * c2], b1 = bias[2 * c2 + 1];
  const int64_t npos = (int64_t)B * T1 * F1;
  const int64_t p0 = (int64_t)blockIdx.x * pos_per_block;
  for (int64_t p = p0 + threadIdx.y; p < min(npos, p0 + pos_per_block); p += blockDim.y) {
    const int f1 = (int)(p % F1);
    const int t1 = (int)((p / F1) % T1);
    const int b = (int)(p / ((int64_t)F1 * T1));
    const float* xb = x + (size_t)b * T * F;
    float a0 = b0, a1 = b1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int t = 2 * t1 + kh - 2;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 + kw - 2;
        const float v = (t >= 0 && t < T && f >= 0 && f < F) ? __ldg(xb + (size_t)t * F + f) : 0.f;
        a0 = fmaf(v, w0[kh * 3 + kw], a0);
        a1 = fmaf(v, w1[kh * 3 + kw], a1);
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(y + (size_t)p * C + 2 * c2) = __floats2bfloat162_rn(a0, a1);
  }
}

// conv 1 weight gradient: dw[k, c] += sum_pos dy[pos, c] * x[pos shifted by tap k]
__global__ void __launch_bounds__(1024) conv1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                                                           int B, int T, int F, int T1, int F1, int C, int pos_per_block) {
  extern __shared__ float s_dw[];  // [9][C]
  const int c2 = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  for (int i = tid; i < 9 * C; i += nth) s_dw[i] = 0.f;
  __syncthreads();
  float g0[9], g1[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) g0[k] = g1[k] = 0.f;
  const int64_t npos = (int64_t)B * T1 * F1;
  const int64_t p0 = (int64_t)blockIdx.x * pos_per_block;
  for (int64_t p = p0 + threadIdx.y; p < min(npos, p0 + pos_per_block); p += blockDim.y) {
    const int f1 = (int)(p % F1);
    const int t1 = (int)((p / F1) % T1);
    const int b = (int)(p / ((int64_t)F1 * T1));
    const float* xb = x + (size_t)b * T * F;
    const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(dy + (size_t)p * C + 2 * c2);
    const float d0 = __low2float(d), d1 = __high2float(d);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int t = 2 * t1 + kh - 2;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 + kw - 2;
        const float v = (t >= 0 && t < T && f >= 0 && f < F) ? __ldg(xb + (size_t)t * F + f) : 0.f;
        g0[kh * 3 + kw] = fmaf(v, d0, g0[kh * 3 + kw]);
        g1[kh * 3 + kw] = fmaf(v, d1, g1[kh * 3 + kw]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    atomicAdd(&s_dw[k * C + 2 * c2], g0[k]);
    atomicAdd(&s_dw[k * C + 2 * c2 + 1], g1[k]);
  }
  __syncthreads();
  for (int i = tid; i < 9 * C; i += nth) atomicAdd(dw + i, s_dw[i]);
}

// ---------------------------------------------------------------------------------------------- im2col / col2im (3x3, stride 2, causal)
// a [B, T1, F1, C] bf16 -> col [B*T2*F2, 9*C] bf16 (column order kh, kw, c = the Keras kernel [kh,kw,cin,cout] flattened).
// One thread per 16-byte chunk (8 channels).
__global__ void __launch_bounds__(256) im2col_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ col, int B, int T1, int F1,
                                                     int T2, int F2, int C) {
  const int c8n = C >> 3;
  const int64_t total = (int64_t)B * T2 * F2 * 9 * c8n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const int k = (int)((i / c8n) % 9);
    const int64_t p = i / ((int64_t)c8n * 9);
    const int f2 = (int)(p % F2);
    const int t2 = (int)((p / F2) % T2);
    const int b = (int)(p / ((int64_t)F2 * T2));
    const int t = 2 * t2 + k / 3 - 2, f = 2 * f2 + k % 3 - 2;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < T1 && f >= 0 && f < F1) v = *reinterpret_cast<const uint4*>(a + (((size_t)b * T1 + t) * F1 + f) * C + c8 * 8);
    *reinterpret_cast<uint4*>(col + (size_t)p * 9 * C + (size_t)k * C + c8 * 8) = v;
  }
}

// dcol [B*T2*F2, 9*C] -> da [B, T1, F1, C]: each input element gathers the <= 4 taps that read it (no atomics).
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ da, int B, int T1, int F1,
                                                     int T2, int F2, int C) {
  const int c8n = C >> 3;
  const int64_t total = (int64_t)B * T1 * F1 * c8n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const int64_t q = i / c8n;
    const int f = (int)(q % F1);
    const int t = (int)((q / F1) % T1);
    const int b = (int)(q / ((int64_t)F1 * T1));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int tt = t + 2 - kh;  // = 2 * t2
      if (tt < 0 || (tt & 1)) continue;
      const int t2 = tt >> 1;
      if (t2 >= T2) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ff = f + 2 - kw;
        if (ff < 0 || (ff & 1)) continue;
        const int f2 = ff >> 1;
        if (f2 >= F2) continue;
        const size_t p = ((size_t)b * T2 + t2) * F2 + f2;
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + p * 9 * C + (size_t)(kh * 3 + kw) * C + c8 * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(da + (size_t)q * C + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------------------------------------- BatchNorm (+ swish), channels-last [R, C]
// stats: sums[0][c] += sum_r x, sums[1][c] += sum_r x^2.   block (C/2, RY).
__global__ void __launch_bounds__(1024) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums, int64_t R, int C,
                                                        int rows_per_block) {
  extern __shared__ float s_acc[];  // [2][C]
  const int c2 = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  for (int i = tid; i < 2 * C; i += nth) s_acc[i] = 0.f;
  __syncthreads();
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t r = r0 + threadIdx.y; r < min(R, r0 + rows_per_block); r += blockDim.y) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + (size_t)r * C + 2 * c2);
    const float a = __low2float(v), b = __high2float(v);
    s0 += a;
    s1 += b;
    q0 = fmaf(a, a, q0);
    q1 = fmaf(b, b, q1);
  }
  atomicAdd(&s_acc[2 * c2], s0);
  atomicAdd(&s_acc[2 * c2 + 1], s1);
  atomicAdd(&s_acc[C + 2 * c2], q0);
  atomicAdd(&s_acc[C + 2 * c2 + 1], q1);
  __syncthreads();
  for (int i = tid; i < 2 * C; i += nth) atomicAdd(sums + i, s_acc[i]);
}

// y = act(x * scale[c] + shift[c]);  act: 0 none, 1 swish, 2 relu.   scale = gamma * rstd, shift = beta - mean * scale.
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, int64_t R, int C, int act) {
  const int c2n = C >> 1;
  const int64_t total = R * c2n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c2 = (int)(i % c2n);
    const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(x)[i];
    float a = fmaf(__low2float(v), scale[2 * c2], shift[2 * c2]);
    float b = fmaf(__high2float(v), scale[2 * c2 + 1], shift[2 * c2 + 1]);
    if (act == 1) {
      a *= sigmoid_f(a);
      b *= sigmoid_f(b);
    } else if (act == 2) {
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    }
    reinterpret_cast<__nv_bfloat162*>(y)[i] = __floats2bfloat162_rn(a, b);
  }
}

// backward pass 1: d = dy * act'(x * scale + shift);  sums[0][c] += sum d, sums[1][c] += sum d * xhat   (xhat = (x - mean) * rstd)
__global__ void __launch_bounds__(1024) bn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float* __restrict__ sums, int64_t R, int C, int act, int rows_per_block) {
  extern __shared__ float s_acc[];
  const int c2 = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  for (int i = tid; i < 2 * C; i += nth) s_acc[i] = 0.f;
  __syncthreads();
  const float sc0 = scale[2 * c2], sc1 = scale[2 * c2 + 1], sh0 = shift[2 * c2], sh1 = shift[2 * c2 + 1];
  const float m0 = mean[2 * c2], m1 = mean[2 * c2 + 1], rs0 = rstd[2 * c2], rs1 = rstd[2 * c2 + 1];
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t r = r0 + threadIdx.y; r < min(R, r0 + rows_per_block); r += blockDim.y) {
    const size_t idx = (size_t)r * C + 2 * c2;
    const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(x + idx);
    const __nv_bfloat162 dv = *reinterpret_cast<const __nv_bfloat162*>(dy + idx);
    const float x0 = __low2float(xv), x1 = __high2float(xv);
    float d0 = __low2float(dv), d1 = __high2float(dv);
    if (act == 1) {
      const float z0 = fmaf(x0, sc0, sh0), z1 = fmaf(x1, sc1, sh1);
      const float g0 = sigmoid_f(z0), g1 = sigmoid_f(z1);
      d0 *= g0 * (1.0f + z0 * (1.0f - g0));
      d1 *= g1 * (1.0f + z1 * (1.0f - g1));
    } else if (act == 2) {
      if (fmaf(x0, sc0, sh0) <= 0.f) d0 = 0.f;
      if (fmaf(x1, sc1, sh1) <= 0.f) d1 = 0.f;
    }
    s0 += d0;
    s1 += d1;
    q0 = fmaf(d0, (x0 - m0) * rs0, q0);
    q1 = fmaf(d1, (x1 - m1) * rs1, q1);
  }
  atomicAdd(&s_acc[2 * c2], s0);
  atomicAdd(&s_acc[2 * c2 + 1], s1);
  atomicAdd(&s_acc[C + 2 * c2], q0);
  atomicAdd(&s_acc[C + 2 * c2 + 1], q1);
  __syncthreads();
  for (int i = tid; i < 2 * C; i += nth) atomicAdd(sums + i, s_acc[i]);
}

// Sentinel for `count`: the row count is the float stored at sums[2 * C] (sync-BatchNorm over replicas whose row counts differ: every
// replica puts its own R there before the all-reduce, so the reduced vector carries the global count; keras counts the same way).
constexpr float kBnCountOnDevice = -2.0f;

// backward pass 2: dx = gamma * rstd * (d - mean_d - xhat * mean_dxhat),  mean_d = sums[0]/count, mean_dxhat = sums[1]/count.
// (inference-style BN with fixed statistics: pass count <= 0 -> dx = scale * d)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ sums, float count, __nv_bfloat16* __restrict__ dx, int64_t R,
                                                           int C, int act) {
  const int c2n = C >> 1;
  const int64_t total = R * c2n;
  if (count == kBnCountOnDevice) count = sums[2 * C];  // the all-reduced row count rides behind the sums
  const float inv = count > 0.f ? 1.0f / count : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c2 = (int)(i % c2n);
    const __nv_bfloat162 xv = reinterpret_cast<const __nv_bfloat162*>(x)[i];
    const __nv_bfloat162 dv = reinterpret_cast<const __nv_bfloat162*>(dy)[i];
    float o[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 2 * c2 + h;
      const float xx = h ? __high2float(xv) : __low2float(xv);
      float d = h ? __high2float(dv) : __low2float(dv);
      if (act == 1) {
        const float z = fmaf(xx, scale[c], shift[c]);
        const float g = sigmoid_f(z);
        d *= g * (1.0f + z * (1.0f - g));
      } else if (act == 2) {
        if (fmaf(xx, scale[c], shift[c]) <= 0.f) d = 0.f;
      }
      const float xh = (xx - mean[c]) * rstd[c];
      o[h] = scale[c] * (d - sums[c] * inv - xh * sums[C + c] * inv);
    }
    reinterpret_cast<__nv_bfloat162*>(dx)[i] = __floats2bfloat162_rn(o[0], o[1]);
  }
}

// ---------------------------------------------------------------------------------------------- vector variants (C % 8 == 0)
// All four BatchNorm passes as one templated kernel: thread = (8-channel vector, row lane), 16-byte loads / stores, per-channel
// parameters in registers, rows of a thread unrolled 4-fold (independent loads in flight), statistics reduced in SMEM then one
// atomic per channel and CTA.  The scalar kernels above stay for odd channel counts.
//   MODE 0 stats      sums[0] += x, sums[1] += x^2
//   MODE 1 apply      y = act(x * scale + shift)
//   MODE 2 bwd stats  d = dy * act'(x * scale + shift); sums[0] += d, sums[1] += d * xhat
//   MODE 3 bwd apply  dx = scale * (d - sums[0] / count - xhat * sums[1] / count)
struct BnVecArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  __nv_bfloat16* out;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* sums_in;
  float* sums_out;
  int64_t R;
  int C, act, rows_per_block;
  float inv_count;
  int count_on_device;  // MODE 3: 1 / sums_in[2 C] instead of inv_count
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    pk[j] = *reinterpret_cast<const uint32_t*>(&t);
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

template <int MODE>
__global__ void __launch_bounds__(256) bn_vec_kernel(const BnVecArgs a) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  extern __shared__ float s_acc[];  // [2][C] (stats modes)
  const int C = a.C, cv = threadIdx.x, ty = threadIdx.y, RY = blockDim.y;
  const int tid = ty * blockDim.x + cv, nth = blockDim.x * blockDim.y;
  if (MODE == 0 || MODE == 2) {
    for (int i = tid; i < 2 * C; i += nth) s_acc[i] = 0.f;
    __syncthreads();
  }
  float sc[8], sh[8], mu[8], rs[8], k0[8], k1[8];
  [[maybe_unused]] float inv_count = a.inv_count;
  if (MODE == 3 && a.count_on_device) {
    const float n = a.sums_in[2 * C];
    inv_count = n > 0.f ? 1.0f / n : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    sc[j] = (MODE != 0) ? a.scale[c] : 0.f;
    sh[j] = (MODE != 0) ? a.shift[c] : 0.f;
    mu[j] = (MODE >= 2) ? a.mean[c] : 0.f;
    rs[j] = (MODE >= 2) ? a.rstd[c] : 0.f;
    k0[j] = (MODE == 3) ? a.sums_in[c] * inv_count : 0.f;
    k1[j] = (MODE == 3) ? a.sums_in[C + c] * inv_count : 0.f;
  }
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_block;
  const int64_t r1 = r0 + a.rows_per_block < a.R ? r0 + a.rows_per_block : a.R;
  for (int64_t rb = r0 + ty; rb < r1; rb += 4 * RY) {
    uint4 xv[4], dv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = rb + (int64_t)k * RY;
      if (r < r1) {
        xv[k] = *reinterpret_cast<const uint4*>(a.x + (size_t)r * C + cv * 8);
        if (MODE >= 2) dv[k] = *reinterpret_cast<const uint4*>(a.dy + (size_t)r * C + cv * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = rb + (int64_t)k * RY;
      if (r >= r1) break;
      float xf[8], df[8], o[8];
      unpack8(xv[k], xf);
      if (MODE >= 2) unpack8(dv[k], df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) {
          s0[j] += xf[j];
          s1[j] = fmaf(xf[j], xf[j], s1[j]);
        } else if (MODE == 1) {
          float z = fmaf(xf[j], sc[j], sh[j]);
          if (a.act == 1) z *= sigmoid_f(z);
          else if (a.act == 2) z = fmaxf(z, 0.f);
          o[j] = z;
        } else {
          float d = df[j];
          if (a.act == 1) {
            const float z = fmaf(xf[j], sc[j], sh[j]);
            const float g = sigmoid_f(z);
            d *= g * (1.0f + z * (1.0f - g));
          } else if (a.act == 2) {
            if (fmaf(xf[j], sc[j], sh[j]) <= 0.f) d = 0.f;
          }
          const float xh = (xf[j] - mu[j]) * rs[j];
          if (MODE == 2) {
            s0[j] += d;
            s1[j] = fmaf(d, xh, s1[j]);
          } else {
            o[j] = sc[j] * (d - k0[j] - xh * k1[j]);
          }
        }
      }
      if (MODE == 1 || MODE == 3) *reinterpret_cast<uint4*>(a.out + (size_t)r * C + cv * 8) = pack8(o);
    }
  }
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_acc[cv * 8 + j], s0[j]);
      atomicAdd(&s_acc[C + cv * 8 + j], s1[j]);
    }
    __syncthreads();
    for (int i = tid; i < 2 * C; i += nth) atomicAdd(a.sums_out + i, s_acc[i]);
  }
}

// -> true when the vector kernel ran
template <int MODE>
static bool launch_bn_vec(BnVecArgs a, cudaStream_t stream, int* status) {
  const int C = a.C;
  if ((C & 7) || C / 8 > 256) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.dy) | reinterpret_cast<uintptr_t>(a.out)) & 15) return false;
  const int cvn = C / 8;
  int ry = 256 / cvn;
  if (ry > 16) ry = 16;
  // ~2 CTAs per SM, at least 4 rows per thread
  int64_t rpb = ceil_div(a.R, (int64_t)2 * kNumSMs);
  if (rpb < 4 * ry) rpb = 4 * ry;
  a.rows_per_block = (int)rpb;
  const unsigned grid = (unsigned)ceil_div(a.R, rpb);
  launch_pdl(bn_vec_kernel<MODE>, dim3(grid), dim3(cvn, ry), (MODE == 0 || MODE == 2) ? sizeof(float) * 2 * C : 0, stream, a);
  count_launch();
  *status = check_launch();
  return true;
}

// Per-channel bookkeeping of a training-mode BatchNorm in one launch (C threads) instead of ~14 elementwise launches:
//   mean = sums[0] / count, var = max(sums[1] / count - mean^2, 0), rstd = rsqrt(var + eps), scale = gamma rstd, shift = beta - mean scale,
//   moving = momentum * moving + (1 - momentum) * batch statistic        (keras BatchNormalization, momentum 0.99)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ out, int C, float count,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (count == kBnCountOnDevice) count = sums[2 * C];
  const float mean = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float scale = gamma[c] * rstd;
  out[c] = scale;                      // scale
  out[C + c] = beta[c] - mean * scale;  // shift
  out[2 * C + c] = mean;
  out[3 * C + c] = rstd;
  if (moving_mean) moving_mean[c] = momentum * moving_mean[c] + (1.0f - momentum) * mean;
  if (moving_var) moving_var[c] = momentum * moving_var[c] + (1.0f - momentum) * var;
}
// dgamma += sums[1], dbeta += sums[0]   (local contribution; taken before the sync-BN all-reduce of the backward sums)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dgamma[c] += sums[C + c];
  dbeta[c] += sums[c];
}

static inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

static int rows_block_cfg(int C, dim3& block, int& rows_per_block, int64_t R, unsigned& grid) {
  if (C < 2 || (C & 1) || C / 2 > 1024) return TFASR_INVALID_VALUE;
  int ry = 1024 / (C / 2);
  if (ry > 16) ry = 16;
  if (ry < 1) ry = 1;
  block = dim3(C / 2, ry);
  int64_t target_blocks = kNumSMs * 4;
  int64_t rpb = ceil_div(R, target_blocks);
  if (rpb < ry) rpb = ry;
  rows_per_block = (int)rpb;
  grid = (unsigned)ceil_div(R, rpb);
  return TFASR_OK;
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_conv1_3x3s2_fwd(const float* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t F, int32_t C,
                                     tfasr_stream_t stream) {
  if (!x || !w || !bias || !y || B < 0 || T < 1 || F < 1) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  dim3 block;
  int ppb;
  unsigned grid;
  int st = rows_block_cfg(C, block, ppb, (int64_t)B * T1 * F1, grid);
  if (st) return st;
  conv1_fwd_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, static_cast<__nv_bfloat16*>(y), B, T, F, T1, F1, C, ppb);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_conv1_3x3s2_wgrad(const float* x, const void* dy, float* dw, int32_t B, int32_t T, int32_t F, int32_t C, tfasr_stream_t stream) {
  if (!x || !dy || !dw || B < 0 || T < 1 || F < 1) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  dim3 block;
  int ppb;
  unsigned grid;
  int st = rows_block_cfg(C, block, ppb, (int64_t)B * T1 * F1, grid);
  if (st) return st;
  conv1_wgrad_kernel<<<grid, block, sizeof(float) * 9 * C, static_cast<cudaStream_t>(stream)>>>(x, static_cast<const __nv_bfloat16*>(dy), dw, B, T, F, T1,
                                                                                                 F1, C, ppb);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_im2col_3x3s2(const void* a, void* col, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream) {
  if (!a || !col || B < 0 || T1 < 1 || F1 < 1 || C < 8 || (C & 7)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const int64_t total = (int64_t)B * T2 * F2 * 9 * (C >> 3);
  unsigned grid = (unsigned)min64(ceil_div(total, 256), kNumSMs * 32);
  im2col_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(col), B, T1, F1, T2,
                                                                      F2, C);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_col2im_3x3s2(const void* dcol, void* da, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream) {
  if (!dcol || !da || B < 0 || T1 < 1 || F1 < 1 || C < 8 || (C & 7)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const int64_t total = (int64_t)B * T1 * F1 * (C >> 3);
  unsigned grid = (unsigned)min64(ceil_div(total, 256), kNumSMs * 32);
  col2im_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dcol), static_cast<__nv_bfloat16*>(da), B, T1, F1,
                                                                      T2, F2, C);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_stats(const void* x, float* sums, int64_t R, int32_t C, tfasr_stream_t stream) {
  if (!x || !sums || R < 0) return TFASR_INVALID_VALUE;
  if (R == 0) return TFASR_OK;
  {
    BnVecArgs a{};
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.sums_out = sums;
    a.R = R;
    a.C = C;
    int vst = 0;
    if (launch_bn_vec<0>(a, static_cast<cudaStream_t>(stream), &vst)) return vst;
  }
  dim3 block;
  int rpb;
  unsigned grid;
  int st = rows_block_cfg(C, block, rpb, R, grid);
  if (st) return st;
  bn_stats_kernel<<<grid, block, sizeof(float) * 2 * C, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), sums, R, C, rpb);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_apply(const void* x, const float* scale, const float* shift, void* y, int64_t R, int32_t C, int32_t act,
                              tfasr_stream_t stream) {
  if (!x || !scale || !shift || !y || R < 0 || C < 2 || (C & 1)) return TFASR_INVALID_VALUE;
  if (R == 0) return TFASR_OK;
  {
    BnVecArgs a{};
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.out = static_cast<__nv_bfloat16*>(y);
    a.scale = scale;
    a.shift = shift;
    a.R = R;
    a.C = C;
    a.act = act;
    int vst = 0;
    if (launch_bn_vec<1>(a, static_cast<cudaStream_t>(stream), &vst)) return vst;
  }
  const int64_t total = R * (C >> 1);
  unsigned grid = (unsigned)min64(ceil_div(total, 256), kNumSMs * 32);
  bn_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), scale, shift, static_cast<__nv_bfloat16*>(y), R,
                                                                        C, act);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_bwd_stats(const void* dy, const void* x, const float* scale, const float* shift, const float* mean, const float* rstd,
                                  float* sums, int64_t R, int32_t C, int32_t act, tfasr_stream_t stream) {
  if (!dy || !x || !scale || !shift || !mean || !rstd || !sums || R < 0) return TFASR_INVALID_VALUE;
  if (R == 0) return TFASR_OK;
  {
    BnVecArgs a{};
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.dy = static_cast<const __nv_bfloat16*>(dy);
    a.scale = scale;
    a.shift = shift;
    a.mean = mean;
    a.rstd = rstd;
    a.sums_out = sums;
    a.R = R;
    a.C = C;
    a.act = act;
    int vst = 0;
    if (launch_bn_vec<2>(a, static_cast<cudaStream_t>(stream), &vst)) return vst;
  }
  dim3 block;
  int rpb;
  unsigned grid;
  int st = rows_block_cfg(C, block, rpb, R, grid);
  if (st) return st;
  bn_bwd_stats_kernel<<<grid, block, sizeof(float) * 2 * C, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), scale, shift, mean, rstd, sums, R, C, act, rpb);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_bwd_apply(const void* dy, const void* x, const float* scale, const float* shift, const float* mean, const float* rstd,
                                  const float* sums, float count, void* dx, int64_t R, int32_t C, int32_t act, tfasr_stream_t stream) {
  if (!dy || !x || !scale || !shift || !mean || !rstd || !sums || !dx || R < 0 || C < 2 || (C & 1)) return TFASR_INVALID_VALUE;
  if (R == 0) return TFASR_OK;
  {
    BnVecArgs a{};
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.dy = static_cast<const __nv_bfloat16*>(dy);
    a.out = static_cast<__nv_bfloat16*>(dx);
    a.scale = scale;
    a.shift = shift;
    a.mean = mean;
    a.rstd = rstd;
    a.sums_in = sums;
    a.inv_count = count > 0.f ? 1.0f / count : 0.f;
    a.count_on_device = (count == kBnCountOnDevice) ? 1 : 0;
    a.R = R;
    a.C = C;
    a.act = act;
    int vst = 0;
    if (launch_bn_vec<3>(a, static_cast<cudaStream_t>(stream), &vst)) return vst;
  }
  const int64_t total = R * (C >> 1);
  unsigned grid = (unsigned)min64(ceil_div(total, 256), kNumSMs * 32);
  bn_bwd_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), scale,
                                                                            shift, mean, rstd, sums, count, static_cast<__nv_bfloat16*>(dx), R, C, act);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_finalize(const float* sums, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* out,
                                 int32_t C, float count, float eps, float momentum, tfasr_stream_t stream) {
  if (!sums || !gamma || !beta || !out || C < 1 || !(count > 0.f || count == kBnCountOnDevice)) return TFASR_INVALID_VALUE;
  bn_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, gamma, beta, moving_mean, moving_var, out, C, count,
                                                                                                  eps, momentum);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_bn_bwd_finalize(const float* sums, float* dgamma, float* dbeta, int32_t C, tfasr_stream_t stream) {
  if (!sums || !dgamma || !dbeta || C < 1) return TFASR_INVALID_VALUE;
  bn_bwd_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, dgamma, dbeta, C);
  count_launch();
  return check_launch();
}
