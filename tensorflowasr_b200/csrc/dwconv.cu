// Causal depthwise 1-D convolution (ConvModule's DepthwiseConv1D, K = 31) for sm_100a: forward, data gradient, weight gradient.
//
// Replaces keras DepthwiseConv1D with the reference's "causal" padding (tensorflow_asr/models/layers/convolution.py:159-216:
// pad K-1 frames on the left, then VALID), which the reference runs on TF's DepthwiseConv2dNative kernels
// (models/encoders/conformer.py:311-323,370).  Channels-last bf16 [B, T, C]; every CTA stages a (TT + K - 1) x C window in SMEM so each
// input element is read from HBM once.  HBM-bound: 2 * B*T*C * 2 bytes per pass.
#include "common.cuh"

namespace tfasr {

constexpr int kDwTT = 64;  // output frames per CTA

// y[b,t,c] = bias[c] + sum_k w[k,c] * x[b, t - (K-1) + k, c]           block (C/2, 4); grid (ceil(T/TT), B)
__global__ void __launch_bounds__(1024) dwconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ y, int T, int C, int K) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  __nv_bfloat162* sx = reinterpret_cast<__nv_bfloat162*>(dw_smem);  // [(TT + K - 1)][C/2]
  const int C2 = C >> 1;
  float* sw = reinterpret_cast<float*>(sx + (size_t)(kDwTT + K - 1) * C2);  // [K][C]
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(x + (size_t)b * T * C);
  for (int i = tid; i < (kDwTT + K - 1) * C2; i += nth) {
    const int r = i / C2, c2 = i % C2;
    const int t = t0 - (K - 1) + r;
    sx[i] = (t >= 0 && t < T) ? xb[(size_t)t * C2 + c2] : __floats2bfloat162_rn(0.f, 0.f);
  }
  for (int i = tid; i < K * C; i += nth) sw[i] = w[i];
  __syncthreads();
  const int c2 = threadIdx.x;
  const float b0 = bias ? bias[2 * c2] : 0.f, b1 = bias ? bias[2 * c2 + 1] : 0.f;
  for (int tt = threadIdx.y; tt < kDwTT; tt += blockDim.y) {
    const int t = t0 + tt;
    if (t >= T) break;
    float a0 = b0, a1 = b1;
    for (int k = 0; k < K; ++k) {
      const __nv_bfloat162 v = sx[(size_t)(tt + k) * C2 + c2];
      a0 = fmaf(__low2float(v), sw[k * C + 2 * c2], a0);
      a1 = fmaf(__high2float(v), sw[k * C + 2 * c2 + 1], a1);
    }
    reinterpret_cast<__nv_bfloat162*>(y + ((size_t)b * T + t) * C)[c2] = __floats2bfloat162_rn(a0, a1);
  }
}

// dx[b,t,c] = sum_k w[k,c] * dy[b, t + (K-1) - k, c]   (frames beyond T contribute nothing)
__global__ void __launch_bounds__(1024) dwconv_bwd_data_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ dx, int T, int C, int K) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  __nv_bfloat162* sd = reinterpret_cast<__nv_bfloat162*>(dw_smem);  // rows t0 .. t0 + TT + K - 2
  const int C2 = C >> 1;
  float* sw = reinterpret_cast<float*>(sd + (size_t)(kDwTT + K - 1) * C2);
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  const __nv_bfloat162* db = reinterpret_cast<const __nv_bfloat162*>(dy + (size_t)b * T * C);
  for (int i = tid; i < (kDwTT + K - 1) * C2; i += nth) {
    const int r = i / C2, c2 = i % C2;
    const int t = t0 + r;
    sd[i] = (t < T) ? db[(size_t)t * C2 + c2] : __floats2bfloat162_rn(0.f, 0.f);
  }
  for (int i = tid; i < K * C; i += nth) sw[i] = w[i];
  __syncthreads();
  const int c2 = threadIdx.x;
  for (int tt = threadIdx.y; tt < kDwTT; tt += blockDim.y) {
    const int t = t0 + tt;
    if (t >= T) break;
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < K; ++k) {
      const __nv_bfloat162 v = sd[(size_t)(tt + (K - 1) - k) * C2 + c2];
      a0 = fmaf(__low2float(v), sw[k * C + 2 * c2], a0);
      a1 = fmaf(__high2float(v), sw[k * C + 2 * c2 + 1], a1);
    }
    reinterpret_cast<__nv_bfloat162*>(dx + ((size_t)b * T + t) * C)[c2] = __floats2bfloat162_rn(a0, a1);
  }
}

// dw[k,c] += sum_{b,t} dy[b,t,c] * x[b, t - (K-1) + k, c];  dbias[c] += sum dy   (dbias may be NULL)
__global__ void __launch_bounds__(1024) dwconv_bwd_weight_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                 float* __restrict__ dw, float* __restrict__ dbias, int T, int C, int K) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  const int C2 = C >> 1;
  __nv_bfloat162* sx = reinterpret_cast<__nv_bfloat162*>(dw_smem);        // [(TT + K - 1)][C/2]
  __nv_bfloat162* sd = sx + (size_t)(kDwTT + K - 1) * C2;                  // [TT][C/2]
  float* sacc = reinterpret_cast<float*>(sd + (size_t)kDwTT * C2);         // [K + 1][C]
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(x + (size_t)b * T * C);
  const __nv_bfloat162* db = reinterpret_cast<const __nv_bfloat162*>(dy + (size_t)b * T * C);
  for (int i = tid; i < (kDwTT + K - 1) * C2; i += nth) {
    const int r = i / C2, c2 = i % C2;
    const int t = t0 - (K - 1) + r;
    sx[i] = (t >= 0 && t < T) ? xb[(size_t)t * C2 + c2] : __floats2bfloat162_rn(0.f, 0.f);
  }
  for (int i = tid; i < kDwTT * C2; i += nth) {
    const int r = i / C2, c2 = i % C2;
    const int t = t0 + r;
    sd[i] = (t < T) ? db[(size_t)t * C2 + c2] : __floats2bfloat162_rn(0.f, 0.f);
  }
  for (int i = tid; i < (K + 1) * C; i += nth) sacc[i] = 0.f;
  __syncthreads();
  // thread (c2, ky) owns taps k = ky, ky + blockDim.y, ... of channel pair c2
  const int c2 = threadIdx.x;
  for (int k = threadIdx.y; k < K + 1; k += blockDim.y) {
    float a0 = 0.f, a1 = 0.f;
    if (k < K) {
      for (int tt = 0; tt < kDwTT; ++tt) {
        const __nv_bfloat162 d = sd[(size_t)tt * C2 + c2];
        const __nv_bfloat162 v = sx[(size_t)(tt + k) * C2 + c2];
        a0 = fmaf(__low2float(d), __low2float(v), a0);
        a1 = fmaf(__high2float(d), __high2float(v), a1);
      }
    } else {
      for (int tt = 0; tt < kDwTT; ++tt) {
        const __nv_bfloat162 d = sd[(size_t)tt * C2 + c2];
        a0 += __low2float(d);
        a1 += __high2float(d);
      }
    }
    sacc[k * C + 2 * c2] = a0;
    sacc[k * C + 2 * c2 + 1] = a1;
  }
  __syncthreads();
  for (int i = tid; i < K * C; i += nth) atomicAdd(dw + i, sacc[i]);
  if (dbias)
    for (int i = tid; i < C; i += nth) atomicAdd(dbias + i, sacc[K * C + i]);
}

// ---------------------------------------------------------------------------------------------- register-blocked forward / data gradient
// One thread = one channel pair x 8 consecutive frames: the K + 7 inputs slide through registers (each SMEM value is read once and
// used by up to 8 outputs) and the tap weights are fetched once per thread.  3x fewer instructions per output than the per-frame
// kernels above, which stay as the generic fallback.
//   REVERSE = false: y[t]  = bias + sum_k w[k] x[t - (K-1) + k]      (window starts K-1 frames before the tile)
//   REVERSE = true : dx[t] =        sum_k w[K-1-k] dy[t + k]         (window starts at the tile; frames >= T are zero)
template <int K, bool REVERSE>
__global__ void __launch_bounds__(640) dwconv_sliding_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                                                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int T, int C) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  constexpr int TPT = 8;  // frames per thread; kDwTT / TPT row groups per CTA
  extern __shared__ __align__(16) uint8_t dw_smem[];
  __nv_bfloat162* sx = reinterpret_cast<__nv_bfloat162*>(dw_smem);  // [(TT + K - 1)][C/2]
  const int C2 = C >> 1;
  const int b = blockIdx.y, t0 = blockIdx.x * kDwTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(in + (size_t)b * T * C);
  const int tstart = REVERSE ? t0 : t0 - (K - 1);
  for (int r = threadIdx.y; r < kDwTT + K - 1; r += blockDim.y) {
    const int t = tstart + r;
    sx[(size_t)r * C2 + threadIdx.x] = (t >= 0 && t < T) ? xb[(size_t)t * C2 + threadIdx.x] : __floats2bfloat162_rn(0.f, 0.f);
  }
  (void)tid;
  (void)nth;
  const int c2 = threadIdx.x;
  float w0[K], w1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int kk = REVERSE ? (K - 1 - k) : k;
    w0[k] = w[kk * C + 2 * c2];
    w1[k] = w[kk * C + 2 * c2 + 1];
  }
  const float b0 = (!REVERSE && bias) ? bias[2 * c2] : 0.f, b1 = (!REVERSE && bias) ? bias[2 * c2 + 1] : 0.f;
  __syncthreads();
  for (int grp = threadIdx.y; grp < kDwTT / TPT; grp += blockDim.y) {
    const int tt0 = grp * TPT;
    if (t0 + tt0 >= T) break;
    float a0[TPT], a1[TPT];
#pragma unroll
    for (int o = 0; o < TPT; ++o) {
      a0[o] = b0;
      a1[o] = b1;
    }
    // out[tt0 + o] = sum_k w[k] * win[tt0 + o + k]: walk the window once, each value feeds the (o, k) pairs with o + k = i
#pragma unroll
    for (int i = 0; i < TPT + K - 1; ++i) {
      const float2 v = __bfloat1622float2(sx[(size_t)(tt0 + i) * C2 + c2]);
#pragma unroll
      for (int o = 0; o < TPT; ++o) {
        const int k = i - o;
        if (k >= 0 && k < K) {
          a0[o] = fmaf(v.x, w0[k], a0[o]);
          a1[o] = fmaf(v.y, w1[k], a1[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < TPT; ++o) {
      const int t = t0 + tt0 + o;
      if (t < T) reinterpret_cast<__nv_bfloat162*>(out + ((size_t)b * T + t) * C)[c2] = __floats2bfloat162_rn(a0[o], a1[o]);
    }
  }
}

static int dw_cfg(int C, int K, dim3& block, size_t& smem_fwd, size_t& smem_w) {
  if (C < 2 || (C & 1) || C / 2 > 512 || K < 1 || K > 128) return TFASR_INVALID_VALUE;
  int ty = 1024 / (C / 2);
  if (ty > 8) ty = 8;
  block = dim3(C / 2, ty);
  smem_fwd = (size_t)(kDwTT + K - 1) * C * 2 + sizeof(float) * K * C;
  smem_w = (size_t)(kDwTT + K - 1) * C * 2 + (size_t)kDwTT * C * 2 + sizeof(float) * (K + 1) * C;
  if (smem_fwd > 200 * 1024 || smem_w > 200 * 1024) return TFASR_INVALID_VALUE;
  return TFASR_OK;
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_dwconv_causal_fwd(const void* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t C, int32_t K,
                                       tfasr_stream_t stream) {
  dim3 block;
  size_t sf, sw;
  if (!x || !w || !y || B < 0 || T < 1 || B > 65535) return TFASR_INVALID_VALUE;
  int st = dw_cfg(C, K, block, sf, sw);
  if (st) return st;
  if (B == 0) return TFASR_OK;
  if (K == 31 && block.x * block.y <= 640) {  // the Conformer kernel size: register-blocked variant (62 tap weights in registers)
    const size_t smem = (size_t)(kDwTT + K - 1) * C * 2;
    if (cudaFuncSetAttribute(dwconv_sliding_kernel<31, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    launch_pdl(dwconv_sliding_kernel<31, false>, dim3((unsigned)ceil_div(T, kDwTT), B), dim3(block), smem, static_cast<cudaStream_t>(stream),
        static_cast<const __nv_bfloat16*>(x), w, bias, static_cast<__nv_bfloat16*>(y), T, C);
    count_launch();
    return check_launch();
  }
  if (cudaFuncSetAttribute(dwconv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  dwconv_fwd_kernel<<<dim3((unsigned)ceil_div(T, kDwTT), B), block, sf, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), w, bias, static_cast<__nv_bfloat16*>(y), T, C, K);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_dwconv_causal_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int32_t B, int32_t T,
                                       int32_t C, int32_t K, tfasr_stream_t stream_) {
  dim3 block;
  size_t sf, sw;
  if (!dy || !x || !w || !dx || !dw || B < 0 || T < 1 || B > 65535) return TFASR_INVALID_VALUE;
  int st = dw_cfg(C, K, block, sf, sw);
  if (st) return st;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cudaFuncSetAttribute(dwconv_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  if (cudaFuncSetAttribute(dwconv_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  dim3 grid((unsigned)ceil_div(T, kDwTT), B);
  if (K == 31 && block.x * block.y <= 640) {
    const size_t smem = (size_t)(kDwTT + K - 1) * C * 2;
    if (cudaFuncSetAttribute(dwconv_sliding_kernel<31, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    launch_pdl(dwconv_sliding_kernel<31, true>, dim3(grid), dim3(block), smem, stream, static_cast<const __nv_bfloat16*>(dy), w, nullptr, static_cast<__nv_bfloat16*>(dx), T, C);
  } else {
    dwconv_bwd_data_kernel<<<grid, block, sf, stream>>>(static_cast<const __nv_bfloat16*>(dy), w, static_cast<__nv_bfloat16*>(dx), T, C, K);
  }
  count_launch();
  dwconv_bwd_weight_kernel<<<grid, block, sw, stream>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw, dbias, T, C, K);
  count_launch();
  return check_launch();
}

// Data gradient alone (the weight / bias gradient then comes from tfasr_dwconv_causal_wgrad_v2, dwconv_wgrad_v2.cu).
extern "C" int tfasr_dwconv_causal_bwd_dx(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t K, tfasr_stream_t stream_) {
  dim3 block;
  size_t sf, sw;
  if (!dy || !w || !dx || B < 0 || T < 1 || B > 65535) return TFASR_INVALID_VALUE;
  int st = dw_cfg(C, K, block, sf, sw);
  if (st) return st;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid((unsigned)ceil_div(T, kDwTT), B);
  if (K == 31 && block.x * block.y <= 640) {
    const size_t smem = (size_t)(kDwTT + K - 1) * C * 2;
    if (cudaFuncSetAttribute(dwconv_sliding_kernel<31, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    launch_pdl(dwconv_sliding_kernel<31, true>, dim3(grid), dim3(block), smem, stream, static_cast<const __nv_bfloat16*>(dy), w, nullptr, static_cast<__nv_bfloat16*>(dx), T, C);
  } else {
    if (cudaFuncSetAttribute(dwconv_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED;
    dwconv_bwd_data_kernel<<<grid, block, sf, stream>>>(static_cast<const __nv_bfloat16*>(dy), w, static_cast<__nv_bfloat16*>(dx), T, C, K);
  }
  count_launch();
  return check_launch();
}
