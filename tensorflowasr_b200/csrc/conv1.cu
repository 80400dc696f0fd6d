// First conv of Conv2dSubsampling (Cin = 1, 3x3, stride 2, causal pad 2/2; subsampling.py:218-230 + convolution.py:25-37,132-144),
// second generation: x rows staged once in SMEM (zero-padded, so the taps need no bounds checks), thread = 8 output channels with its 72
// tap weights in registers, 16-byte stores / loads, no 64-bit index arithmetic in the loop.  The first-generation kernels in conv_bn.cu
// (thread = channel pair, 9 global loads and a 4-byte store per position) run at ~7 % of the HBM roofline of their 184 MB of output
// (0.43 ms forward, 0.42 ms weight gradient per step) and stay the default until these have been measured on a B200:
// opt-in with ops.set_conv1_v2(True) / TFASR_CONV1_V2=1.
//   fwd  : y[b,t1,f1,c] = bias[c] + sum_{kh,kw} w[kh,kw,0,c] * x[b, 2 t1 + kh - 2, 2 f1 + kw - 2]     HBM: 4 B T F + 2 B T1 F1 C bytes
//   wgrad: dw[kh,kw,0,c] += sum_{b,t1,f1} dy[b,t1,f1,c] * x[b, 2 t1 + kh - 2, 2 f1 + kw - 2]          HBM: 4 B T F + 2 B T1 F1 C bytes
// Plain SIMT: builds twice (nvcc -> libtfasr_b200.so; g++ -DTFASR_EMU_BUILD -> tests/emu).
#include "simt.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

constexpr int kC1PY = 16;  // position lanes per CTA

// stage rows t = 2*t1_0 - 2 .. 2*t1_0 - 2 + nrows - 1 of utterance b, columns f = -2 .. F-1 (row pitch F + 2), zeros outside the signal
__device__ __forceinline__ void conv1_stage_rows(const float* __restrict__ xb, float* sx, int T, int F, int t_first, int nrows, int tid, int nth) {
  const int W = F + 2;
  for (int i = tid; i < nrows * W; i += nth) {
    const int r = i / W, j = i - r * W;
    const int t = t_first + r, f = j - 2;
    sx[i] = (t >= 0 && t < T && f >= 0) ? xb[(size_t)t * F + f] : 0.f;
  }
}

// grid (ceil(T1 / RT), B), block (C/8, kC1PY).  Dynamic SMEM: (2 RT + 1) (F + 2) floats.
template <int RT>
__global__ void __launch_bounds__(512) conv1_fwd_v2_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int T, int F, int T1, int F1, int C) {
  TFASR_DYN_SMEM(smem_raw);
  float* sx = reinterpret_cast<float*>(smem_raw);
  const int b = blockIdx.y, t1_0 = blockIdx.x * RT;
  const int c8 = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * blockDim.x + c8, nth = blockDim.x * blockDim.y;
  const int W = F + 2;
  conv1_stage_rows(x + (size_t)b * T * F, sx, T, F, 2 * t1_0 - 2, 2 * RT + 1, tid, nth);
  float wr[9][8], br[8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[k][e] = w[k * C + c8 * 8 + e];
#pragma unroll
  for (int e = 0; e < 8; ++e) br[e] = bias[c8 * 8 + e];
  __syncthreads();
  for (int pos = ty; pos < RT * F1; pos += kC1PY) {
    const int r1 = pos / F1, f1 = pos - r1 * F1;
    const int t1 = t1_0 + r1;
    if (t1 >= T1) break;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = br[e];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = sx[(2 * r1 + kh) * W + 2 * f1 + kw];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(v, wr[kh * 3 + kw][e], acc[e]);
      }
    st16(y + (((size_t)b * T1 + t1) * F1 + f1) * C + c8 * 8, bf8_pack(acc));
  }
}

// grid (ceil(T1 / RT), B), block (C/8, kC1PY).  Dynamic SMEM: (2 RT + 1) (F + 2) + 9 C floats.
template <int RT>
__global__ void __launch_bounds__(512) conv1_wgrad_v2_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                                                              int T, int F, int T1, int F1, int C) {
  TFASR_DYN_SMEM(smem_raw);
  float* sx = reinterpret_cast<float*>(smem_raw);
  const int W = F + 2;
  float* s_dw = sx + (2 * RT + 1) * W;  // [9][C]
  const int b = blockIdx.y, t1_0 = blockIdx.x * RT;
  const int c8 = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * blockDim.x + c8, nth = blockDim.x * blockDim.y;
  conv1_stage_rows(x + (size_t)b * T * F, sx, T, F, 2 * t1_0 - 2, 2 * RT + 1, tid, nth);
  for (int i = tid; i < 9 * C; i += nth) s_dw[i] = 0.f;
  float g[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) g[k][e] = 0.f;
  __syncthreads();
  for (int pos = ty; pos < RT * F1; pos += kC1PY) {
    const int r1 = pos / F1, f1 = pos - r1 * F1;
    const int t1 = t1_0 + r1;
    if (t1 >= T1) break;
    float d[8];
    bf8_unpack(ld16(dy + (((size_t)b * T1 + t1) * F1 + f1) * C + c8 * 8), d);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = sx[(2 * r1 + kh) * W + 2 * f1 + kw];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[kh * 3 + kw][e] = fmaf(v, d[e], g[kh * 3 + kw][e]);
      }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_dw[k * C + c8 * 8 + e], g[k][e]);
  __syncthreads();
  for (int i = tid; i < 9 * C; i += nth) atomicAdd(dw + i, s_dw[i]);
}

constexpr int kC1FwdRT = 4, kC1WgRT = 8;

}  // namespace tfasr

using namespace tfasr;

namespace {
bool conv1_args_ok(const void* x, const void* a, const void* b, int B, int T, int F, int C) {
  return x && a && b && B >= 0 && B <= 65535 && T >= 1 && F >= 1 && C >= 8 && !(C & 7) && (C / 8) * kC1PY <= 512 && !(reinterpret_cast<uintptr_t>(b) & 15);
}
}  // namespace

extern "C" int tfasr_conv1_3x3s2_fwd_v2(const float* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t F, int32_t C,
                                        tfasr_stream_t stream_) {
  if (!conv1_args_ok(x, w, y, B, T, F, C) || !bias) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  const dim3 grid((unsigned)ceil_div(T1, kC1FwdRT), (unsigned)B), block((unsigned)(C / 8), kC1PY);
  const size_t smem = sizeof(float) * (size_t)(2 * kC1FwdRT + 1) * (F + 2);
  if (smem > 48 * 1024) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, smem, [&] { conv1_fwd_v2_kernel<kC1FwdRT>(x, w, bias, static_cast<__nv_bfloat16*>(y), T, F, T1, F1, C); });
  return TFASR_OK;
#else
  conv1_fwd_v2_kernel<kC1FwdRT><<<grid, block, smem, static_cast<cudaStream_t>(stream_)>>>(x, w, bias, static_cast<__nv_bfloat16*>(y), T, F, T1, F1, C);
  count_launch();
  return check_launch();
#endif
}

extern "C" int tfasr_conv1_3x3s2_wgrad_v2(const float* x, const void* dy, float* dw, int32_t B, int32_t T, int32_t F, int32_t C, tfasr_stream_t stream_) {
  if (!conv1_args_ok(x, dw, dy, B, T, F, C)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T1 = (T + 1) / 2, F1 = (F + 1) / 2;
  const dim3 grid((unsigned)ceil_div(T1, kC1WgRT), (unsigned)B), block((unsigned)(C / 8), kC1PY);
  const size_t smem = sizeof(float) * ((size_t)(2 * kC1WgRT + 1) * (F + 2) + (size_t)9 * C);
  if (smem > 48 * 1024) return TFASR_INVALID_VALUE;
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, smem, [&] { conv1_wgrad_v2_kernel<kC1WgRT>(x, static_cast<const __nv_bfloat16*>(dy), dw, T, F, T1, F1, C); });
  return TFASR_OK;
#else
  conv1_wgrad_v2_kernel<kC1WgRT><<<grid, block, smem, static_cast<cudaStream_t>(stream_)>>>(x, static_cast<const __nv_bfloat16*>(dy), dw, T, F, T1, F1, C);
  count_launch();
  return check_launch();
#endif
}
