// Weight (and bias) gradient of the ConvModule's causal depthwise conv (convolution.py:159-216, conformer.py:311-323), second generation
// (opt-in: ops.set_dwconv_wgrad_v2 / TFASR_DWCONV_WGRAD_V2=1).  Same contract as the weight half of tfasr_dwconv_causal_bwd (dwconv.cu):
//   dw[k,c] += sum_{b,t} dy[b,t,c] * x[b, t - (K-1) + k, c]        dbias[c] += sum_{b,t} dy[b,t,c]   (dbias may be NULL)
// Why: the first-generation kernel stages (64 + K - 1) frames per CTA in SMEM and gives each thread up to 4 taps x 64 frames of serial
// work behind two barriers: 21 us per layer on a 2.3 MB tensor (16 launches, 0.34 ms of the step).  The whole working set lives in L2/L1,
// so here every thread simply owns one (tap, channel pair) and walks the frames of its utterance slice straight from global memory —
// loads coalesced over the channel pairs, re-reads of x across taps served by L1 — and ends with two atomics.
// Latency / L1-bound byte work (2 * B*T*C * 2 bytes from HBM at most); no tensor cores.  Plain SIMT: builds twice (see simt.cuh).
#include "simt.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

// grid (ceil((K + has_bias) / KY), B, TS), block (C/2, KY): thread (c2, k).  k == K: the bias row.
__global__ void __launch_bounds__(1024) dwconv_wgrad_v2_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                               float* __restrict__ dw, float* __restrict__ dbias, int T, int C, int K, int frames) {
  const int C2 = C >> 1;
  const int c2 = threadIdx.x;
  const int k = blockIdx.x * blockDim.y + threadIdx.y;
  const int b = blockIdx.y;
  if (k > K || (k == K && !dbias)) return;
  const int t_lo = blockIdx.z * frames, t_hi = min(T, t_lo + frames);
  const __nv_bfloat162* dy2 = reinterpret_cast<const __nv_bfloat162*>(dy) + (size_t)b * T * C2 + c2;
  float a0 = 0.f, a1 = 0.f;
  if (k == K) {
    for (int t = t_lo; t < t_hi; ++t) {
      const float2 d = __bfloat1622float2(dy2[(size_t)t * C2]);
      a0 += d.x;
      a1 += d.y;
    }
    atomicAdd(dbias + 2 * c2, a0);
    atomicAdd(dbias + 2 * c2 + 1, a1);
    return;
  }
  const int shift = (K - 1) - k;  // x frame = t - shift
  const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(x) + (size_t)b * T * C2 + c2;
  int t = max(t_lo, shift);
  for (; t + 4 <= t_hi; t += 4) {  // four independent load pairs in flight
    float2 d[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      d[i] = __bfloat1622float2(dy2[(size_t)(t + i) * C2]);
      v[i] = __bfloat1622float2(x2[(size_t)(t + i - shift) * C2]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a0 = fmaf(d[i].x, v[i].x, a0);
      a1 = fmaf(d[i].y, v[i].y, a1);
    }
  }
  for (; t < t_hi; ++t) {
    const float2 d = __bfloat1622float2(dy2[(size_t)t * C2]);
    const float2 v = __bfloat1622float2(x2[(size_t)(t - shift) * C2]);
    a0 = fmaf(d.x, v.x, a0);
    a1 = fmaf(d.y, v.y, a1);
  }
  atomicAdd(dw + (size_t)k * C + 2 * c2, a0);
  atomicAdd(dw + (size_t)k * C + 2 * c2 + 1, a1);
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_dwconv_causal_wgrad_v2(const void* dy, const void* x, float* dw, float* dbias, int32_t B, int32_t T, int32_t C, int32_t K,
                                            tfasr_stream_t stream_) {
  if (!dy || !x || !dw || B < 0 || B > 65535 || T < 1 || C < 2 || (C & 1) || C / 2 > 512 || K < 1 || K > 1024 ||
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 3))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int C2 = C / 2;
  int ky = 1024 / C2;
  if (ky > 8) ky = 8;
  const int rows = K + (dbias ? 1 : 0);
  const int gx = (int)ceil_div(rows, ky);
  // enough CTAs for ~2 per SM: split the frames when the batch is small
  int ts = (int)ceil_div((int64_t)2 * kNumSMs, (int64_t)gx * B);
  if (ts < 1) ts = 1;
  if (ts > 8) ts = 8;
  int frames = (int)ceil_div(T, ts);
  if (frames < 32) frames = T < 32 ? T : 32;
  ts = (int)ceil_div(T, frames);
  const dim3 grid((unsigned)gx, (unsigned)B, (unsigned)ts), block((unsigned)C2, (unsigned)ky);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, 0, [&] { dwconv_wgrad_v2_kernel(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw, dbias, T, C, K, frames); });
  return TFASR_OK;
#else
  dwconv_wgrad_v2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw,
                                                                                dbias, T, C, K, frames);
  count_launch();
  return check_launch();
#endif
}
