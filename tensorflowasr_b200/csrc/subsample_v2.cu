// im2col / col2im of the second subsampling conv (3x3, stride 2, causal pad 2/2; subsampling.py:218-230), second generation
// (opt-in: ops.set_im2col_v2 / TFASR_IM2COL_V2=1).  Same contracts as tfasr_im2col_3x3s2 / tfasr_col2im_3x3s2 (conv_bn.cu):
//   im2col: a [B,T1,F1,C] bf16 -> col [B*T2*F2, 9C] bf16, column (kh*3 + kw)*C + c, zeros outside the input
//   col2im: dcol -> da [B,T1,F1,C]: every input element gathers the <= 4 taps that read it (no atomics)
// Why: the first-generation kernels index one flat 64-bit counter per 16-byte chunk (three 64-bit divisions each) and move their
// 0.6 GB at 2.7 TB/s (41 % of the HBM copy peak; 0.22 ms each, three launches per step).  Here one CTA owns one output row (b, t2) resp.
// input row (b, t): thread = (8-channel vector, position lane), 32-bit index arithmetic with a compile-time divisor, 16-byte accesses.
// HBM-bound byte work: 2 B T1 F1 C + 2 * 9 B T2 F2 C bytes per pass.  Plain SIMT: builds twice (see simt.cuh).
#include "simt.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

// grid (T2, B), block (C/8, PY): ty walks the (f2, tap) pairs of the row, tap fastest: consecutive threads write consecutive 16-byte chunks
__global__ void __launch_bounds__(512) im2col_v2_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ col, int T1, int F1, int T2,
                                                        int F2, int C) {
  const int t2 = blockIdx.x, b = blockIdx.y;
  const int cv = threadIdx.x;
  const __nv_bfloat16* ab = a + (size_t)b * T1 * F1 * C + cv * 8;
  __nv_bfloat16* cb = col + ((size_t)b * T2 + t2) * F2 * 9 * C + cv * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int pair = threadIdx.y; pair < 9 * F2; pair += blockDim.y) {
    const int f2 = pair / 9, k = pair - 9 * f2;
    const int kh = k / 3, kw = k - 3 * kh;
    const int t = 2 * t2 + kh - 2, f = 2 * f2 + kw - 2;
    uint4 v = zero;
    if (t >= 0 && t < T1 && f >= 0 && f < F1) v = ld16(ab + ((size_t)t * F1 + f) * C);
    st16(cb + (size_t)pair * C, v);
  }
}

// grid (T1, B), block (C/8, PY): thread (cv, f lane) gathers the <= 4 taps of input element (t, f)
__global__ void __launch_bounds__(512) col2im_v2_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ da, int T1, int F1, int T2,
                                                        int F2, int C) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int cv = threadIdx.x;
  for (int f = threadIdx.y; f < F1; f += blockDim.y) {
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
        float v[8];
        bf8_unpack(ld16(dcol + (((size_t)b * T2 + t2) * F2 + f2) * 9 * C + (size_t)(kh * 3 + kw) * C + cv * 8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    st16(da + (((size_t)b * T1 + t) * F1 + f) * C + cv * 8, bf8_pack(acc));
  }
}

}  // namespace tfasr

using namespace tfasr;

namespace {
bool sub_args_ok(const void* x, const void* y, int B, int T1, int F1, int C) {
  return x && y && B >= 0 && B <= 65535 && T1 >= 1 && F1 >= 1 && C >= 8 && !(C & 7) && C / 8 <= 512 &&
         !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15);
}
dim3 sub_block(int C, int rows) {
  int py = 512 / (C / 8);
  if (py > rows) py = rows;
  if (py < 1) py = 1;
  return dim3((unsigned)(C / 8), (unsigned)py);
}
}  // namespace

extern "C" int tfasr_im2col_3x3s2_v2(const void* a, void* col, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream_) {
  if (!sub_args_ok(a, col, B, T1, F1, C)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const dim3 grid((unsigned)T2, (unsigned)B), block = sub_block(C, 9 * F2);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, 0, [&] { im2col_v2_kernel(static_cast<const __nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(col), T1, F1, T2, F2, C); });
  return TFASR_OK;
#else
  im2col_v2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(col), T1, F1, T2,
                                                                          F2, C);
  count_launch();
  return check_launch();
#endif
}

extern "C" int tfasr_col2im_3x3s2_v2(const void* dcol, void* da, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream_) {
  if (!sub_args_ok(dcol, da, B, T1, F1, C)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int T2 = (T1 + 1) / 2, F2 = (F1 + 1) / 2;
  const dim3 grid((unsigned)T1, (unsigned)B), block = sub_block(C, F1);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, 0, [&] { col2im_v2_kernel(static_cast<const __nv_bfloat16*>(dcol), static_cast<__nv_bfloat16*>(da), T1, F1, T2, F2, C); });
  return TFASR_OK;
#else
  col2im_v2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(dcol), static_cast<__nv_bfloat16*>(da), T1, F1, T2,
                                                                          F2, C);
  count_launch();
  return check_launch();
#endif
}
