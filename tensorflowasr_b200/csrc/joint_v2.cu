// Transducer joint, elementwise / reduction halves, second generation (opt-in: ops.set_joint_v2 / TFASR_JOINT_V2=1).  Same contracts as
// joint.cu (base_transducer.py:199-207,291 and its backward):
//   joint_tanh_fwd_v2    h[b,t,u,:] = tanh(e[b,t,:] + p[b,u,:])           bf16 [B,T,U1,J]
//   joint_reduce_bwd_v2  de[b,t,:] = sum_u dz[b,t,u,:],  dp[b,u,:] += sum_t dz[b,t,u,:]     (dp zeroed by the caller, as before)
// Why: the first-generation kernels move the 333 MB lattice at 2.1-2.3 TB/s (32-35 % of the HBM copy peak): the forward writes 4 bytes per
// thread with a modulo in the loop; the backward reduces de through SMEM with two CTA barriers per frame and runs 160 CTAs of 10 warps.
// Here: thread = 8 joint columns (16-byte accesses) with the column vector fixed per thread (e / partial sums stay in registers);
// in the backward the 8 label lanes of a column vector sit in ONE warp, so de is a 3-step shuffle reduction and the frame loop has no barrier.
// HBM-bound byte work: 2 B per lattice node and joint column, each pass.  Plain SIMT: builds twice (see simt.cuh).
#include "simt.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

__device__ __forceinline__ float jv2_tanh(float x) {
#ifdef TFASR_EMU_BUILD
  return tanhf(x);
#else
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));  // same MUFU path as joint.cu: the bf16 rounding of h hides its 2^-11 error
  return y;
#endif
}

// grid (T, B), block (J/8, UL): thread (cv, ul) owns column vector cv for labels u = ul, ul + UL, ...
__global__ void __launch_bounds__(512) joint_tanh_fwd_v2_kernel(const __nv_bfloat16* __restrict__ e, const __nv_bfloat16* __restrict__ p,
                                                                __nv_bfloat16* __restrict__ h, int T, int U1, int J) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int cv = threadIdx.x;
  float ev[8];
  bf8_unpack(ld16(e + ((size_t)b * T + t) * J + cv * 8), ev);
  const __nv_bfloat16* pb = p + (size_t)b * U1 * J + cv * 8;
  __nv_bfloat16* hb = h + (((size_t)b * T + t) * U1) * J + cv * 8;
  for (int u = threadIdx.y; u < U1; u += blockDim.y) {
    float pv[8];
    bf8_unpack(ld16(pb + (size_t)u * J), pv);
#pragma unroll
    for (int c = 0; c < 8; ++c) pv[c] = jv2_tanh(ev[c] + pv[c]);
    st16(hb + (size_t)u * J, bf8_pack(pv));
  }
}

// grid (chunks, B), block 32 * W: lane = (cvl = lane / 8 in 0..3, uw = lane % 8); warp w covers column vectors cv = 4 w' + cvl for
// w' = w, w + W, ...  Per frame a thread loads NU = ceil(U1 / 8) vectors (labels uw, uw + 8, ...), adds them into its dp partials and
// into a per-frame sum that is reduced over the 8 uw lanes by shuffles.
constexpr int kJv2NU = 9;  // labels per thread: U1 <= 72
__global__ void __launch_bounds__(256) joint_reduce_bwd_v2_kernel(const __nv_bfloat16* __restrict__ dz, float* __restrict__ de, float* __restrict__ dp,
                                                                  int T, int U1, int J, int tch) {
  const int JV = J >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cvl = lane >> 3, uw = lane & 7;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tch, t1 = min(T, t0 + tch);
  for (int cv0 = warp * 4; cv0 < JV; cv0 += nw * 4) {
    const int cv = cv0 + cvl;
    const bool live = cv < JV;
    float dpa[kJv2NU][8];
#pragma unroll
    for (int k = 0; k < kJv2NU; ++k)
#pragma unroll
      for (int c = 0; c < 8; ++c) dpa[k][c] = 0.f;
    for (int t = t0; t < t1; ++t) {
      const __nv_bfloat16* base = dz + (((size_t)b * T + t) * U1) * J + (live ? cv : 0) * 8;
      float es[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < kJv2NU; ++k) {
        const int u = uw + 8 * k;
        if (live && u < U1) {
          float v[8];
          bf8_unpack(ld16(base + (size_t)u * J), v);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            es[c] += v[c];
            dpa[k][c] += v[c];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // sum over the 8 label lanes of this column vector (lanes differing in bits 0..2)
        es[c] += __shfl_xor_sync(0xffffffffu, es[c], 1);
        es[c] += __shfl_xor_sync(0xffffffffu, es[c], 2);
        es[c] += __shfl_xor_sync(0xffffffffu, es[c], 4);
      }
      if (live && uw == 0) {
        float* der = de + ((size_t)b * T + t) * J + cv * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) der[c] = es[c];
      }
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < kJv2NU; ++k) {
        const int u = uw + 8 * k;
        if (u < U1) {
          float4* dst = reinterpret_cast<float4*>(dp + ((size_t)b * U1 + u) * J + cv * 8);
          atomicAdd(dst, make_float4(dpa[k][0], dpa[k][1], dpa[k][2], dpa[k][3]));  // red.global.add.v4.f32
          atomicAdd(dst + 1, make_float4(dpa[k][4], dpa[k][5], dpa[k][6], dpa[k][7]));
        }
      }
    }
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_joint_tanh_fwd_v2(const void* e, const void* p, void* h, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream_) {
  if (!e || !p || !h || B < 0 || T < 1 || U1 < 1 || J < 8 || (J & 7) || B > 65535 || T > 2147483647 / 2 ||
      ((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(h)) & 15))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const int JV = J / 8;
  if (JV > 512) return TFASR_INVALID_VALUE;
  int ul = 512 / JV;
  if (ul > U1) ul = U1;
  if (ul > 8) ul = 8;
  const dim3 grid((unsigned)T, (unsigned)B), block((unsigned)JV, (unsigned)ul);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, 0, [&] {
    joint_tanh_fwd_v2_kernel(static_cast<const __nv_bfloat16*>(e), static_cast<const __nv_bfloat16*>(p), static_cast<__nv_bfloat16*>(h), T, U1, J);
  });
  return TFASR_OK;
#else
  joint_tanh_fwd_v2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(e), static_cast<const __nv_bfloat16*>(p),
                                                                                  static_cast<__nv_bfloat16*>(h), T, U1, J);
  count_launch();
  return check_launch();
#endif
}

extern "C" int tfasr_joint_reduce_bwd_v2(const void* dz, float* de, float* dp, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream_) {
  if (!dz || !de || !dp || B < 0 || T < 1 || U1 < 1 || U1 > 8 * kJv2NU || J < 8 || (J & 7) || B > 65535 ||
      ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dp)) & 15))
    return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  // ~2 CTAs per SM (what the 128 registers allow): frames per CTA from the batch size; every CTA ends with U1 * J / 4 vector reds into dp
  int64_t want = (int64_t)2 * kNumSMs / B;
  if (want < 1) want = 1;
  if (want > T) want = T;
  const int tch = (int)ceil_div(T, want);
  const dim3 grid((unsigned)ceil_div(T, tch), (unsigned)B), block(256);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
  emu::launch(grid, block, 0, [&] { joint_reduce_bwd_v2_kernel(static_cast<const __nv_bfloat16*>(dz), de, dp, T, U1, J, tch); });
  return TFASR_OK;
#else
  joint_reduce_bwd_v2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(dz), de, dp, T, U1, J, tch);
  count_launch();
  return check_launch();
#endif
}
