// Transducer joint network, training path (sm_100a): the elementwise / reduction halves around the tcgen05 GEMMs.
//
// Replaces TransducerJointMerge + tanh (tensorflow_asr/models/transducer/base_transducer.py:199-207,291) and its backward:
//   joint_tanh_fwd    h[b,t,u,:] = tanh(e[b,t,:] + p[b,u,:])                  bf16, never materialised in fp32
//   joint_reduce_bwd  de[b,t,:] = sum_u dz[b,t,u,:],  dp[b,u,:] = sum_t dz[b,t,u,:]     (dz = dh * (1 - h^2) comes out of the
//                     data-gradient GEMM's epilogue)
// Both are HBM-bound streaming kernels over [B, T, U1, J] bf16 (333 MB at B=32, T=250, U1=65, J=320).
#include "common.cuh"

namespace tfasr {

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// grid (T, B); block 256.  J even.
__global__ void __launch_bounds__(256) joint_tanh_fwd_kernel(const __nv_bfloat16* __restrict__ e, const __nv_bfloat16* __restrict__ p,
                                                             __nv_bfloat16* __restrict__ h, int T, int U1, int J) {
  extern __shared__ float s_e[];  // [J]
  const int t = blockIdx.x, b = blockIdx.y;
  const __nv_bfloat16* er = e + ((size_t)b * T + t) * J;
  for (int j = threadIdx.x; j < J; j += blockDim.x) s_e[j] = __bfloat162float(er[j]);
  __syncthreads();
  const int J2 = J >> 1;
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(p + (size_t)b * U1 * J);
  __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(h + (((size_t)b * T + t) * U1) * J);
  const int total = U1 * J2;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int j2 = i % J2;
    const __nv_bfloat162 pv = pb[i];
    const float a = tanh_fast(s_e[2 * j2] + __low2float(pv));
    const float c = tanh_fast(s_e[2 * j2 + 1] + __high2float(pv));
    hb[i] = __floats2bfloat162_rn(a, c);
  }
}

// grid (ceil(T / TCH), B); block = J/2 threads (each owns a column pair); dp partial sums in SMEM, one atomic per (u, j) per CTA.
constexpr int kJointTch = 8;
__global__ void __launch_bounds__(512) joint_reduce_bwd_kernel(const __nv_bfloat16* __restrict__ dz, float* __restrict__ de, float* __restrict__ dp,
                                                               int T, int U1, int J) {
  extern __shared__ float s_dp[];  // [U1][J]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kJointTch;
  const int t1 = min(T, t0 + kJointTch);
  const int J2 = J >> 1;
  for (int i = threadIdx.x; i < U1 * J; i += blockDim.x) s_dp[i] = 0.f;
  __syncthreads();
  for (int j2 = threadIdx.x; j2 < J2; j2 += blockDim.x) {
    for (int t = t0; t < t1; ++t) {
      const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(dz + (((size_t)b * T + t) * U1) * J) + j2;
      float ea = 0.f, eb = 0.f;
      for (int u = 0; u < U1; ++u) {
        const __nv_bfloat162 v = row[(size_t)u * J2];
        const float a = __low2float(v), c = __high2float(v);
        ea += a;
        eb += c;
        s_dp[u * J + 2 * j2] += a;  // this thread is the only writer of its column pair
        s_dp[u * J + 2 * j2 + 1] += c;
      }
      float* der = de + ((size_t)b * T + t) * J + 2 * j2;
      der[0] = ea;
      der[1] = eb;
    }
  }
  __syncthreads();
  float* dpb = dp + (size_t)b * U1 * J;
  for (int i = threadIdx.x; i < U1 * J; i += blockDim.x) atomicAdd(dpb + i, s_dp[i]);
}


// Wide variant (the one the step uses): the lattice of one utterance is streamed once with 16-byte loads.  Threads = (J/8 column vectors) x UW
// label lanes; a thread keeps the prediction-side partial sums of its <= 9 labels x 8 columns in registers over the CTA's frames, the
// encoder-side sums go through a small SMEM reduction per frame.  HBM-bound: 2 bytes per lattice node and joint column.
constexpr int kJrNU = 9;
template <int UW>
__global__ void __launch_bounds__(512) joint_reduce_bwd_wide_kernel(const __nv_bfloat16* __restrict__ dz, float* __restrict__ de,
                                                                     float* __restrict__ dp, int T, int U1, int J, int tch) {
  extern __shared__ float s_e[];  // [UW][J]
  const int JV = J >> 3;
  const int cv = threadIdx.x % JV, uw = threadIdx.x / JV;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tch, t1 = min(T, t0 + tch);
  float dpa[kJrNU][8];
#pragma unroll
  for (int k = 0; k < kJrNU; ++k)
#pragma unroll
    for (int c = 0; c < 8; ++c) dpa[k][c] = 0.f;
  for (int t = t0; t < t1; ++t) {
    const __nv_bfloat16* base = dz + (((size_t)b * T + t) * U1) * J + cv * 8;
    uint4 raw[kJrNU];
#pragma unroll
    for (int k = 0; k < kJrNU; ++k) {
      const int u = uw + UW * k;
      raw[k] = (u < U1) ? ld_stream_u4(base + (size_t)u * J) : make_uint4(0u, 0u, 0u, 0u);
    }
    float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kJrNU; ++k) {
      const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float lo = __uint_as_float(w[c] << 16), hi = __uint_as_float(w[c] & 0xFFFF0000u);
        e[2 * c] += lo;
        e[2 * c + 1] += hi;
        dpa[k][2 * c] += lo;
        dpa[k][2 * c + 1] += hi;
      }
    }
    float4* se = reinterpret_cast<float4*>(s_e + uw * J + cv * 8);
    se[0] = make_float4(e[0], e[1], e[2], e[3]);
    se[1] = make_float4(e[4], e[5], e[6], e[7]);
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < UW; ++w) acc += s_e[w * J + j];
      de[((size_t)b * T + t) * J + j] = acc;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kJrNU; ++k) {
    const int u = uw + UW * k;
    if (u < U1) {
      float4* dst = reinterpret_cast<float4*>(dp + ((size_t)b * U1 + u) * J + cv * 8);
      atomicAdd(dst, make_float4(dpa[k][0], dpa[k][1], dpa[k][2], dpa[k][3]));  // red.global.add.v4.f32
      atomicAdd(dst + 1, make_float4(dpa[k][4], dpa[k][5], dpa[k][6], dpa[k][7]));
    }
  }
}

template <int UW>
static int launch_joint_reduce_wide(const __nv_bfloat16* dz, float* de, float* dp, int B, int T, int U1, int J, cudaStream_t stream) {
  const int threads = (J / 8) * UW;
  const int chunks = T < 5 ? 1 : 5;  // ~1 CTA per SM at B = 32; each keeps 9 x 16-byte loads per thread in flight
  const int tch = (T + chunks - 1) / chunks;
  joint_reduce_bwd_wide_kernel<UW><<<dim3((unsigned)ceil_div(T, tch), B), threads, sizeof(float) * UW * J, stream>>>(dz, de, dp, T, U1, J, tch);
  count_launch();
  return check_launch();
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_joint_tanh_fwd(const void* e, const void* p, void* h, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream) {
  if (!e || !p || !h || B < 0 || T < 1 || U1 < 1 || J < 2 || (J & 1) || B > 65535) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  joint_tanh_fwd_kernel<<<dim3(T, B), 256, sizeof(float) * J, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(e), static_cast<const __nv_bfloat16*>(p), static_cast<__nv_bfloat16*>(h), T, U1, J);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_joint_reduce_bwd(const void* dz, float* de, float* dp, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream) {
  if (!dz || !de || !dp || B < 0 || T < 1 || U1 < 1 || J < 2 || (J & 1) || B > 65535) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  if ((J & 7) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (reinterpret_cast<uintptr_t>(dp) & 15) == 0) {
    const __nv_bfloat16* dzb = static_cast<const __nv_bfloat16*>(dz);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int jv = J / 8;
    // 72 + 36 live registers per thread: at most 512 threads per CTA
    if (U1 <= 8 * kJrNU && jv * 8 <= 512) return launch_joint_reduce_wide<8>(dzb, de, dp, B, T, U1, J, st);
    if (U1 <= 16 * kJrNU && jv * 16 <= 512) return launch_joint_reduce_wide<16>(dzb, de, dp, B, T, U1, J, st);
  }
  const size_t smem = sizeof(float) * (size_t)U1 * J;
  if (smem > 200 * 1024) return TFASR_INVALID_VALUE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(joint_reduce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  int threads = J / 2;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  joint_reduce_bwd_kernel<<<dim3((unsigned)ceil_div(T, kJointTch), B), threads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dz), de, dp, T, U1, J);
  count_launch();
  return check_launch();
}
