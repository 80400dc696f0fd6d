// Row-wise fused kernels of the Conformer block (sm_100a, HBM-bound, bf16 activations / fp32 statistics).
//
// Replace the chains of TF/Keras elementwise + reduction kernels around every Dense layer of the reference:
//   LayerNormalization fwd/bwd      tensorflow_asr/models/encoders/conformer.py:63,152,284,424-499 (keras LN, eps 1e-3)
//   residual add                    models/layers/residual.py:58-62   (fused into the LN backward / GEMM epilogues)
//   dropout backward + bias grads   models/layers/general.py Dropout   (mask regenerated from the counter-based hash)
//   GLU backward                    models/activations/glu.py:25-28
// One warp per row; rows are D <= 1024 wide (D = 144 in Conformer-S, 320 in the prediction net).
#include "common.cuh"
#include "dropout.cuh"

namespace tfasr {

constexpr int kRowsWarps = 8;
constexpr int kMaxPerLane = 32;  // D <= 1024

// ---------------------------------------------------------------------------------------------- LayerNorm forward
// NPL = values per lane (compile time, so the row lives in registers): D <= 32 * NPL.
template <typename TIn, int NPL>
__global__ void __launch_bounds__(kRowsWarps * 32) layernorm_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
                                                                        float eps) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const TIn* xr = x + (size_t)row * D;
  float v[NPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    v[i] = (j < D) ? (float)xr[j] : 0.f;
    s += v[i];
  }
  s = warp_sum(s);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    const float d = (j < D) ? v[i] - mean : 0.f;
    q += d * d;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)D + eps);
  __nv_bfloat16* yr = y + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    if (j < D) yr[j] = __float2bfloat16((v[i] - mean) * rstd * gamma[j] + beta[j]);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma          (+ dres: the residual branch's gradient)
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (CTA-level partial sums in SMEM, then one atomic per column per CTA)
template <typename TIn, int NPL>
__global__ void __launch_bounds__(kRowsWarps * 32) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const TIn* __restrict__ x,
                                                                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                        const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                                                                        __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                        float* __restrict__ dbeta, int M, int D, int rows_per_warp) {
  extern __shared__ float s_acc[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < 2 * D; j += blockDim.x) s_acc[j] = 0.f;
  __syncthreads();
  float ag[NPL], ab[NPL], gm[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    ag[i] = ab[i] = 0.f;
    gm[i] = (j < D) ? gamma[j] : 0.f;
  }
  const int row0 = (blockIdx.x * kRowsWarps + warp) * rows_per_warp;
  for (int r = 0; r < rows_per_warp; ++r) {
    const int row = row0 + r;
    if (row >= M) break;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const __nv_bfloat16* dyr = dy + (size_t)row * D;
    const TIn* xr = x + (size_t)row * D;
    float g[NPL], xh[NPL];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int j = lane + 32 * i;
      const float d = (j < D) ? __bfloat162float(dyr[j]) : 0.f;
      xh[i] = (j < D) ? ((float)xr[j] - mean) * rstd : 0.f;
      g[i] = d * gm[i];
      sg += g[i];
      sgx += g[i] * xh[i];
      ag[i] += d * xh[i];
      ab[i] += d;
    }
    sg = warp_sum(sg) / (float)D;
    sgx = warp_sum(sgx) / (float)D;
    __nv_bfloat16* dxr = dx + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int j = lane + 32 * i;
      if (j < D) {
        float o = rstd * (g[i] - sg - xh[i] * sgx);
        if (dres) o += __bfloat162float(dres[(size_t)row * D + j]);
        dxr[j] = __float2bfloat16(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int j = lane + 32 * i;
    if (j < D) {
      atomicAdd(&s_acc[j], ag[i]);
      atomicAdd(&s_acc[D + j], ab[i]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + j, s_acc[j]);
    if (dbeta) atomicAdd(dbeta + j, s_acc[D + j]);
  }
}

// bf16 x, even D: column pairs per lane (4-byte loads, 128 B per warp instruction), the rows of a warp loaded up front (RPW x 2 x NP2
// independent loads in flight per lane) - the scalar kernel above ran at 0.7 TB/s on [8000, 144].
template <int NP2, int RPW>
__global__ void __launch_bounds__(kRowsWarps * 32) layernorm_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                            const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                            float* __restrict__ dbeta, int M, int D) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  extern __shared__ float s_acc[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D2 = D >> 1;
  for (int j = threadIdx.x; j < 2 * D; j += blockDim.x) s_acc[j] = 0.f;
  __syncthreads();
  float2 ag[NP2], ab[NP2], gm[NP2];
#pragma unroll
  for (int i = 0; i < NP2; ++i) {
    const int pi = lane + 32 * i;
    ag[i] = ab[i] = make_float2(0.f, 0.f);
    gm[i] = (pi < D2) ? make_float2(gamma[2 * pi], gamma[2 * pi + 1]) : make_float2(0.f, 0.f);  // parameters are only 4-byte aligned in general
  }
  const int row0 = (blockIdx.x * kRowsWarps + warp) * RPW;
  __nv_bfloat162 dv[RPW][NP2], xv[RPW][NP2], rv[RPW][NP2];
  float mu[RPW], rs[RPW];
  const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = row0 + r;
    const bool ok = row < M;
    mu[r] = ok ? mean_in[row] : 0.f;
    rs[r] = ok ? rstd_in[row] : 0.f;
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      const int pi = lane + 32 * i;
      const bool in = ok && pi < D2;
      dv[r][i] = in ? *reinterpret_cast<const __nv_bfloat162*>(dy + (size_t)row * D + 2 * pi) : z2;
      xv[r][i] = in ? *reinterpret_cast<const __nv_bfloat162*>(x + (size_t)row * D + 2 * pi) : z2;
      rv[r][i] = (in && dres) ? *reinterpret_cast<const __nv_bfloat162*>(dres + (size_t)row * D + 2 * pi) : z2;
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = row0 + r;
    if (row >= M) break;
    float2 g[NP2], xh[NP2];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      const int pi = lane + 32 * i;
      const float2 d = __bfloat1622float2(dv[r][i]), xx = __bfloat1622float2(xv[r][i]);
      xh[i] = (pi < D2) ? make_float2((xx.x - mu[r]) * rs[r], (xx.y - mu[r]) * rs[r]) : make_float2(0.f, 0.f);
      g[i] = make_float2(d.x * gm[i].x, d.y * gm[i].y);
      sg += g[i].x + g[i].y;
      sgx += g[i].x * xh[i].x + g[i].y * xh[i].y;
      ag[i].x += d.x * xh[i].x;
      ag[i].y += d.y * xh[i].y;
      ab[i].x += d.x;
      ab[i].y += d.y;
    }
    sg = warp_sum(sg) / (float)D;
    sgx = warp_sum(sgx) / (float)D;
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      const int pi = lane + 32 * i;
      if (pi < D2) {
        const float2 rr = __bfloat1622float2(rv[r][i]);
        const float o0 = rs[r] * (g[i].x - sg - xh[i].x * sgx) + rr.x, o1 = rs[r] * (g[i].y - sg - xh[i].y * sgx) + rr.y;
        *reinterpret_cast<__nv_bfloat162*>(dx + (size_t)row * D + 2 * pi) = __floats2bfloat162_rn(o0, o1);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NP2; ++i) {
    const int pi = lane + 32 * i;
    if (pi < D2) {
      atomicAdd(&s_acc[2 * pi], ag[i].x);
      atomicAdd(&s_acc[2 * pi + 1], ag[i].y);
      atomicAdd(&s_acc[D + 2 * pi], ab[i].x);
      atomicAdd(&s_acc[D + 2 * pi + 1], ab[i].y);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + j, s_acc[j]);
    if (dbeta) atomicAdd(dbeta + j, s_acc[D + j]);
  }
}

// ---------------------------------------------------------------------------------------------- elementwise helpers
// out = scale * dropout_mask * in (bf16) [+ add], colsum += column sums of out.   [M, N] row-major, N even.
__global__ void __launch_bounds__(256) scale_mask_colsum_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ add,
                                                                __nv_bfloat16* __restrict__ out, float* __restrict__ colsum, int M, int N,
                                                                float scale, float drop_p, const unsigned long long* __restrict__ seed_ptr,
                                                                unsigned long long salt, int rows_per_block) {
  // block = [rows_per_block rows] x all columns; thread t handles column pair (2t, 2t+1) (+ stride blockDim for wide N)
  extern __shared__ uint32_t s_rowkey[];  // dropout row keys of this block's rows (64-bit work once per row, not per element)
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const unsigned long long seed = (drop_p > 0.f && seed_ptr) ? (*seed_ptr + salt) : salt;
  const uint32_t drop_thr = dropout_threshold(drop_p);
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  if (drop_p > 0.f) {
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) s_rowkey[r - r0] = dropout_rowkey(seed, (unsigned long long)r);
    __syncthreads();
  }
  for (int c = 2 * threadIdx.x; c < N; c += 2 * blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const size_t idx = (size_t)r * N + c;
      const __nv_bfloat162 vi = *reinterpret_cast<const __nv_bfloat162*>(in + idx);
      float a = __low2float(vi) * scale, b = __high2float(vi) * scale;
      if (drop_p > 0.f) {
        const uint32_t hsh = dropout_hash(s_rowkey[r - r0], (uint32_t)c);  // c is even: one hash serves the column pair
        a = ((hsh & 0xFFFFu) >= drop_thr) ? a * keep_scale : 0.f;
        b = ((hsh >> 16) >= drop_thr) ? b * keep_scale : 0.f;
      }
      if (add) {
        const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(add + idx);
        a += __low2float(va);
        b += __high2float(va);
      }
      if (out) *reinterpret_cast<__nv_bfloat162*>(out + idx) = __floats2bfloat162_rn(a, b);
      s0 += a;
      s1 += b;
    }
    if (colsum) {
      atomicAdd(colsum + c, s0);
      atomicAdd(colsum + c + 1, s1);
    }
  }
}

// Vector variant (N % 8 == 0, N / 8 <= 256): thread = (row lane, 8-column vector), 16-byte loads / stores, column sums reduced in SMEM.
__global__ void __launch_bounds__(256) scale_mask_colsum_vec_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ add,
                                                                    __nv_bfloat16* __restrict__ out, float* __restrict__ colsum, int M, int N,
                                                                    float scale, float drop_p, const unsigned long long* __restrict__ seed_ptr,
                                                                    unsigned long long salt, int rows_per_block) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  extern __shared__ float s_col[];  // [N]
  const int NV = N >> 3, RL = blockDim.x / NV;
  const int cv = threadIdx.x % NV, rl = threadIdx.x / NV;
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const unsigned long long seed = (drop_p > 0.f && seed_ptr) ? (*seed_ptr + salt) : salt;
  const uint32_t drop_thr = dropout_threshold(drop_p);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (colsum) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_col[i] = 0.f;
    __syncthreads();
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < RL) {
    for (int rb = r0 + rl; rb < r1; rb += 4 * RL) {
      // four rows per trip, all loads issued before the first use (the scalar loop exposed one L2 round trip per row)
      uint4 raw4[4], add4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rb + k * RL;
        if (r < r1) {
          raw4[k] = *reinterpret_cast<const uint4*>(in + (size_t)r * N + cv * 8);
          if (add) add4[k] = *reinterpret_cast<const uint4*>(add + (size_t)r * N + cv * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rb + k * RL;
        if (r >= r1) break;
        const size_t idx = (size_t)r * N + cv * 8;
        const uint32_t w[4] = {raw4[k].x, raw4[k].y, raw4[k].z, raw4[k].w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] = __uint_as_float(w[j] << 16) * scale;
          v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u) * scale;
        }
        if (drop_p > 0.f) {
          const uint32_t key = dropout_rowkey(seed, (unsigned long long)r);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t hsh = dropout_hash(key, (uint32_t)(cv * 8 + 2 * j));
            v[2 * j] = ((hsh & 0xFFFFu) >= drop_thr) ? v[2 * j] * keep_scale : 0.f;
            v[2 * j + 1] = ((hsh >> 16) >= drop_thr) ? v[2 * j + 1] * keep_scale : 0.f;
          }
        }
        if (add) {
          const uint32_t wa[4] = {add4[k].x, add4[k].y, add4[k].z, add4[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] += __uint_as_float(wa[j] << 16);
            v[2 * j + 1] += __uint_as_float(wa[j] & 0xFFFF0000u);
          }
        }
        if (out) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(out + idx) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
  }
  if (colsum) {
    if (rl < RL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_col[cv * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(colsum + i, s_col[i]);
  }
}

// Column sums only, any width with N % 8 == 0: grid (row blocks, column blocks of NVB 8-column vectors); the same four-rows-in-flight loads as
// the vector kernel above.  Used when N / 8 > 256 (the all-blocks position projection: N = 16 x 144).
__global__ void __launch_bounds__(256) colsum_wide_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ colsum, int M, int N, int NVB,
                                                          int rows_per_block) {
  extern __shared__ float s_col[];  // [8 NVB]
  const int RL = blockDim.x / NVB;
  const int cvl = threadIdx.x % NVB, rl = threadIdx.x / NVB;
  const int c0 = (blockIdx.y * NVB + cvl) * 8;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int i = threadIdx.x; i < 8 * NVB; i += blockDim.x) s_col[i] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < RL) {
    for (int rb = r0 + rl; rb < r1; rb += 4 * RL) {
      uint4 raw4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rb + k * RL;
        if (r < r1) raw4[k] = *reinterpret_cast<const uint4*>(in + (size_t)r * N + c0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (rb + k * RL >= r1) break;
        const uint32_t w[4] = {raw4[k].x, raw4[k].y, raw4[k].z, raw4[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_col[cvl * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * NVB; i += blockDim.x) atomicAdd(colsum + (size_t)blockIdx.y * NVB * 8 + i, s_col[i]);
}

// GLU backward on interleaved (a, b) pairs: aux [M, 2D] pre-GLU values, dg [M, D] -> daux [M, 2D]; colsum [2D] += column sums.
__global__ void __launch_bounds__(256) glu_bwd_kernel(const __nv_bfloat16* __restrict__ aux, const __nv_bfloat16* __restrict__ dg,
                                                      __nv_bfloat16* __restrict__ daux, float* __restrict__ colsum, int M, int D,
                                                      int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sa = 0.f, sb = 0.f;
    for (int r = r0; r < r1; ++r) {
      const __nv_bfloat162 ab = *reinterpret_cast<const __nv_bfloat162*>(aux + (size_t)r * 2 * D + 2 * c);
      const float a = __low2float(ab), b = __high2float(ab);
      const float g = __bfloat162float(dg[(size_t)r * D + c]);
      const float sg = 1.0f / (1.0f + __expf(-b));
      const float da = g * sg;
      const float db = g * a * sg * (1.0f - sg);
      *reinterpret_cast<__nv_bfloat162*>(daux + (size_t)r * 2 * D + 2 * c) = __floats2bfloat162_rn(da, db);
      sa += da;
      sb += db;
    }
    if (colsum) {
      atomicAdd(colsum + 2 * c, sa);
      atomicAdd(colsum + 2 * c + 1, sb);
    }
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_layernorm_fwd(const void* x, int32_t x_is_f32, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                   int32_t M, int32_t D, float eps, tfasr_stream_t stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd || M < 0 || D < 1 || D > 32 * kMaxPerLane) return TFASR_INVALID_VALUE;
  if (M == 0) return TFASR_OK;
  const unsigned grid = (unsigned)ceil_div(M, kRowsWarps);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_FWD(T, NPL) launch_pdl(layernorm_fwd_kernel<T, NPL>, dim3(grid), dim3(kRowsWarps * 32), 0, st, static_cast<const T*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, M, D, eps)
  if (x_is_f32) {
    if (D <= 160) LN_FWD(float, 5); else if (D <= 320) LN_FWD(float, 10); else if (D <= 640) LN_FWD(float, 20); else LN_FWD(float, 32);
  } else {
    if (D <= 160) LN_FWD(__nv_bfloat16, 5); else if (D <= 320) LN_FWD(__nv_bfloat16, 10); else if (D <= 640) LN_FWD(__nv_bfloat16, 20); else LN_FWD(__nv_bfloat16, 32);
  }
#undef LN_FWD
  count_launch();
  return check_launch();
}

extern "C" int tfasr_layernorm_bwd(const void* dy, const void* x, int32_t x_is_f32, const float* mean, const float* rstd, const float* gamma,
                                   const void* dres, void* dx, float* dgamma, float* dbeta, int32_t M, int32_t D, tfasr_stream_t stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dx || M < 0 || D < 1 || D > 32 * kMaxPerLane) return TFASR_INVALID_VALUE;
  if (M == 0) return TFASR_OK;
  const int rows_per_warp = 4;
  const unsigned grid = (unsigned)ceil_div(M, kRowsWarps * rows_per_warp);
  const size_t smem = sizeof(float) * 2 * D;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!x_is_f32 && (D & 1) == 0 && D <= 512) {
#define LN_BWD_VEC(NP2)                                                                                                                 \
  launch_pdl(layernorm_bwd_vec_kernel<NP2, 4>, dim3(grid), dim3(kRowsWarps * 32), smem, st, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), \
                                                                        mean, rstd, gamma, static_cast<const __nv_bfloat16*>(dres),       \
                                                                        static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, M, D)
    if (D <= 192) LN_BWD_VEC(3); else if (D <= 320) LN_BWD_VEC(5); else LN_BWD_VEC(8);
#undef LN_BWD_VEC
    count_launch();
    return check_launch();
  }
#define LN_BWD(T, NPL)                                                                                                                  \
  layernorm_bwd_kernel<T, NPL><<<grid, kRowsWarps * 32, smem, st>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const T*>(x), mean, rstd, gamma, \
                                                                     static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, M, D, rows_per_warp)
  if (x_is_f32) {
    if (D <= 160) LN_BWD(float, 5); else if (D <= 320) LN_BWD(float, 10); else if (D <= 640) LN_BWD(float, 20); else LN_BWD(float, 32);
  } else {
    if (D <= 160) LN_BWD(__nv_bfloat16, 5); else if (D <= 320) LN_BWD(__nv_bfloat16, 10); else if (D <= 640) LN_BWD(__nv_bfloat16, 20); else LN_BWD(__nv_bfloat16, 32);
  }
#undef LN_BWD
  count_launch();
  return check_launch();
}

extern "C" int tfasr_scale_mask_colsum(const void* in, const void* add, void* out, float* colsum, int32_t M, int32_t N, float scale, float dropout_p,
                                       const unsigned long long* seed_ptr, unsigned long long salt, tfasr_stream_t stream) {
  if (!in || (!out && !colsum) || M < 0 || N < 2 || (N & 1)) return TFASR_INVALID_VALUE;  // out == NULL: column sums only
  if (M == 0) return TFASR_OK;
  if ((N & 7) == 0 && (N >> 3) <= 256 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int RL = 256 / (N >> 3);
    int rows_per_block = RL * 4;
    while ((long long)ceil_div(M, rows_per_block) > 8LL * kNumSMs) rows_per_block *= 2;  // bounded atomic traffic on colsum
    launch_pdl(scale_mask_colsum_vec_kernel, dim3((unsigned)ceil_div(M, rows_per_block)), dim3(256), sizeof(float) * N, static_cast<cudaStream_t>(stream),
        static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(out), colsum, M, N, scale, dropout_p,
        seed_ptr, salt, rows_per_block);
    count_launch();
    return check_launch();
  }
  if (!out && !add && scale == 1.0f && dropout_p == 0.f && (N & 7) == 0 && (N >> 3) > 256 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    int nvb = 0;  // widest column block (in 8-column vectors) that divides N / 8 and fits the CTA
    for (int d = 256; d >= 32 && !nvb; --d)
      if ((N >> 3) % d == 0) nvb = d;
    if (nvb) {
      const int RL = 256 / nvb;
      int rows_per_block = RL * 4;
      while ((long long)ceil_div(M, rows_per_block) * ((N >> 3) / nvb) > 8LL * kNumSMs) rows_per_block *= 2;
      const dim3 grid((unsigned)ceil_div(M, rows_per_block), (unsigned)((N >> 3) / nvb));
      colsum_wide_kernel<<<grid, 256, sizeof(float) * 8 * nvb, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in), colsum, M, N,
                                                                                                   nvb, rows_per_block);
      count_launch();
      return check_launch();
    }
  }
  const int rows_per_block = 32;
  scale_mask_colsum_kernel<<<(unsigned)ceil_div(M, rows_per_block), 256, sizeof(uint32_t) * rows_per_block, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(out), colsum, M, N, scale, dropout_p,
      seed_ptr, salt, rows_per_block);
  count_launch();
  return check_launch();
}

namespace tfasr {
// Vector variant (D % 4 == 0, D / 4 <= 256): thread = (row lane, 4 GLU outputs): 16-byte aux / daux accesses (4 interleaved (a, b) pairs),
// 8-byte dg loads, four rows per trip with the loads issued first, column sums reduced in SMEM.
__global__ void __launch_bounds__(256) glu_bwd_vec_kernel(const __nv_bfloat16* __restrict__ aux, const __nv_bfloat16* __restrict__ dg,
                                                          __nv_bfloat16* __restrict__ daux, float* __restrict__ colsum, int M, int D,
                                                          int rows_per_block) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  extern __shared__ float s_col[];  // [2 D]
  const int NV = D >> 2, RL = blockDim.x / NV;
  const int cv = threadIdx.x % NV, rl = threadIdx.x / NV;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (colsum) {
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) s_col[i] = 0.f;
    __syncthreads();
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < RL) {
    for (int rb = r0 + rl; rb < r1; rb += 4 * RL) {
      uint4 ab4[4];
      uint2 g4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rb + k * RL;
        if (r < r1) {
          ab4[k] = *reinterpret_cast<const uint4*>(aux + (size_t)r * 2 * D + cv * 8);
          g4[k] = *reinterpret_cast<const uint2*>(dg + (size_t)r * D + cv * 4);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = rb + k * RL;
        if (r >= r1) break;
        const uint32_t w[4] = {ab4[k].x, ab4[k].y, ab4[k].z, ab4[k].w};
        const uint32_t gw[2] = {g4[k].x, g4[k].y};
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = __uint_as_float(w[j] << 16), b = __uint_as_float(w[j] & 0xFFFF0000u);
          const float g = (j & 1) ? __uint_as_float(gw[j >> 1] & 0xFFFF0000u) : __uint_as_float(gw[j >> 1] << 16);
          const float sg = 1.0f / (1.0f + __expf(-b));
          const float da = g * sg, db = g * a * sg * (1.0f - sg);
          const __nv_bfloat162 t = __floats2bfloat162_rn(da, db);
          pk[j] = *reinterpret_cast<const uint32_t*>(&t);
          acc[2 * j] += da;
          acc[2 * j + 1] += db;
        }
        *reinterpret_cast<uint4*>(daux + (size_t)r * 2 * D + cv * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  if (colsum) {
    if (rl < RL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_col[cv * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) atomicAdd(colsum + i, s_col[i]);
  }
}
}  // namespace tfasr

extern "C" int tfasr_glu_bwd(const void* aux, const void* dg, void* daux, float* colsum, int32_t M, int32_t D, tfasr_stream_t stream) {
  if (!aux || !dg || !daux || M < 0 || D < 1) return TFASR_INVALID_VALUE;
  if (M == 0) return TFASR_OK;
  if ((D & 3) == 0 && (D >> 2) <= 256 &&
      ((reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(daux)) & 15) == 0 && (reinterpret_cast<uintptr_t>(dg) & 7) == 0) {
    const int RL = 256 / (D >> 2);
    int rpb = RL * 4;
    while ((long long)ceil_div(M, rpb) > 8LL * kNumSMs) rpb *= 2;
    launch_pdl(glu_bwd_vec_kernel, dim3((unsigned)ceil_div(M, rpb)), dim3(256), sizeof(float) * 2 * D, static_cast<cudaStream_t>(stream),
        static_cast<const __nv_bfloat16*>(aux), static_cast<const __nv_bfloat16*>(dg), static_cast<__nv_bfloat16*>(daux), colsum, M, D, rpb);
    count_launch();
    return check_launch();
  }
  const int rows_per_block = 32;
  glu_bwd_kernel<<<(unsigned)ceil_div(M, rows_per_block), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(aux), static_cast<const __nv_bfloat16*>(dg), static_cast<__nv_bfloat16*>(daux), colsum, M, D, rows_per_block);
  count_launch();
  return check_launch();
}
