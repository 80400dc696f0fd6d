// Relative-position multi-head self-attention (Transformer-XL style) for sm_100a, forward and backward.
//
// Replaces MultiHeadRelativeAttention._compute_attention (tensorflow_asr/models/layers/multihead_attention.py:543-582): two einsums,
// rel_left_shift (:27-64: pad / reshape / slice chain materialising [B,N,T,2T-1] twice), masked softmax
// (models/layers/general.py:30-41, fill -1e9) and the value einsum — ~12 kernels and three [B,N,T,T]-sized round trips in the
// reference — by one forward kernel and two backward kernels.  The shift is index arithmetic: key j is at relative row
// r = T - 1 - t + j of the projected position table.
//
//   scores[t,j] = ((q_t + u) . k_j + (q_t + v) . pos_{T-1-t+j}) / sqrt(H);   query-side mask only (t >= len_b -> every key -1e9),
//   optional streaming mask (chunk / history, multihead_attention.py:104-143);   out_t = softmax_j(scores[t,:]) . V
//
// Sizes here are small (T' = 250, H = 36 per head): the kernels keep K, V and the needed window of the position table in SMEM and
// run on the FP32 pipes; they are latency / SMEM-bandwidth bound, not tensor-core candidates at H = 36.
#include "common.cuh"

namespace tfasr {

constexpr int kAttThreads = 128;  // 4 warps
constexpr int kAttTQ = 16;        // query rows per CTA (forward / backward-A)
constexpr int kMaxH = 64;

struct AttParams {
  const __nv_bfloat16* qkv;  // [B*T, ld_qkv]: q at col n*H, k at NH + n*H, v at 2NH + n*H
  const __nv_bfloat16* pos;  // [B*R, NH]
  const float* ubias;        // [N, H]
  const float* vbias;        // [N, H]
  const int32_t* length;     // [B] or null
  int B, T, N, H, R, ld_qkv;
  int chunk, history;        // streaming mask if chunk > 0
  float scale;
};

__device__ __forceinline__ bool att_allowed(int t, int j, int len, int T, int chunk, int history) {
  if (t >= len) return false;  // query-side mask
  if (chunk > 0) {
    const int c0 = (t / chunk) * chunk;
    const int hist = history < 0 ? T : history;
    const int start = max(0, c0 - hist), end = min(T, c0 + chunk);
    return j >= start && j < end;
  }
  return true;
}

// smem row stride (in bf16 elements) for K / V / P tiles: H + 2 keeps 4-byte alignment and an odd word stride (H = 36 -> 19 words)
__device__ __forceinline__ int att_stride(int H) { return ((H + 2) | 2); }

// ---------------------------------------------------------------------------------------------- forward
// grid (ceil(T/TQ), N, B).  MAXIT * 32 >= T.
template <int MAXIT>
__global__ void __launch_bounds__(kAttThreads) rel_attention_fwd_kernel(const AttParams p, __nv_bfloat16* __restrict__ ctx,
                                                                        __nv_bfloat16* __restrict__ probs) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int b = blockIdx.z, n = blockIdx.y, t0 = blockIdx.x * kAttTQ;
  const int S = att_stride(H);
  const int PW = T + kAttTQ - 1;  // rows of the position table this CTA can touch
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + (size_t)T * S;
  __nv_bfloat16* sP = sV + (size_t)T * S;
  float* sProb = reinterpret_cast<float*>(sP + (size_t)PW * S);  // [4 warps][T]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = p.length ? p.length[b] : T;
  // position-table window: rows r in [T-1-(t0+TQ-1), T-1-t0+T-1]
  const int r_lo = T - 1 - min(T - 1, t0 + kAttTQ - 1);
  {
    const int H2 = H >> 1;
    for (int i = threadIdx.x; i < T * H2; i += kAttThreads) {
      const int j = i / H2, h2 = i % H2;
      const size_t g = ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H + 2 * h2;
      *reinterpret_cast<__nv_bfloat162*>(sK + j * S + 2 * h2) = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + g + NH);
      *reinterpret_cast<__nv_bfloat162*>(sV + j * S + 2 * h2) = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + g + 2 * NH);
    }
    for (int i = threadIdx.x; i < PW * H2; i += kAttThreads) {
      const int rr = i / H2, h2 = i % H2;
      const int r = r_lo + rr;
      __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
      if (r < R) v = *reinterpret_cast<const __nv_bfloat162*>(p.pos + ((size_t)b * R + r) * NH + (size_t)n * H + 2 * h2);
      *reinterpret_cast<__nv_bfloat162*>(sP + rr * S + 2 * h2) = v;
    }
  }
  __syncthreads();
  float* myProb = sProb + (size_t)warp * T;
  for (int rt = warp; rt < kAttTQ; rt += 4) {
    const int t = t0 + rt;
    if (t >= T) break;
    // (q + u) * scale and (q + v) * scale in registers (every lane holds the whole head vector)
    float qu[kMaxH], qv[kMaxH];
    const __nv_bfloat16* qrow = p.qkv + ((size_t)b * T + t) * p.ld_qkv + (size_t)n * H;
#pragma unroll
    for (int h = 0; h < kMaxH; ++h) {
      if (h < H) {
        const float q = __bfloat162float(qrow[h]);
        qu[h] = (q + p.ubias[n * H + h]) * p.scale;
        qv[h] = (q + p.vbias[n * H + h]) * p.scale;
      }
    }
    float sc[MAXIT];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = lane + 32 * it;
      float s = -CUDART_INF_F;
      if (j < T) {
        const __nv_bfloat16* kr = sK + j * S;
        const __nv_bfloat16* pr = sP + (T - 1 - t + j - r_lo) * S;
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < kMaxH; h += 2) {
          if (h < H) {
            const __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(kr + h);
            const __nv_bfloat162 pp = *reinterpret_cast<const __nv_bfloat162*>(pr + h);
            acc = fmaf(qu[h], __low2float(kk), acc);
            acc = fmaf(qu[h + 1], __high2float(kk), acc);
            acc = fmaf(qv[h], __low2float(pp), acc);
            acc = fmaf(qv[h + 1], __high2float(pp), acc);
          }
        }
        s = att_allowed(t, j, len, T, p.chunk, p.history) ? acc : -1e9f;  // masked_fill value of general.py:30-41
      }
      sc[it] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = lane + 32 * it;
      const float e = (j < T) ? __expf(sc[it] - m) : 0.f;
      sc[it] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __nv_bfloat16* prow = probs + (((size_t)b * p.N + n) * T + t) * T;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = lane + 32 * it;
      if (j < T) {
        const __nv_bfloat16 pb = __float2bfloat16(sc[it] * inv);
        prow[j] = pb;
        myProb[j] = __bfloat162float(pb);  // the backward sees the bf16-rounded probabilities: use the same here
      }
    }
    __syncwarp();
    // out[h] = sum_j prob[j] * V[j][h]; lanes over h
    for (int h = lane; h < H; h += 32) {
      float acc = 0.f;
      for (int j = 0; j < T; ++j) acc = fmaf(myProb[j], __bfloat162float(sV[j * S + h]), acc);
      ctx[((size_t)b * T + t) * NH + (size_t)n * H + h] = __float2bfloat16(acc);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- backward A (per query row)
// ds[t,j] = p[t,j] * (dp[t,j] - sum_j p dp),  dp[t,j] = dctx_t . v_j  (zero where masked);  dq_t = scale * sum_j ds (k_j + pos_idx)
// du[n,:] += sum_t scale * sum_j ds k_j ;  dvb[n,:] += sum_t scale * sum_j ds pos_idx
template <int MAXIT>
__global__ void __launch_bounds__(kAttThreads) rel_attention_bwd_rows_kernel(const AttParams p, const __nv_bfloat16* __restrict__ probs,
                                                                             const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ ds_out,
                                                                             __nv_bfloat16* __restrict__ dqkv, float* __restrict__ du,
                                                                             float* __restrict__ dvb) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int b = blockIdx.z, n = blockIdx.y, t0 = blockIdx.x * kAttTQ;
  const int S = att_stride(H);
  const int PW = T + kAttTQ - 1;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + (size_t)T * S;
  __nv_bfloat16* sP = sV + (size_t)T * S;
  float* sDs = reinterpret_cast<float*>(sP + (size_t)PW * S);  // [4 warps][T]
  float* sAcc = sDs + 4 * T;                                   // [2][H] du / dvb partial sums of this CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = p.length ? p.length[b] : T;
  const int r_lo = T - 1 - min(T - 1, t0 + kAttTQ - 1);
  {
    const int H2 = H >> 1;
    for (int i = threadIdx.x; i < T * H2; i += kAttThreads) {
      const int j = i / H2, h2 = i % H2;
      const size_t g = ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H + 2 * h2;
      *reinterpret_cast<__nv_bfloat162*>(sK + j * S + 2 * h2) = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + g + NH);
      *reinterpret_cast<__nv_bfloat162*>(sV + j * S + 2 * h2) = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + g + 2 * NH);
    }
    for (int i = threadIdx.x; i < PW * H2; i += kAttThreads) {
      const int rr = i / H2, h2 = i % H2;
      const int r = r_lo + rr;
      __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
      if (r < R) v = *reinterpret_cast<const __nv_bfloat162*>(p.pos + ((size_t)b * R + r) * NH + (size_t)n * H + 2 * h2);
      *reinterpret_cast<__nv_bfloat162*>(sP + rr * S + 2 * h2) = v;
    }
    for (int i = threadIdx.x; i < 2 * H; i += kAttThreads) sAcc[i] = 0.f;
  }
  __syncthreads();
  float* myDs = sDs + (size_t)warp * T;
  for (int rt = warp; rt < kAttTQ; rt += 4) {
    const int t = t0 + rt;
    if (t >= T) break;
    float dc[kMaxH];
    const __nv_bfloat16* drow = dctx + ((size_t)b * T + t) * NH + (size_t)n * H;
#pragma unroll
    for (int h = 0; h < kMaxH; ++h)
      if (h < H) dc[h] = __bfloat162float(drow[h]);
    const __nv_bfloat16* prow = probs + (((size_t)b * p.N + n) * T + t) * T;
    float pr[MAXIT], dp[MAXIT];
    float delta = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = lane + 32 * it;
      pr[it] = 0.f;
      dp[it] = 0.f;
      if (j < T) {
        pr[it] = __bfloat162float(prow[j]);
        const __nv_bfloat16* vr = sV + j * S;
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < kMaxH; h += 2) {
          if (h < H) {
            const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(vr + h);
            acc = fmaf(dc[h], __low2float(vv), acc);
            acc = fmaf(dc[h + 1], __high2float(vv), acc);
          }
        }
        dp[it] = acc;
        delta = fmaf(pr[it], acc, delta);
      }
    }
    delta = warp_sum(delta);
    __nv_bfloat16* dsrow = ds_out + (((size_t)b * p.N + n) * T + t) * T;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int j = lane + 32 * it;
      if (j < T) {
        float d = pr[it] * (dp[it] - delta);
        if (!att_allowed(t, j, len, T, p.chunk, p.history)) d = 0.f;  // masked scores are constants
        const __nv_bfloat16 db = __float2bfloat16(d);
        dsrow[j] = db;
        myDs[j] = __bfloat162float(db);
      }
    }
    __syncwarp();
    // dq[h] = scale * sum_j ds[j] * (K[j][h] + P[idx(j)][h]); lanes over h
    for (int h = lane; h < H; h += 32) {
      float ac = 0.f, ap = 0.f;
      const int pbase = T - 1 - t - r_lo;
      for (int j = 0; j < T; ++j) {
        const float d = myDs[j];
        ac = fmaf(d, __bfloat162float(sK[j * S + h]), ac);
        ap = fmaf(d, __bfloat162float(sP[(pbase + j) * S + h]), ap);
      }
      ac *= p.scale;
      ap *= p.scale;
      dqkv[((size_t)b * T + t) * p.ld_qkv + (size_t)n * H + h] = __float2bfloat16(ac + ap);
      atomicAdd(&sAcc[h], ac);
      atomicAdd(&sAcc[H + h], ap);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += kAttThreads) {
    atomicAdd(du + n * H + i, sAcc[i]);
    atomicAdd(dvb + n * H + i, sAcc[H + i]);
  }
}

// ---------------------------------------------------------------------------------------------- backward B (per key / per position)
// columns c in [0, T): key j = c:   dV[j] = sum_t p[t,j] dctx_t ;  dK[j] = scale * sum_t ds[t,j] (q_t + u)
// columns c in [T, T+R): position r = c - T:   dP[r] = scale * sum_t ds[t, r-(T-1)+t] (q_t + v)
// One lane per column, 36 accumulators per lane; q / dctx rows are broadcast from SMEM.
constexpr int kAttColsPerCta = 64;  // 2 warps
__global__ void __launch_bounds__(64) rel_attention_bwd_cols_kernel(const AttParams p, const __nv_bfloat16* __restrict__ probs,
                                                                    const __nv_bfloat16* __restrict__ ds, const __nv_bfloat16* __restrict__ dctx,
                                                                    __nv_bfloat16* __restrict__ dqkv, __nv_bfloat16* __restrict__ dpos) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int b = blockIdx.z, n = blockIdx.y;
  float* sQ = reinterpret_cast<float*>(att_smem);  // [T][H] raw q (bias added per use)
  float* sD = sQ + (size_t)T * H;                  // [T][H] dctx
  for (int i = threadIdx.x; i < T * H; i += blockDim.x) {
    const int t = i / H, h = i % H;
    sQ[i] = __bfloat162float(p.qkv[((size_t)b * T + t) * p.ld_qkv + (size_t)n * H + h]);
    sD[i] = __bfloat162float(dctx[((size_t)b * T + t) * NH + (size_t)n * H + h]);
  }
  __syncthreads();
  const int c = blockIdx.x * kAttColsPerCta + threadIdx.x;
  if (c >= T + R) return;
  const __nv_bfloat16* pbase = probs + ((size_t)b * p.N + n) * T * T;
  const __nv_bfloat16* dbase = ds + ((size_t)b * p.N + n) * T * T;
  if (c < T) {
    const int j = c;
    float av[kMaxH], ak[kMaxH];
#pragma unroll
    for (int h = 0; h < kMaxH; ++h) av[h] = ak[h] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float pp = __bfloat162float(pbase[(size_t)t * T + j]);
      const float dd = __bfloat162float(dbase[(size_t)t * T + j]);
      const float* qr = sQ + t * H;
      const float* dr = sD + t * H;
#pragma unroll
      for (int h = 0; h < kMaxH; ++h) {
        if (h < H) {
          av[h] = fmaf(pp, dr[h], av[h]);
          ak[h] = fmaf(dd, qr[h] + p.ubias[n * H + h], ak[h]);
        }
      }
    }
    __nv_bfloat16* krow = dqkv + ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H;
#pragma unroll
    for (int h = 0; h < kMaxH; ++h) {
      if (h < H) {
        krow[NH + h] = __float2bfloat16(ak[h] * p.scale);
        krow[2 * NH + h] = __float2bfloat16(av[h]);
      }
    }
  } else {
    const int r = c - T;
    float ap[kMaxH];
#pragma unroll
    for (int h = 0; h < kMaxH; ++h) ap[h] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int j = r - (T - 1) + t;
      if (j < 0 || j >= T) continue;
      const float dd = __bfloat162float(dbase[(size_t)t * T + j]);
      const float* qr = sQ + t * H;
#pragma unroll
      for (int h = 0; h < kMaxH; ++h)
        if (h < H) ap[h] = fmaf(dd, qr[h] + p.vbias[n * H + h], ap[h]);
    }
    __nv_bfloat16* prow = dpos + ((size_t)b * R + r) * NH + (size_t)n * H;
#pragma unroll
    for (int h = 0; h < kMaxH; ++h)
      if (h < H) prow[h] = __float2bfloat16(ap[h] * p.scale);
  }
}

static size_t att_rows_smem(int T, int H, bool bwd) {
  const int S = ((H + 2) | 2);
  size_t bytes = (size_t)(2 * T + T + kAttTQ - 1) * S * 2;
  bytes = align_up(bytes, 16);
  bytes += sizeof(float) * 4 * T;
  if (bwd) bytes += sizeof(float) * 2 * H;
  return bytes;
}

static int att_check(const AttParams& p) {
  if (p.B < 0 || p.T < 1 || p.N < 1 || p.H < 2 || (p.H & 1) || p.H > kMaxH || p.R != 2 * p.T - 1 || p.T > 1024) return TFASR_INVALID_VALUE;
  if (!p.qkv || !p.pos || !p.ubias || !p.vbias || (p.ld_qkv & 1) || p.B > 65535) return TFASR_INVALID_VALUE;
  return TFASR_OK;
}

}  // namespace tfasr

using namespace tfasr;

static AttParams make_att(const void* qkv, int ld_qkv, const void* pos, const float* u, const float* v, const int32_t* length, int B, int T, int N,
                          int H, int chunk, int history) {
  AttParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.pos = static_cast<const __nv_bfloat16*>(pos);
  p.ubias = u;
  p.vbias = v;
  p.length = length;
  p.B = B;
  p.T = T;
  p.N = N;
  p.H = H;
  p.R = 2 * T - 1;
  p.ld_qkv = ld_qkv;
  p.chunk = chunk;
  p.history = history;
  p.scale = 1.0f / sqrtf((float)H);
  return p;
}

#define ATT_DISPATCH(KERNEL, ...)                                                                                  \
  do {                                                                                                             \
    const int its = (T + 31) / 32;                                                                                 \
    if (its <= 8) { LAUNCH(KERNEL<8>, __VA_ARGS__); }                                                              \
    else if (its <= 16) { LAUNCH(KERNEL<16>, __VA_ARGS__); }                                                       \
    else if (its <= 24) { LAUNCH(KERNEL<24>, __VA_ARGS__); }                                                       \
    else { LAUNCH(KERNEL<32>, __VA_ARGS__); }                                                                      \
  } while (0)

extern "C" int tfasr_rel_attention_fwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias,
                                       const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx,
                                       void* probs, tfasr_stream_t stream_) {
  AttParams p = make_att(qkv, ld_qkv, pos, ubias, vbias, length, B, T, N, H, chunk, history);
  int st = att_check(p);
  if (st || !ctx || !probs) return st ? st : TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = att_rows_smem(T, H, false);
  if (smem > 220 * 1024) return TFASR_INVALID_VALUE;
  dim3 grid((unsigned)ceil_div(T, kAttTQ), N, B);
#define LAUNCH(K, ...)                                                                                           \
  {                                                                                                              \
    if (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED; \
    K<<<grid, kAttThreads, smem, stream>>>(__VA_ARGS__);                                                         \
  }
  ATT_DISPATCH(rel_attention_fwd_kernel, p, static_cast<__nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(probs));
#undef LAUNCH
  count_launch();
  return check_launch();
}

extern "C" int tfasr_rel_attention_bwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias,
                                       const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history,
                                       const void* probs, const void* dctx, void* ds_ws, void* dqkv, void* dpos, float* du, float* dvb,
                                       tfasr_stream_t stream_) {
  AttParams p = make_att(qkv, ld_qkv, pos, ubias, vbias, length, B, T, N, H, chunk, history);
  int st = att_check(p);
  if (st || !probs || !dctx || !ds_ws || !dqkv || !dpos || !du || !dvb) return st ? st : TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = att_rows_smem(T, H, true);
  if (smem > 220 * 1024) return TFASR_INVALID_VALUE;
  dim3 grid((unsigned)ceil_div(T, kAttTQ), N, B);
#define LAUNCH(K, ...)                                                                                           \
  {                                                                                                              \
    if (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) return TFASR_EXECUTION_FAILED; \
    K<<<grid, kAttThreads, smem, stream>>>(__VA_ARGS__);                                                         \
  }
  ATT_DISPATCH(rel_attention_bwd_rows_kernel, p, static_cast<const __nv_bfloat16*>(probs), static_cast<const __nv_bfloat16*>(dctx),
               static_cast<__nv_bfloat16*>(ds_ws), static_cast<__nv_bfloat16*>(dqkv), du, dvb);
#undef LAUNCH
  count_launch();
  const size_t smem2 = sizeof(float) * 2 * (size_t)T * H;
  if (smem2 > 220 * 1024) return TFASR_INVALID_VALUE;
  if (cudaFuncSetAttribute(rel_attention_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
    return TFASR_EXECUTION_FAILED;
  dim3 grid2((unsigned)ceil_div(T + p.R, kAttColsPerCta), N, B);
  rel_attention_bwd_cols_kernel<<<grid2, kAttColsPerCta, smem2, stream>>>(p, static_cast<const __nv_bfloat16*>(probs),
                                                                          static_cast<const __nv_bfloat16*>(ds_ws),
                                                                          static_cast<const __nv_bfloat16*>(dctx), static_cast<__nv_bfloat16*>(dqkv),
                                                                          static_cast<__nv_bfloat16*>(dpos));
  count_launch();
  return check_launch();
}
