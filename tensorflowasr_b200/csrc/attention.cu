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
#include "attention_mma.cuh"

namespace tfasr {

constexpr int kAttThreads = 128;  // 4 warps
constexpr int kAttTQ = 16;        // query rows per CTA (forward / backward-A)
constexpr int kMaxH = 64;

struct AttParams {
  const __nv_bfloat16* qkv;  // [B*T, ld_qkv]: q at col n*H, k at NH + n*H, v at 2NH + n*H
  const __nv_bfloat16* pos;  // [B*R, NH] with row stride ld_pos (a column slice of the all-blocks projection when ld_pos > NH)
  const float* ubias;        // [N, H]
  const float* vbias;        // [N, H]
  const int32_t* length;     // [B] or null
  int B, T, N, H, R, ld_qkv;
  int ld_pos, ld_dpos;       // row strides (elements) of pos and of the dpos output
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

// The same predicate as a per-row key window, derived ONCE per row: key j is allowed iff (unsigned)(j - lo) < (unsigned)n  (n = 0: the whole
// row is masked).  The tensor-core kernels test it per score: one subtract + one unsigned compare instead of a division by the chunk size.
struct AttWin {
  int lo;
  unsigned n;
};
__device__ __forceinline__ AttWin att_window(int t, int len, int T, int chunk, int history) {
  AttWin w;
  w.lo = 0;
  w.n = 0u;
  if (t >= len) return w;  // query-side mask
  if (chunk > 0) {
    const int c0 = (t / chunk) * chunk;
    const int hist = history < 0 ? T : history;
    const int start = max(0, c0 - hist), end = min(T, c0 + chunk);
    w.lo = start;
    w.n = (unsigned)max(end - start, 0);
    return w;
  }
  w.n = 0x7fffffffu;  // no key-side mask without streaming (padded key columns are handled by the callers)
  return w;
}
__device__ __forceinline__ bool att_in(const AttWin& w, int j) { return (unsigned)(j - w.lo) < w.n; }

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
      if (r < R) v = *reinterpret_cast<const __nv_bfloat162*>(p.pos + ((size_t)b * R + r) * p.ld_pos + (size_t)n * H + 2 * h2);
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
      if (r < R) v = *reinterpret_cast<const __nv_bfloat162*>(p.pos + ((size_t)b * R + r) * p.ld_pos + (size_t)n * H + 2 * h2);
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
    __nv_bfloat16* prow = dpos + ((size_t)b * R + r) * p.ld_dpos + (size_t)n * H;
#pragma unroll
    for (int h = 0; h < kMaxH; ++h)
      if (h < H) prow[h] = __float2bfloat16(ap[h] * p.scale);
  }
}

// ================================================================================================ tensor-core forward
// One CTA (16 warps) = one (utterance, head): all T <= 256 query rows, 16 per warp, so K, V and the projected position table are staged
// once and four warps share every SM scheduler.  Operands live in SMEM as bf16 with a row stride that is an odd multiple of 16 bytes
// (conflict-free ldmatrix).  Each warp walks the keys in chunks of 64, twice (row max / sum, then probabilities), recomputing the
// scores instead of keeping a [16 x T] FP32 row buffer - the MMAs are cheap, SMEM and latency are not:
//   positional block  S_pos[16 x 80] = Qv . Pos[window]^T  -> per-warp FP32 scratch            (mma.sync m16n8k16)
//   content tiles     S[16 x 8] = Qu . K^T + S_pos[i][jl + 15 - i]   (the rel-shift is this read index), masks
//   pass 2            P = exp(S - m) / l  -> bf16 to HBM (kept for backward) and, straight from the accumulator registers, the A operand
//                     of O += P . V   (V through ldmatrix.trans)
template <int KS>  // k-steps of 16 over the head dimension: H <= 16 * KS
struct AttMma {
  static constexpr int HS = KS == 3 ? 56 : (KS == 4 ? 72 : (KS == 2 ? 40 : 24));  // bf16 row stride: odd multiple of 16 B
  static constexpr int CH = 64;                                                    // keys per chunk
  static constexpr int SCW = 84;                                                   // f32 scratch row stride (80 columns used)
  static constexpr int SKW = 88;                                                   // bf16 skew-buffer row stride (backward)
};

__host__ __device__ inline int att_mma_tk(int T) { return (T + 15) / 16 * 16; }
__host__ __device__ inline int att_mma_tc(int T) { return (T + 63) / 64 * 64; }       // keys padded to whole chunks
__host__ __device__ inline int att_mma_prow(int T) { return att_mma_tk(T) + att_mma_tc(T) + 16; }  // staged position-table rows
__host__ __device__ inline int att_mma_ps(int T) { return att_mma_tk(T) + 8; }        // bf16 row stride of the P / dS blocks (columns kernel)

template <int KS>
static size_t att_mma_fwd_smem(int T) {
  using C = AttMma<KS>;
  const size_t kv = (size_t)2 * att_mma_tc(T) * C::HS * 2, pos = (size_t)att_mma_prow(T) * C::HS * 2;
  const size_t q = (size_t)2 * att_mma_tk(T) * C::HS * 2, scratch = (size_t)16 * 16 * C::SCW * 4;
  return kv + pos + (q > scratch ? q : scratch);
}

// 8-byte (4 x bf16) staging copies; H % 4 == 0 keeps every head slice 8-byte aligned (checked by the launcher)
__device__ __forceinline__ uint2 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, uint2 v) { *reinterpret_cast<uint2*>(p) = v; }

// K, V [Tc][HS] and the position window [prow][HS]: window row rr holds table row r = rr - (Tk - T)
template <int KS>
__device__ __forceinline__ void att_stage_kvp(const AttParams& p, int b, int n, __nv_bfloat16* sK, __nv_bfloat16* sV, __nv_bfloat16* sP,
                                              bool want_v) {
  constexpr int HS = AttMma<KS>::HS, HP4 = KS * 4;
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int Tk = att_mma_tk(T), Tc = att_mma_tc(T), PR = att_mma_prow(T);
  const uint2 z = make_uint2(0u, 0u);
  // valid elements travel as 8-byte cp.async copies (all in flight at once; the caller waits with att_stage_wait()), padding is zeroed
  for (int i = threadIdx.x; i < Tc * HP4; i += blockDim.x) {
    const int j = i / HP4, h4 = i % HP4;
    if (j < T && 4 * h4 < H) {
      const __nv_bfloat16* g = p.qkv + ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H + 4 * h4;
      mma::cp_async8(sK + j * HS + 4 * h4, g + NH);
      if (want_v) mma::cp_async8(sV + j * HS + 4 * h4, g + 2 * NH);
    } else {
      st8(sK + j * HS + 4 * h4, z);
      if (want_v) st8(sV + j * HS + 4 * h4, z);
    }
  }
  for (int i = threadIdx.x; i < PR * HP4; i += blockDim.x) {
    const int rr = i / HP4, h4 = i % HP4;
    const int r = rr - (Tk - T);
    if (r >= 0 && r < R && 4 * h4 < H) mma::cp_async8(sP + rr * HS + 4 * h4, p.pos + ((size_t)b * R + r) * p.ld_pos + (size_t)n * H + 4 * h4);
    else st8(sP + rr * HS + 4 * h4, z);
  }
  mma::cp_async_commit();
}

__device__ __forceinline__ void att_stage_wait() { mma::cp_async_wait<0>(); }

// (q + u) * scale and (q + v) * scale rows [Tk][HS], zero beyond T / H
template <int KS>
__device__ __forceinline__ void att_stage_q(const AttParams& p, int b, int n, __nv_bfloat16* sQu, __nv_bfloat16* sQv) {
  constexpr int HS = AttMma<KS>::HS, HP2 = KS * 8;
  const int T = p.T, H = p.H, Tk = att_mma_tk(T);
  const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
  for (int i = threadIdx.x; i < Tk * HP2; i += blockDim.x) {
    const int t = i / HP2, h2 = i % HP2;
    __nv_bfloat162 qu = z2, qv = z2;
    if (t < T && 2 * h2 < H) {
      const __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + ((size_t)b * T + t) * p.ld_qkv + (size_t)n * H + 2 * h2);
      const float q0 = __low2float(q), q1 = __high2float(q);
      const float* ub = p.ubias + n * H + 2 * h2;
      const float* vb = p.vbias + n * H + 2 * h2;
      qu = __floats2bfloat162_rn((q0 + ub[0]) * p.scale, (q1 + ub[1]) * p.scale);
      qv = __floats2bfloat162_rn((q0 + vb[0]) * p.scale, (q1 + vb[1]) * p.scale);
    }
    *reinterpret_cast<__nv_bfloat162*>(sQu + t * HS + 2 * h2) = qu;
    *reinterpret_cast<__nv_bfloat162*>(sQv + t * HS + 2 * h2) = qv;
  }
}

// Scores of this warp's 16 rows against keys [c0, c0 + 64): s[nt][0..1] row g, s[nt][2..3] row g + 8, columns c0 + 8 nt + 2 tq (+1).
template <int KS>
__device__ __forceinline__ void att_chunk_scores(float (&s)[8][4], const uint32_t (&aqu)[KS][4], const uint32_t (&aqv)[KS][4],
                                                 const __nv_bfloat16* sK, const __nv_bfloat16* sP, float* scratch, int rr0, int c0, int lane,
                                                 const AttWin& wA, const AttWin& wB, const AttParams& p) {
  constexpr int HS = AttMma<KS>::HS, SCW = AttMma<KS>::SCW;
  const int g = lane >> 2, tq = lane & 3;
  const int brow = lane & 7, bcol = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int nt = 0; nt < 10; ++nt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t bfr[2];
      mma::ldsm_x2(bfr, sP + (rr0 + c0 + nt * 8 + brow) * HS + ks * 16 + bcol);
      mma::mma_16816(acc, aqv[ks], bfr);
    }
    *reinterpret_cast<float2*>(scratch + g * SCW + nt * 8 + 2 * tq) = make_float2(acc[0], acc[1]);
    *reinterpret_cast<float2*>(scratch + (g + 8) * SCW + nt * 8 + 2 * tq) = make_float2(acc[2], acc[3]);
  }
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t bfr[2];
      mma::ldsm_x2(bfr, sK + (c0 + nt * 8 + brow) * HS + ks * 16 + bcol);
      mma::mma_16816(acc, aqu[ks], bfr);
    }
    const int jl = nt * 8 + 2 * tq, j = c0 + jl;
    const float* ra = scratch + g * SCW + jl + 15 - g;
    const float* rb = scratch + (g + 8) * SCW + jl + 7 - g;
    float s0 = acc[0] + ra[0], s1 = acc[1] + ra[1], s2 = acc[2] + rb[0], s3 = acc[3] + rb[1];
    // masked_fill(-1e9) of general.py:30-41; padded key columns do not exist (-inf)
    if (!att_in(wA, j)) s0 = -1e9f;
    if (!att_in(wA, j + 1)) s1 = -1e9f;
    if (!att_in(wB, j)) s2 = -1e9f;
    if (!att_in(wB, j + 1)) s3 = -1e9f;
    if (j >= p.T) s0 = s2 = -CUDART_INF_F;
    if (j + 1 >= p.T) s1 = s3 = -CUDART_INF_F;
    s[nt][0] = s0;
    s[nt][1] = s1;
    s[nt][2] = s2;
    s[nt][3] = s3;
  }
  __syncwarp();  // the scratch is rewritten by the next chunk
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <int KS>
__global__ void __launch_bounds__(512, 1) rel_attention_fwd_mma_kernel(const AttParams p, __nv_bfloat16* __restrict__ ctx,
                                                                       __nv_bfloat16* __restrict__ probs) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  using C = AttMma<KS>;
  constexpr int HS = C::HS;
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H;
  const int b = blockIdx.y, n = blockIdx.x;
  const int Tk = att_mma_tk(T), Tc = att_mma_tc(T);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + (size_t)Tc * HS;
  __nv_bfloat16* sP = sV + (size_t)Tc * HS;
  uint8_t* U = reinterpret_cast<uint8_t*>(sP + (size_t)att_mma_prow(T) * HS);
  __nv_bfloat16* sQu = reinterpret_cast<__nv_bfloat16*>(U);
  __nv_bfloat16* sQv = sQu + (size_t)Tk * HS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int len = p.length ? p.length[b] : T;
  att_stage_kvp<KS>(p, b, n, sK, sV, sP, true);
  att_stage_q<KS>(p, b, n, sQu, sQv);
  att_stage_wait();
  __syncthreads();
  const int tw = warp * 16;
  const bool active = tw < T;
  uint32_t aqu[KS][4], aqv[KS][4];
  if (active) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      mma::ldsm_x4(aqu[ks], sQu + (tw + (lane & 15)) * HS + ks * 16 + (lane >> 4) * 8);
      mma::ldsm_x4(aqv[ks], sQv + (tw + (lane & 15)) * HS + ks * 16 + (lane >> 4) * 8);
    }
  }
  __syncthreads();  // every warp holds its Q fragments: the Q region becomes the per-warp scratch
  if (!active) return;
  float* scratch = reinterpret_cast<float*>(U) + (size_t)warp * 16 * C::SCW;
  const int rr0 = Tk - 16 - tw;
  const int tA = tw + g, tB = tw + g + 8;
  const AttWin wA = att_window(tA, len, T, p.chunk, p.history), wB = att_window(tB, len, T, p.chunk, p.history);
  // ---- pass 1: row max and sum
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
  for (int c0 = 0; c0 < Tc; c0 += C::CH) {
    float s[8][4];
    att_chunk_scores<KS>(s, aqu, aqv, sK, sP, scratch, rr0, c0, lane, wA, wB, p);
    float x0 = -CUDART_INF_F, x1 = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      x0 = fmaxf(x0, fmaxf(s[nt][0], s[nt][1]));
      x1 = fmaxf(x1, fmaxf(s[nt][2], s[nt][3]));
    }
    const float n0 = fmaxf(m0, quad_max(x0)), n1 = fmaxf(m1, quad_max(x1));  // finite from the first chunk on (key 0 exists)
    // a term is FADD + FMUL + ex2.approx.ftz instead of __expf's six instructions (the exponentials were a quarter of this kernel's issue
    // slots, profiles/attention_ncu_r01.txt).  The subtraction stays a separate, exact step: a fully masked row has s = m = -1e9, and folding
    // m into an FFMA would leave the rounding error of 1e9 * log2(e) in the exponent
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      a0 += fast_ex2((s[nt][0] - n0) * kLog2e) + fast_ex2((s[nt][1] - n0) * kLog2e);
      a1 += fast_ex2((s[nt][2] - n1) * kLog2e) + fast_ex2((s[nt][3] - n1) * kLog2e);
    }
    l0 = l0 * fast_ex2((m0 - n0) * kLog2e) + a0;  // m = -inf on the first chunk: 2^-inf = 0
    l1 = l1 * fast_ex2((m1 - n1) * kLog2e) + a1;
    m0 = n0;
    m1 = n1;
  }
  // p = 2^((s - m) log2e - log2 l): the normaliser rides in the exponent
  const float e0 = -log2f(quad_sum(l0)), e1 = -log2f(quad_sum(l1));
  // ---- pass 2: probabilities and O = P . V
  float o[KS * 2][4];
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) o[nh][0] = o[nh][1] = o[nh][2] = o[nh][3] = 0.f;
  __nv_bfloat16* prA = probs + (((size_t)b * p.N + n) * T + min(tA, T - 1)) * T;
  __nv_bfloat16* prB = probs + (((size_t)b * p.N + n) * T + min(tB, T - 1)) * T;
  const bool even = (T & 1) == 0;
  for (int c0 = 0; c0 < Tc; c0 += C::CH) {
    float s[8][4];
    att_chunk_scores<KS>(s, aqu, aqv, sK, sP, scratch, rr0, c0, lane, wA, wB, p);
    uint32_t pa[8], pb[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      pa[nt] = pack_bf16(fast_ex2(fmaf(s[nt][0] - m0, kLog2e, e0)), fast_ex2(fmaf(s[nt][1] - m0, kLog2e, e0)));
      pb[nt] = pack_bf16(fast_ex2(fmaf(s[nt][2] - m1, kLog2e, e1)), fast_ex2(fmaf(s[nt][3] - m1, kLog2e, e1)));
      const int j = c0 + nt * 8 + 2 * tq;
      if (even) {
        if (j < T) {
          if (tA < T) *reinterpret_cast<uint32_t*>(prA + j) = pa[nt];
          if (tB < T) *reinterpret_cast<uint32_t*>(prB + j) = pb[nt];
        }
      } else {
        const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(&pa[nt]), vb = *reinterpret_cast<const __nv_bfloat162*>(&pb[nt]);
        if (tA < T) {
          if (j < T) prA[j] = __low2bfloat16(va);
          if (j + 1 < T) prA[j + 1] = __high2bfloat16(va);
        }
        if (tB < T) {
          if (j < T) prB[j] = __low2bfloat16(vb);
          if (j + 1 < T) prB[j + 1] = __high2bfloat16(vb);
        }
      }
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const uint32_t afr[4] = {pa[2 * kt], pb[2 * kt], pa[2 * kt + 1], pb[2 * kt + 1]};
#pragma unroll
      for (int nh = 0; nh < KS * 2; ++nh) {
        uint32_t bfr[2];
        mma::ldsm_x2_trans(bfr, sV + (c0 + kt * 16 + (lane & 15)) * HS + nh * 8);
        mma::mma_16816(o[nh], afr, bfr);
      }
    }
  }
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) {
    const int h = nh * 8 + 2 * tq;
    if (h < H) {
      if (tA < T) *reinterpret_cast<__nv_bfloat162*>(ctx + ((size_t)b * T + tA) * NH + (size_t)n * H + h) = __floats2bfloat162_rn(o[nh][0], o[nh][1]);
      if (tB < T) *reinterpret_cast<__nv_bfloat162*>(ctx + ((size_t)b * T + tB) * NH + (size_t)n * H + h) = __floats2bfloat162_rn(o[nh][2], o[nh][3]);
    }
  }
}

template <int KS>
static int launch_att_fwd_mma(const AttParams& p, __nv_bfloat16* ctx, __nv_bfloat16* probs, cudaStream_t stream) {
  const size_t smem = att_mma_fwd_smem<KS>(p.T);
  if (smem > 227 * 1024 || p.T > 256 || (p.H & 3) || (p.ld_qkv & 3) || (p.ld_pos & 3) || (reinterpret_cast<uintptr_t>(p.pos) & 7)) return -1;  // caller falls back to the FP32 kernel
  if (cudaFuncSetAttribute(rel_attention_fwd_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return TFASR_EXECUTION_FAILED;
  dim3 grid(p.N, p.B);
  launch_pdl(rel_attention_fwd_mma_kernel<KS>, dim3(grid), dim3(512), smem, stream, p, ctx, probs);
  count_launch();
  return check_launch();
}

// ================================================================================================ tensor-core backward, rows
// Same CTA shape as the forward.  Per warp (16 query rows), one pass over the keys in chunks of 64:
//   delta_i = dO_i . O_i   (= sum_j P_ij dP_ij, from the forward's output instead of a second sweep)
//   dP = dO . V^T  ->  dS = P (dP - delta), masked -> 0;   bf16 dS -> HBM (for the column kernel), kept in registers as the A operand of
//   dQ += dS . K, and written skewed (column jl + 15 - i) to a per-warp SMEM buffer for  dQ += dS_skewed . Pos[window]
//   du / dvb = column sums of the two dQ parts
template <int KS>
static size_t att_mma_bwd_rows_smem(int T) {
  using C = AttMma<KS>;
  const size_t kv = (size_t)2 * att_mma_tc(T) * C::HS * 2, pos = (size_t)att_mma_prow(T) * C::HS * 2;
  const size_t dox = (size_t)att_mma_tk(T) * C::HS * 2, skew = (size_t)16 * 16 * C::SKW * 2;
  return kv + pos + (dox > skew ? dox : skew) + 2 * 64 * sizeof(float);
}

template <int KS>
__global__ void __launch_bounds__(512, 1) rel_attention_bwd_rows_mma_kernel(const AttParams p, const __nv_bfloat16* __restrict__ probs,
                                                                            const __nv_bfloat16* __restrict__ dctx,
                                                                            const __nv_bfloat16* __restrict__ octx,
                                                                            __nv_bfloat16* __restrict__ ds_out, __nv_bfloat16* __restrict__ dqkv,
                                                                            float* __restrict__ du, float* __restrict__ dvb) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  using C = AttMma<KS>;
  constexpr int HS = C::HS, SKW = C::SKW;
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H;
  const int b = blockIdx.y, n = blockIdx.x;
  const int Tk = att_mma_tk(T), Tc = att_mma_tc(T);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + (size_t)Tc * HS;
  __nv_bfloat16* sP = sV + (size_t)Tc * HS;
  uint8_t* U = reinterpret_cast<uint8_t*>(sP + (size_t)att_mma_prow(T) * HS);
  const size_t usz = max((size_t)Tk * HS * 2, (size_t)16 * 16 * SKW * 2);
  float* sAcc = reinterpret_cast<float*>(U + usz);
  __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(U);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int len = p.length ? p.length[b] : T;
  att_stage_kvp<KS>(p, b, n, sK, sV, sP, true);
  {
    constexpr int HP4 = KS * 4;
    const uint2 z = make_uint2(0u, 0u);
    for (int i = threadIdx.x; i < Tk * HP4; i += 512) {
      const int t = i / HP4, h4 = i % HP4;
      if (t < T && 4 * h4 < H) mma::cp_async8(sdO + t * HS + 4 * h4, dctx + ((size_t)b * T + t) * NH + (size_t)n * H + 4 * h4);
      else st8(sdO + t * HS + 4 * h4, z);
    }
    mma::cp_async_commit();
    if (threadIdx.x < 128) sAcc[threadIdx.x] = 0.f;
  }
  att_stage_wait();
  __syncthreads();
  const int tw = warp * 16;
  const bool active = tw < T;
  uint32_t ado[KS][4];
  if (active) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) mma::ldsm_x4(ado[ks], sdO + (tw + (lane & 15)) * HS + ks * 16 + (lane >> 4) * 8);
  }
  __syncthreads();  // dO fragments are in registers: the region becomes the per-warp skew buffers
  float accC[KS * 2][4], accP[KS * 2][4];
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) {
    accC[nh][0] = accC[nh][1] = accC[nh][2] = accC[nh][3] = 0.f;
    accP[nh][0] = accP[nh][1] = accP[nh][2] = accP[nh][3] = 0.f;
  }
  if (active) {
    __nv_bfloat16* skew = reinterpret_cast<__nv_bfloat16*>(U) + (size_t)warp * 16 * SKW;
    for (int i = lane; i < 16 * SKW / 2; i += 32) reinterpret_cast<uint32_t*>(skew)[i] = 0u;  // outside the band stays zero
    const int rr0 = Tk - 16 - tw;
    const int tA = tw + g, tB = tw + g + 8;
    const int rA = min(tA, T - 1), rB = min(tB, T - 1);
    // rows past the sequence (tA >= T) have an empty window: len <= T, so att_window masks them through its query-side test
    const AttWin wA = att_window(tA, min(len, T), T, p.chunk, p.history), wB = att_window(tB, min(len, T), T, p.chunk, p.history);
    // delta = dO . O over the head dimension (quad-split)
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nh = 0; nh < KS * 2; ++nh) {
      const int h = nh * 8 + 2 * tq;
      if (h < H) {
        const size_t ia = ((size_t)b * T + rA) * NH + (size_t)n * H + h, ib = ((size_t)b * T + rB) * NH + (size_t)n * H + h;
        const __nv_bfloat162 xa = *reinterpret_cast<const __nv_bfloat162*>(dctx + ia), ya = *reinterpret_cast<const __nv_bfloat162*>(octx + ia);
        const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(dctx + ib), yb = *reinterpret_cast<const __nv_bfloat162*>(octx + ib);
        d0 += __low2float(xa) * __low2float(ya) + __high2float(xa) * __high2float(ya);
        d1 += __low2float(xb) * __low2float(yb) + __high2float(xb) * __high2float(yb);
      }
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
    const __nv_bfloat16* prA = probs + (((size_t)b * p.N + n) * T + rA) * T;
    const __nv_bfloat16* prB = probs + (((size_t)b * p.N + n) * T + rB) * T;
    __nv_bfloat16* dsA = ds_out + (((size_t)b * p.N + n) * T + rA) * T;
    __nv_bfloat16* dsB = ds_out + (((size_t)b * p.N + n) * T + rB) * T;
    const bool even = (T & 1) == 0;
    const int brow = lane & 7, bcol = ((lane >> 3) & 1) * 8;
    __syncwarp();
    for (int c0 = 0; c0 < Tc; c0 += C::CH) {
      // this chunk's probabilities (issued first: their latency hides behind the dP MMAs)
      float2 pA[8], pB[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = c0 + nt * 8 + 2 * tq;
        pA[nt] = pB[nt] = make_float2(0.f, 0.f);
        if (even) {
          if (j < T) {
            pA[nt] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(prA + j));
            pB[nt] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(prB + j));
          }
        } else {
          if (j < T) {
            pA[nt].x = __bfloat162float(prA[j]);
            pB[nt].x = __bfloat162float(prB[j]);
          }
          if (j + 1 < T) {
            pA[nt].y = __bfloat162float(prA[j + 1]);
            pB[nt].y = __bfloat162float(prB[j + 1]);
          }
        }
      }
      uint32_t da[8], db[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bfr[2];
          mma::ldsm_x2(bfr, sV + (c0 + nt * 8 + brow) * HS + ks * 16 + bcol);
          mma::mma_16816(acc, ado[ks], bfr);
        }
        const int jl = nt * 8 + 2 * tq, j = c0 + jl;
        float s0 = pA[nt].x * (acc[0] - d0), s1 = pA[nt].y * (acc[1] - d0);
        float s2 = pB[nt].x * (acc[2] - d1), s3 = pB[nt].y * (acc[3] - d1);
        if (!att_in(wA, j)) s0 = 0.f;  // masked scores are constants
        if (!att_in(wA, j + 1)) s1 = 0.f;
        if (!att_in(wB, j)) s2 = 0.f;
        if (!att_in(wB, j + 1)) s3 = 0.f;
        da[nt] = pack_bf16(s0, s1);
        db[nt] = pack_bf16(s2, s3);
        const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(&da[nt]), vb = *reinterpret_cast<const __nv_bfloat162*>(&db[nt]);
        skew[g * SKW + jl + 15 - g] = __low2bfloat16(va);
        skew[g * SKW + jl + 16 - g] = __high2bfloat16(va);
        skew[(g + 8) * SKW + jl + 7 - g] = __low2bfloat16(vb);
        skew[(g + 8) * SKW + jl + 8 - g] = __high2bfloat16(vb);
        if (even) {
          if (j < T) {
            if (tA < T) *reinterpret_cast<uint32_t*>(dsA + j) = da[nt];
            if (tB < T) *reinterpret_cast<uint32_t*>(dsB + j) = db[nt];
          }
        } else {
          if (tA < T) {
            if (j < T) dsA[j] = __low2bfloat16(va);
            if (j + 1 < T) dsA[j + 1] = __high2bfloat16(va);
          }
          if (tB < T) {
            if (j < T) dsB[j] = __low2bfloat16(vb);
            if (j + 1 < T) dsB[j + 1] = __high2bfloat16(vb);
          }
        }
      }
      __syncwarp();
      // dQ (content): dS . K, A operand straight from the registers
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const uint32_t afr[4] = {da[2 * kt], db[2 * kt], da[2 * kt + 1], db[2 * kt + 1]};
#pragma unroll
        for (int nh = 0; nh < KS * 2; ++nh) {
          uint32_t bfr[2];
          mma::ldsm_x2_trans(bfr, sK + (c0 + kt * 16 + (lane & 15)) * HS + nh * 8);
          mma::mma_16816(accC[nh], afr, bfr);
        }
      }
      // dQ (positional): dS_skewed [16 x 80] . Pos[window rows rr0 + c0 ..]
#pragma unroll
      for (int kt = 0; kt < 5; ++kt) {
        uint32_t afr[4];
        mma::ldsm_x4(afr, skew + (lane & 15) * SKW + kt * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int nh = 0; nh < KS * 2; ++nh) {
          uint32_t bfr[2];
          mma::ldsm_x2_trans(bfr, sP + (rr0 + c0 + kt * 16 + (lane & 15)) * HS + nh * 8);
          mma::mma_16816(accP[nh], afr, bfr);
        }
      }
      __syncwarp();  // the skew buffer is rewritten by the next chunk
    }
    // scores came from (q + u) * scale and (q + v) * scale: dq carries the scale once
#pragma unroll
    for (int nh = 0; nh < KS * 2; ++nh) {
      const int h = nh * 8 + 2 * tq;
      if (h < H) {
        if (tA < T)
          *reinterpret_cast<__nv_bfloat162*>(dqkv + ((size_t)b * T + tA) * p.ld_qkv + (size_t)n * H + h) =
              __floats2bfloat162_rn((accC[nh][0] + accP[nh][0]) * p.scale, (accC[nh][1] + accP[nh][1]) * p.scale);
        if (tB < T)
          *reinterpret_cast<__nv_bfloat162*>(dqkv + ((size_t)b * T + tB) * p.ld_qkv + (size_t)n * H + h) =
              __floats2bfloat162_rn((accC[nh][2] + accP[nh][2]) * p.scale, (accC[nh][3] + accP[nh][3]) * p.scale);
      }
    }
  }
  // ---- du / dvb: column sums over the CTA's rows (rows >= T contribute exact zeros)
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float c = accC[nh][e] + accC[nh][e + 2], q = accP[nh][e] + accP[nh][e + 2];
#pragma unroll
      for (int m = 4; m <= 16; m <<= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, m);
        q += __shfl_xor_sync(0xffffffffu, q, m);
      }
      const int h = nh * 8 + 2 * tq + e;
      if (g == 0 && h < H) {
        atomicAdd(&sAcc[h], c * p.scale);
        atomicAdd(&sAcc[64 + h], q * p.scale);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += 512) {
    atomicAdd(du + n * H + i, sAcc[i]);
    atomicAdd(dvb + n * H + i, sAcc[64 + i]);
  }
}

// ================================================================================================ tensor-core backward, columns
// One CTA (16 warps) per (utterance, head) streams P and dS once, 16 query rows at a time, double-buffered in SMEM:
//   dV[j]  += P^T  . dO        warp w owns keys 16w .. 16w+15        (A = P block through ldmatrix.trans)
//   dK[j]  += dS^T . (q+u)s
//   dPos[r] += dS_skew^T . (q+v)s   with r = 16 (q0 - blk) + c',  c' = j + 15 - i + rem:   the diagonal gather of the rel-shift becomes
//                                   a block-constant row offset, so warp w accumulates position rows 32w .. 32w+31 in registers.
template <int KS>
struct AttCols {
  static constexpr int HS = AttMma<KS>::HS;
  static constexpr int THREADS = 512;
};
__host__ __device__ inline int att_cols_qs(int T) { return att_mma_tk(T) + 32 + 8; }

template <int KS>
static size_t att_mma_bwd_cols_smem(int T) {
  return (size_t)2 * 16 * (2 * att_mma_ps(T) + att_cols_qs(T) + 3 * AttMma<KS>::HS) * 2;
}

template <int KS>
__global__ void __launch_bounds__(512) rel_attention_bwd_cols_mma_kernel(const AttParams p, const __nv_bfloat16* __restrict__ probs,
                                                                         const __nv_bfloat16* __restrict__ ds,
                                                                         const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dqkv,
                                                                         __nv_bfloat16* __restrict__ dpos) {
  constexpr int HS = AttMma<KS>::HS;
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int b = blockIdx.y, n = blockIdx.x;
  const int Tk = att_mma_tk(T), PS = att_mma_ps(T), QS = att_cols_qs(T);
  const int stage_elems = 16 * (2 * PS + QS + 3 * HS);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int q0 = (T - 16) / 16, rem = (T - 16) % 16;  // T >= 16
  const int nct = (rem + T + 15 + 15) / 16;           // 16-wide tiles of the skewed coordinate c'
  const int nblk = Tk / 16;
  const __nv_bfloat16* pbase = probs + ((size_t)b * p.N + n) * T * T;
  const __nv_bfloat16* dbase = ds + ((size_t)b * p.N + n) * T * T;
  const __nv_bfloat16 z1 = __float2bfloat16(0.f);
  const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);

  auto stage = [&](int blk, int buf) {
    __nv_bfloat16* sPb = reinterpret_cast<__nv_bfloat16*>(att_smem) + (size_t)buf * stage_elems;
    __nv_bfloat16* sDb = sPb + 16 * PS;
    __nv_bfloat16* sSk = sDb + 16 * PS;
    __nv_bfloat16* sQu = sSk + 16 * QS;
    __nv_bfloat16* sQv = sQu + 16 * HS;
    __nv_bfloat16* sdO = sQv + 16 * HS;
    const int tb = blk * 16;
    if ((T & 1) == 0) {
      const int PS2 = PS / 2;
      for (int i = threadIdx.x; i < 16 * PS2; i += 512) {
        const int r = i / PS2, j2 = i % PS2;
        const int t = tb + r;
        __nv_bfloat162 a = z2, d = z2;
        if (t < T && 2 * j2 < T) {
          a = *reinterpret_cast<const __nv_bfloat162*>(pbase + (size_t)t * T + 2 * j2);
          d = *reinterpret_cast<const __nv_bfloat162*>(dbase + (size_t)t * T + 2 * j2);
        }
        *reinterpret_cast<__nv_bfloat162*>(sPb + r * PS + 2 * j2) = a;
        *reinterpret_cast<__nv_bfloat162*>(sDb + r * PS + 2 * j2) = d;
      }
    } else {
      for (int i = threadIdx.x; i < 16 * PS; i += 512) {
        const int r = i / PS, j = i % PS;
        const int t = tb + r;
        const bool ok = t < T && j < T;
        sPb[r * PS + j] = ok ? pbase[(size_t)t * T + j] : z1;
        sDb[r * PS + j] = ok ? dbase[(size_t)t * T + j] : z1;
      }
    }
    for (int i = threadIdx.x; i < 16 * QS; i += 512) {
      const int r = i / QS, c = i % QS;
      const int t = tb + r, j = c - 15 + r - rem;
      sSk[r * QS + c] = (t < T && j >= 0 && j < T) ? dbase[(size_t)t * T + j] : z1;
    }
    constexpr int HP2 = KS * 8;
    for (int i = threadIdx.x; i < 16 * HP2; i += 512) {
      const int r = i / HP2, h2 = i % HP2;
      const int t = tb + r;
      __nv_bfloat162 qu = z2, qv = z2, dd = z2;
      if (t < T && 2 * h2 < H) {
        const __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + ((size_t)b * T + t) * p.ld_qkv + (size_t)n * H + 2 * h2);
        const float qa = __low2float(q), qb = __high2float(q);
        const float* ub = p.ubias + n * H + 2 * h2;
        const float* vb = p.vbias + n * H + 2 * h2;
        qu = __floats2bfloat162_rn((qa + ub[0]) * p.scale, (qb + ub[1]) * p.scale);
        qv = __floats2bfloat162_rn((qa + vb[0]) * p.scale, (qb + vb[1]) * p.scale);
        dd = *reinterpret_cast<const __nv_bfloat162*>(dctx + ((size_t)b * T + t) * NH + (size_t)n * H + 2 * h2);
      }
      *reinterpret_cast<__nv_bfloat162*>(sQu + r * HS + 2 * h2) = qu;
      *reinterpret_cast<__nv_bfloat162*>(sQv + r * HS + 2 * h2) = qv;
      *reinterpret_cast<__nv_bfloat162*>(sdO + r * HS + 2 * h2) = dd;
    }
  };

  float aV[KS * 2][4], aK[KS * 2][4], aP[2][KS * 2][4];
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh)
#pragma unroll
    for (int e = 0; e < 4; ++e) aV[nh][e] = aK[nh][e] = aP[0][nh][e] = aP[1][nh][e] = 0.f;

  stage(0, 0);
  __syncthreads();
  const int ak = (lane & 7) + (lane >> 4) * 8;   // ldmatrix.trans A: source row (k)
  const int am = ((lane >> 3) & 1) * 8;          // and column offset (m)
  for (int blk = 0; blk < nblk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < nblk) stage(blk + 1, buf ^ 1);
    const __nv_bfloat16* sPb = reinterpret_cast<const __nv_bfloat16*>(att_smem) + (size_t)buf * stage_elems;
    const __nv_bfloat16* sDb = sPb + 16 * PS;
    const __nv_bfloat16* sSk = sDb + 16 * PS;
    const __nv_bfloat16* sQu = sSk + 16 * QS;
    const __nv_bfloat16* sQv = sQu + 16 * HS;
    const __nv_bfloat16* sdO = sQv + 16 * HS;
    if (warp < nblk) {
      uint32_t afp[4], afd[4];
      mma::ldsm_x4_trans(afp, sPb + ak * PS + warp * 16 + am);
      mma::ldsm_x4_trans(afd, sDb + ak * PS + warp * 16 + am);
#pragma unroll
      for (int nh = 0; nh < KS * 2; ++nh) {
        uint32_t bo[2], bq[2];
        mma::ldsm_x2_trans(bo, sdO + (lane & 15) * HS + nh * 8);
        mma::ldsm_x2_trans(bq, sQu + (lane & 15) * HS + nh * 8);
        mma::mma_16816(aV[nh], afp, bo);
        mma::mma_16816(aK[nh], afd, bq);
      }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int ct = (2 * warp + mi) - q0 + blk;  // skewed-coordinate tile feeding position rows 16 (2w + mi) ..
      if (ct >= 0 && ct < nct) {
        uint32_t afs[4];
        mma::ldsm_x4_trans(afs, sSk + ak * QS + ct * 16 + am);
#pragma unroll
        for (int nh = 0; nh < KS * 2; ++nh) {
          uint32_t bq[2];
          mma::ldsm_x2_trans(bq, sQv + (lane & 15) * HS + nh * 8);
          mma::mma_16816(aP[mi][nh], afs, bq);
        }
      }
    }
    __syncthreads();
  }
  // ---- outputs
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) {
    const int h = nh * 8 + 2 * tq;
    if (h >= H) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = warp * 16 + g + half * 8;
      if (j < T) {
        __nv_bfloat16* row = dqkv + ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H + h;
        *reinterpret_cast<__nv_bfloat162*>(row + NH) = __floats2bfloat162_rn(aK[nh][2 * half], aK[nh][2 * half + 1]);
        *reinterpret_cast<__nv_bfloat162*>(row + 2 * NH) = __floats2bfloat162_rn(aV[nh][2 * half], aV[nh][2 * half + 1]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = (2 * warp + mi) * 16 + g + half * 8;
        if (r < R)
          *reinterpret_cast<__nv_bfloat162*>(dpos + ((size_t)b * R + r) * p.ld_dpos + (size_t)n * H + h) =
              __floats2bfloat162_rn(aP[mi][nh][2 * half], aP[mi][nh][2 * half + 1]);
      }
    }
  }
}

// Pipelined variant of the column kernel for even T (probability rows are 4-byte aligned): the P and dS blocks of the next two
// 16-row steps are always in flight as cp.async copies (three SMEM stages), the skewed copy is built SMEM -> SMEM, and the (q+u), (q+v),
// dO rows of the next step wait in registers - the global-load latency that dominated the first version (16 exposed round trips per
// CTA) is off the critical path.
template <int KS>
static size_t att_mma_bwd_cols2_smem(int T) {
  return (size_t)(3 * 2 * 16 * att_mma_ps(T) + 16 * att_cols_qs(T) + 2 * 3 * 16 * AttMma<KS>::HS) * 2;
}

template <int KS>
__global__ void __launch_bounds__(512) rel_attention_bwd_cols_mma2_kernel(const AttParams p, const __nv_bfloat16* __restrict__ probs,
                                                                          const __nv_bfloat16* __restrict__ ds,
                                                                          const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dqkv,
                                                                          __nv_bfloat16* __restrict__ dpos) {
  grid_dependency_wait();  // launched through launch_pdl: nothing global before this line
  constexpr int HS = AttMma<KS>::HS;
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int T = p.T, H = p.H, NH = p.N * p.H, R = p.R;
  const int b = blockIdx.y, n = blockIdx.x;
  const int Tk = att_mma_tk(T), PS = att_mma_ps(T), QS = att_cols_qs(T);
  const int T2 = T >> 1;
  __nv_bfloat16* sRaw = reinterpret_cast<__nv_bfloat16*>(att_smem);  // [3 stages][P block | dS block][16][PS]
  __nv_bfloat16* sSk = sRaw + (size_t)3 * 2 * 16 * PS;                // [16][QS]
  __nv_bfloat16* sQ = sSk + (size_t)16 * QS;                          // [2][Qu | Qv | dO][16][HS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int q0 = (T - 16) / 16, rem = (T - 16) % 16;  // T >= 16
  const int nct = (rem + T + 15 + 15) / 16;
  const int nblk = Tk / 16;
  const __nv_bfloat16* pbase = probs + ((size_t)b * p.N + n) * T * T;
  const __nv_bfloat16* dbase = ds + ((size_t)b * p.N + n) * T * T;
  const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
  // zero everything once: padding columns / rows are never written by the async copies
  for (int i = threadIdx.x; i < (3 * 2 * 16 * PS + 16 * QS + 2 * 3 * 16 * HS) / 2; i += 512) reinterpret_cast<__nv_bfloat162*>(att_smem)[i] = z2;
  __syncthreads();

  auto issue = [&](int blk) {  // P and dS rows of step blk -> stage blk % 3; warp w copies row w (no index divisions in the loop)
    __nv_bfloat16* dP = sRaw + (size_t)(blk % 3) * 2 * 16 * PS + warp * PS;
    __nv_bfloat16* dD = dP + 16 * PS;
    const int t = blk * 16 + warp;
    if (t < T) {
      const __nv_bfloat16* gp = pbase + (size_t)t * T;
      const __nv_bfloat16* gd = dbase + (size_t)t * T;
      for (int j2 = lane; j2 < T2; j2 += 32) {
        mma::cp_async4(dP + 2 * j2, gp + 2 * j2);
        mma::cp_async4(dD + 2 * j2, gd + 2 * j2);
      }
    } else {  // rows beyond T (last step only): the stage may hold an earlier step's rows
      for (int j2 = lane; j2 < T2; j2 += 32) {
        *reinterpret_cast<__nv_bfloat162*>(dP + 2 * j2) = z2;
        *reinterpret_cast<__nv_bfloat162*>(dD + 2 * j2) = z2;
      }
    }
    mma::cp_async_commit();
  };
  // (q+u) s, (q+v) s, dO of one step: 16 rows x KS*8 column pairs <= 512 items: one per thread, prefetched into registers
  constexpr int HP2 = KS * 8;
  const bool qd_lane = threadIdx.x < 16 * HP2;
  const int qr = threadIdx.x / HP2, qh2 = threadIdx.x % HP2;
  __nv_bfloat162 pqu = z2, pqv = z2, pdo = z2;
  auto fetch_q = [&](int blk) {
    pqu = pqv = pdo = z2;
    const int t = blk * 16 + qr;
    if (qd_lane && blk < nblk && t < T && 2 * qh2 < H) {
      const __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(p.qkv + ((size_t)b * T + t) * p.ld_qkv + (size_t)n * H + 2 * qh2);
      const float qa = __low2float(q), qb = __high2float(q);
      const float* ub = p.ubias + n * H + 2 * qh2;
      const float* vb = p.vbias + n * H + 2 * qh2;
      pqu = __floats2bfloat162_rn((qa + ub[0]) * p.scale, (qb + ub[1]) * p.scale);
      pqv = __floats2bfloat162_rn((qa + vb[0]) * p.scale, (qb + vb[1]) * p.scale);
      pdo = *reinterpret_cast<const __nv_bfloat162*>(dctx + ((size_t)b * T + t) * NH + (size_t)n * H + 2 * qh2);
    }
  };

  float aV[KS * 2][4], aK[KS * 2][4], aP[2][KS * 2][4];
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh)
#pragma unroll
    for (int e = 0; e < 4; ++e) aV[nh][e] = aK[nh][e] = aP[0][nh][e] = aP[1][nh][e] = 0.f;

  issue(0);
  if (nblk > 1) issue(1);
  fetch_q(0);
  const int ak = (lane & 7) + (lane >> 4) * 8;  // ldmatrix.trans A: source row (k)
  const int am = ((lane >> 3) & 1) * 8;         // and column offset (m)
  for (int blk = 0; blk < nblk; ++blk) {
    if (blk + 1 < nblk) mma::cp_async_wait<1>();
    else mma::cp_async_wait<0>();
    __nv_bfloat16* sQu = sQ + (size_t)(blk & 1) * 3 * 16 * HS;
    __nv_bfloat16* sQv = sQu + 16 * HS;
    __nv_bfloat16* sdO = sQv + 16 * HS;
    if (qd_lane) {
      *reinterpret_cast<__nv_bfloat162*>(sQu + qr * HS + 2 * qh2) = pqu;
      *reinterpret_cast<__nv_bfloat162*>(sQv + qr * HS + 2 * qh2) = pqv;
      *reinterpret_cast<__nv_bfloat162*>(sdO + qr * HS + 2 * qh2) = pdo;
    }
    __syncthreads();  // step blk's rows are visible; every warp is done with step blk - 1
    if (blk + 2 < nblk) issue(blk + 2);
    fetch_q(blk + 1);
    const __nv_bfloat16* sPb = sRaw + (size_t)(blk % 3) * 2 * 16 * PS;
    const __nv_bfloat16* sDb = sPb + 16 * PS;
    // skewed copy: sSk[r][c'] = dS[t][c' - 15 + r - rem]
    {  // warp w skews row w
      const int shift = 15 - warp + rem;  // c' = j + shift
      const __nv_bfloat16* src = sDb + warp * PS;
      __nv_bfloat16* dst = sSk + warp * QS;
      for (int c = lane; c < QS; c += 32) {
        const int j = c - shift;
        dst[c] = (j >= 0 && j < T) ? src[j] : __float2bfloat16(0.f);
      }
    }
    __syncthreads();
    if (warp < nblk) {
      uint32_t afp[4], afd[4];
      mma::ldsm_x4_trans(afp, sPb + ak * PS + warp * 16 + am);
      mma::ldsm_x4_trans(afd, sDb + ak * PS + warp * 16 + am);
#pragma unroll
      for (int nh = 0; nh < KS * 2; ++nh) {
        uint32_t bo[2], bq[2];
        mma::ldsm_x2_trans(bo, sdO + (lane & 15) * HS + nh * 8);
        mma::ldsm_x2_trans(bq, sQu + (lane & 15) * HS + nh * 8);
        mma::mma_16816(aV[nh], afp, bo);
        mma::mma_16816(aK[nh], afd, bq);
      }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int ct = (2 * warp + mi) - q0 + blk;
      if (ct >= 0 && ct < nct) {
        uint32_t afs[4];
        mma::ldsm_x4_trans(afs, sSk + ak * QS + ct * 16 + am);
#pragma unroll
        for (int nh = 0; nh < KS * 2; ++nh) {
          uint32_t bq[2];
          mma::ldsm_x2_trans(bq, sQv + (lane & 15) * HS + nh * 8);
          mma::mma_16816(aP[mi][nh], afs, bq);
        }
      }
    }
  }
#pragma unroll
  for (int nh = 0; nh < KS * 2; ++nh) {
    const int h = nh * 8 + 2 * tq;
    if (h >= H) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = warp * 16 + g + half * 8;
      if (j < T) {
        __nv_bfloat16* row = dqkv + ((size_t)b * T + j) * p.ld_qkv + (size_t)n * H + h;
        *reinterpret_cast<__nv_bfloat162*>(row + NH) = __floats2bfloat162_rn(aK[nh][2 * half], aK[nh][2 * half + 1]);
        *reinterpret_cast<__nv_bfloat162*>(row + 2 * NH) = __floats2bfloat162_rn(aV[nh][2 * half], aV[nh][2 * half + 1]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = (2 * warp + mi) * 16 + g + half * 8;
        if (r < R)
          *reinterpret_cast<__nv_bfloat162*>(dpos + ((size_t)b * R + r) * p.ld_dpos + (size_t)n * H + h) =
              __floats2bfloat162_rn(aP[mi][nh][2 * half], aP[mi][nh][2 * half + 1]);
      }
    }
  }
}

template <int KS>
static int launch_att_bwd_mma(const AttParams& p, const __nv_bfloat16* probs, const __nv_bfloat16* dctx, const __nv_bfloat16* octx,
                              __nv_bfloat16* ds_ws, __nv_bfloat16* dqkv, __nv_bfloat16* dpos, float* du, float* dvb, cudaStream_t stream) {
  const size_t smem1 = att_mma_bwd_rows_smem<KS>(p.T), smem2 = att_mma_bwd_cols_smem<KS>(p.T);
  if (smem1 > 227 * 1024 || smem2 > 227 * 1024 || p.T > 256 || p.T < 16 || (p.H & 3) || (p.ld_qkv & 3) || (p.ld_pos & 3) ||
      (reinterpret_cast<uintptr_t>(p.pos) & 7) || !octx)
    return -1;
  if (cudaFuncSetAttribute(rel_attention_bwd_rows_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(rel_attention_bwd_cols_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return TFASR_EXECUTION_FAILED;
  dim3 grid(p.N, p.B);
  launch_pdl(rel_attention_bwd_rows_mma_kernel<KS>, dim3(grid), dim3(512), smem1, stream, p, probs, dctx, octx, ds_ws, dqkv, du, dvb);
  count_launch();
  const size_t smem3 = att_mma_bwd_cols2_smem<KS>(p.T);
  if ((p.T & 1) == 0 && smem3 <= 227 * 1024) {
    if (cudaFuncSetAttribute(rel_attention_bwd_cols_mma2_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    launch_pdl(rel_attention_bwd_cols_mma2_kernel<KS>, dim3(grid), dim3(512), smem3, stream, p, probs, ds_ws, dctx, dqkv, dpos);
  } else {
    rel_attention_bwd_cols_mma_kernel<KS><<<grid, 512, smem2, stream>>>(p, probs, ds_ws, dctx, dqkv, dpos);
  }
  count_launch();
  return check_launch();
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
  if (p.ld_pos < p.N * p.H || p.ld_dpos < p.N * p.H || (p.ld_pos & 1) || (p.ld_dpos & 1)) return TFASR_INVALID_VALUE;
  return TFASR_OK;
}

}  // namespace tfasr

using namespace tfasr;

static AttParams make_att(const void* qkv, int ld_qkv, const void* pos, const float* u, const float* v, const int32_t* length, int B, int T, int N,
                          int H, int chunk, int history, int ld_pos, int ld_dpos) {
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
  p.ld_pos = ld_pos;
  p.ld_dpos = ld_dpos;
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

static int att_fwd_impl(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                        const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx, void* probs,
                        tfasr_stream_t stream_) {
  AttParams p = make_att(qkv, ld_qkv, pos, ubias, vbias, length, B, T, N, H, chunk, history, ld_pos, N * H);
  int st = att_check(p);
  if (st || !ctx || !probs) return st ? st : TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  {
    // tensor-core path when the operands fit in SMEM (T <= ~256 at H <= 48); otherwise the FP32 kernel below
    int r = -1;
    const int ks = (H + 15) / 16;
    if (ks <= 1) r = launch_att_fwd_mma<1>(p, static_cast<__nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(probs), stream);
    else if (ks == 2) r = launch_att_fwd_mma<2>(p, static_cast<__nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(probs), stream);
    else if (ks == 3) r = launch_att_fwd_mma<3>(p, static_cast<__nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(probs), stream);
    else if (ks == 4) r = launch_att_fwd_mma<4>(p, static_cast<__nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(probs), stream);
    if (r >= 0) return r;
  }
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

extern "C" int tfasr_rel_attention_fwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias,
                                       const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx,
                                       void* probs, tfasr_stream_t stream) {
  return att_fwd_impl(qkv, ld_qkv, pos, N * H, ubias, vbias, length, B, T, N, H, chunk, history, ctx, probs, stream);
}
extern "C" int tfasr_rel_attention_fwd_ld(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                                          const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history,
                                          void* ctx, void* probs, tfasr_stream_t stream) {
  return att_fwd_impl(qkv, ld_qkv, pos, ld_pos, ubias, vbias, length, B, T, N, H, chunk, history, ctx, probs, stream);
}

static int att_bwd_impl(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                        const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, const void* probs,
                        const void* dctx, const void* ctx, void* ds_ws, void* dqkv, void* dpos, int32_t ld_dpos, float* du, float* dvb,
                        tfasr_stream_t stream_) {
  AttParams p = make_att(qkv, ld_qkv, pos, ubias, vbias, length, B, T, N, H, chunk, history, ld_pos, ld_dpos);
  int st = att_check(p);
  if (st || !probs || !dctx || !ds_ws || !dqkv || !dpos || !du || !dvb) return st ? st : TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  {
    int r = -1;
    const int ks = (H + 15) / 16;
#define BWD_MMA(KS_)                                                                                                                  \
  r = launch_att_bwd_mma<KS_>(p, static_cast<const __nv_bfloat16*>(probs), static_cast<const __nv_bfloat16*>(dctx),                   \
                              static_cast<const __nv_bfloat16*>(ctx), static_cast<__nv_bfloat16*>(ds_ws), static_cast<__nv_bfloat16*>(dqkv), static_cast<__nv_bfloat16*>(dpos), \
                              du, dvb, stream)
    if (ks <= 1) BWD_MMA(1);
    else if (ks == 2) BWD_MMA(2);
    else if (ks == 3) BWD_MMA(3);
    else if (ks == 4) BWD_MMA(4);
#undef BWD_MMA
    if (r >= 0) return r;
  }
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

extern "C" int tfasr_rel_attention_bwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias,
                                       const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history,
                                       const void* probs, const void* dctx, const void* ctx, void* ds_ws, void* dqkv, void* dpos, float* du,
                                       float* dvb, tfasr_stream_t stream) {
  return att_bwd_impl(qkv, ld_qkv, pos, N * H, ubias, vbias, length, B, T, N, H, chunk, history, probs, dctx, ctx, ds_ws, dqkv, dpos, N * H, du, dvb,
                      stream);
}
extern "C" int tfasr_rel_attention_bwd_ld(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                                          const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history,
                                          const void* probs, const void* dctx, const void* ctx, void* ds_ws, void* dqkv, void* dpos, int32_t ld_dpos,
                                          float* du, float* dvb, tfasr_stream_t stream) {
  return att_bwd_impl(qkv, ld_qkv, pos, ld_pos, ubias, vbias, length, B, T, N, H, chunk, history, probs, dctx, ctx, ds_ws, dqkv, dpos, ld_dpos, du,
                      dvb, stream);
}
