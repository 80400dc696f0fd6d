// RNN-T gradient kernel K3', second generation (opt-in: ops.set_rnnt_grad_v2 / TFASR_RNNT_GRAD_V2=1).  Same contract as
// tfasr_rnnt_loss_bwd_bf16 (rnnt_loss.cu): grads = g - softmax * sum(g) (losses/impl/rnnt.py:233-275) scaled by the upstream gradient
// (:321-324), written as bf16 with row pitch ld_grads, optionally IN PLACE over the fp32 logits, plus the vocab-bias gradient.
//
// Why: the first-generation kernel holds a whole row (8 float4 per lane at V = 1000) AND its per-column bias partial sums (8 more float4)
// in registers: 126 registers -> 2 CTAs per SM = 25 % occupancy, 4.2 TB/s on 3.09 GB (profiles/rnnt_ncu_r01.txt).  Here
//   * a row is processed in two halves (4 float4 per lane live at a time).  In-place safety: the bf16 gradient of columns [0, 4 h4) lands
//     in bytes [0, 8 h4) of the row = fp32 columns [0, 2 h4) — inside the first half, which every lane has in registers (one __syncwarp)
//     before any lane stores; the second half's gradient lands in fp32 columns [2 h4, 2 n4) <= first half again, already consumed, and
//     the second half's loads (fp32 float4 index >= h4) are never touched by first-half stores (float4 index < h4 / 2).
//   * the bias partial sums live in warp-private SMEM rows (plain read-modify-write, no atomics, no bank conflicts within a warp)
//     and are reduced once per CTA.
// HBM-bound byte work (12 B per lattice node and vocabulary entry incl. the second read); no tensor cores.  Plain SIMT: builds twice.
#include "simt.cuh"
#include "lattice_lens.cuh"

#ifdef TFASR_EMU_BUILD
#include "../../include/tfasr_b200.h"
typedef void* cudaStream_t;
#endif

namespace tfasr {

constexpr int kG2Warps = 8;

struct RnntPlanes {  // the workspace planes of rnnt_loss.cu (rnnt_ws_carve): [B, T, U1] f32 each
  const float *lse, *blank, *truth, *alpha, *beta;
};

// NV4H = float4 per lane per half row (V <= 256 * NV4H)
template <int NV4H>
__global__ void __launch_bounds__(kG2Warps * 32, 3) rnnt_grad_bf16_v2_kernel(const float* logits, const int32_t* __restrict__ labels,
                                                                             const int32_t* __restrict__ logit_len,
                                                                             const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                             RnntPlanes ws, const float* __restrict__ costs,
                                                                             const float* __restrict__ grad_scale, __nv_bfloat16* grads, int64_t ld_g,
                                                                             float* __restrict__ dbias) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_cs = reinterpret_cast<float*>(smem_raw);  // [kG2Warps][V] when dbias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nrows = (int64_t)B * T * U1;
  const int n4 = V >> 2, h4 = (n4 + 1) >> 1;
  if (dbias) {
    for (int j = threadIdx.x; j < kG2Warps * V; j += blockDim.x) s_cs[j] = 0.f;
    __syncthreads();
  }
  float4* cs4 = reinterpret_cast<float4*>(s_cs + (size_t)warp * V);
  for (int64_t row = (int64_t)blockIdx.x * kG2Warps + warp; row < nrows; row += (int64_t)gridDim.x * kG2Warps) {
    const uint32_t r32 = (uint32_t)row;  // B T U1 < 2^31 (checked by the launcher): 32-bit divisions
    const int u = (int)(r32 % (uint32_t)U1);
    const int t = (int)((r32 / (uint32_t)U1) % (uint32_t)T);
    const int b = (int)(r32 / ((uint32_t)U1 * (uint32_t)T));
    int Tb, Ub;
  lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
    uint2* g2 = reinterpret_cast<uint2*>(grads + row * ld_g);
    if (!((t < Tb) && (u <= Ub))) {
      for (int j = lane; j < n4; j += 32) g2[j] = make_uint2(0u, 0u);
      continue;
    }
    const float logz = -costs[b];
    const float a = ws.alpha[row];
    const float lse = ws.lse[row];
    float gb = 0.f, gt = 0.f;
    int label = -1;
    if (t < Tb - 1) gb = -expf(a + ws.beta[row + U1] + ws.blank[row] - logz);
    else if (u == Ub) gb = -1.0f;
    if (u < Ub) {
      gt = -expf(a + ws.beta[row + 1] + ws.truth[row] - logz);
      label = labels[(int64_t)b * (U1 - 1) + u];
    }
    const float sc = grad_scale ? grad_scale[b] : 1.0f;
    gb *= sc;
    gt *= sc;
    const float ng = -(gb + gt), nl = -lse * kLog2e;
    const float4* x4 = reinterpret_cast<const float4*>(logits + row * V);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j0 = half ? h4 : 0, j1 = half ? n4 : h4;
      float4 v[NV4H];
#pragma unroll
      for (int i = 0; i < NV4H; ++i) {
        const int j = j0 + lane + 32 * i;
        if (j < j1) v[i] = simt_ld_stream_f4(x4 + j);
      }
      if (half == 0) __syncwarp();  // in-place mode: the whole first half is in registers before any lane overwrites part of it
#pragma unroll
      for (int i = 0; i < NV4H; ++i) {
        const int j = j0 + lane + 32 * i;
        if (j < j1) {
          float4 o;
          o.x = ng * simt_ex2(fmaf(v[i].x, kLog2e, nl));  // -softmax * sum(g): FFMA + ex2.approx.ftz + FMUL per logit
          o.y = ng * simt_ex2(fmaf(v[i].y, kLog2e, nl));
          o.z = ng * simt_ex2(fmaf(v[i].z, kLog2e, nl));
          o.w = ng * simt_ex2(fmaf(v[i].w, kLog2e, nl));
          const int c0 = 4 * j;
          if (c0 == 0) o.x += gb;
          if (label >= c0 && label < c0 + 4) {
            const int r = label - c0;
            if (r == 0) o.x += gt;
            else if (r == 1) o.y += gt;
            else if (r == 2) o.z += gt;
            else o.w += gt;
          }
          if (dbias) {
            float4 c = cs4[j];
            c.x += o.x;
            c.y += o.y;
            c.z += o.z;
            c.w += o.w;
            cs4[j] = c;
          }
          const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          g2[j] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
      }
    }
  }
  if (dbias) {
    __syncthreads();
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kG2Warps; ++w) s += s_cs[(size_t)w * V + j];
      atomicAdd(dbias + j, s);
    }
  }
}

// bf16 logits in, bf16 gradient out (ops.bf16_logits): element j of a row is read and written by the same lane, so grads may alias logits
// (ld_g = V) with no ordering concern.  NV2 = uint2 (4 bf16) per lane (V <= 128 * NV2): 4 B per lattice node and vocabulary entry.
template <int NV2>
__global__ void __launch_bounds__(kG2Warps * 32, 3) rnnt_grad_bf16_io_kernel(const __nv_bfloat16* logits, int64_t ld_x, const int32_t* __restrict__ labels,
                                                                             const int32_t* __restrict__ logit_len,
                                                                             const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                             RnntPlanes ws, const float* __restrict__ costs,
                                                                             const float* __restrict__ grad_scale, __nv_bfloat16* grads, int64_t ld_g,
                                                                             float* __restrict__ dbias) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_cs = reinterpret_cast<float*>(smem_raw);  // [kG2Warps][V] when dbias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nrows = (int64_t)B * T * U1;
  const int n4 = V >> 2;
  if (dbias) {
    for (int j = threadIdx.x; j < kG2Warps * V; j += blockDim.x) s_cs[j] = 0.f;
    __syncthreads();
  }
  float4* cs4 = reinterpret_cast<float4*>(s_cs + (size_t)warp * V);
  for (int64_t row = (int64_t)blockIdx.x * kG2Warps + warp; row < nrows; row += (int64_t)gridDim.x * kG2Warps) {
    const uint32_t r32 = (uint32_t)row;  // B T U1 < 2^31 (checked by the launcher): 32-bit divisions
    const int u = (int)(r32 % (uint32_t)U1);
    const int t = (int)((r32 / (uint32_t)U1) % (uint32_t)T);
    const int b = (int)(r32 / ((uint32_t)U1 * (uint32_t)T));
    int Tb, Ub;
    lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
    uint2* g2 = reinterpret_cast<uint2*>(grads + row * ld_g);
    if (!((t < Tb) && (u <= Ub))) {
      for (int j = lane; j < n4; j += 32) g2[j] = make_uint2(0u, 0u);
      continue;
    }
    const float logz = -costs[b];
    const float a = ws.alpha[row];
    const float lse = ws.lse[row];
    float gb = 0.f, gt = 0.f;
    int label = -1;
    if (t < Tb - 1) gb = -expf(a + ws.beta[row + U1] + ws.blank[row] - logz);
    else if (u == Ub) gb = -1.0f;
    if (u < Ub) {
      gt = -expf(a + ws.beta[row + 1] + ws.truth[row] - logz);
      label = labels[(int64_t)b * (U1 - 1) + u];
    }
    const float sc = grad_scale ? grad_scale[b] : 1.0f;
    gb *= sc;
    gt *= sc;
    const float gsum = gb + gt;
    const uint2* x2 = reinterpret_cast<const uint2*>(logits + row * ld_x);
    uint2 v[NV2];
#pragma unroll
    for (int i = 0; i < NV2; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) v[i] = simt_ld_stream_u2(x2 + j);
    }
#pragma unroll
    for (int i = 0; i < NV2; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        float4 o;
        o.x = -__expf(__uint_as_float(v[i].x << 16) - lse) * gsum;
        o.y = -__expf(__uint_as_float(v[i].x & 0xffff0000u) - lse) * gsum;
        o.z = -__expf(__uint_as_float(v[i].y << 16) - lse) * gsum;
        o.w = -__expf(__uint_as_float(v[i].y & 0xffff0000u) - lse) * gsum;
        const int c0 = 4 * j;
        if (c0 == 0) o.x += gb;
        if (label >= c0 && label < c0 + 4) {
          const int r = label - c0;
          if (r == 0) o.x += gt;
          else if (r == 1) o.y += gt;
          else if (r == 2) o.z += gt;
          else o.w += gt;
        }
        if (dbias) {
          float4 c = cs4[j];
          c.x += o.x;
          c.y += o.y;
          c.z += o.z;
          c.w += o.w;
          cs4[j] = c;
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        g2[j] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
    }
  }
  if (dbias) {
    __syncthreads();
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kG2Warps; ++w) s += s_cs[(size_t)w * V + j];
      atomicAdd(dbias + j, s);
    }
  }
}

// The same, 16 bytes (8 logits) per access and a short instruction stream (V % 8 == 0, 16-byte aligned rows): the first version spent
// 22 thread-instructions per logit (profiles/head_bf16_ncu_r02.txt: 355 M warp instructions, issue slots 57 % busy, 3.8 TB/s) on the
// range fix-up of __expf, the per-group label compares and the 8-byte accesses.  Here a softmax term is FFMA + ex2.approx.ftz + FMUL, and
// the two columns that get an extra term (blank: column 0, truth: column `label`) are patched after the vector pass by the lane that
// owns them — it re-reads its own logit before its own stores, so the in-place update needs no cross-lane ordering.
template <int NV8>
__global__ void __launch_bounds__(kG2Warps * 32, 3) rnnt_grad_bf16_io8_kernel(const __nv_bfloat16* logits, int64_t ld_x, const int32_t* __restrict__ labels,
                                                                              const int32_t* __restrict__ logit_len,
                                                                              const int32_t* __restrict__ label_len, int B, int T, int U1, int V,
                                                                              RnntPlanes ws, const float* __restrict__ costs,
                                                                              const float* __restrict__ grad_scale, __nv_bfloat16* grads, int64_t ld_g,
                                                                              float* __restrict__ dbias) {
  TFASR_DYN_SMEM(smem_raw);
  float* s_cs = reinterpret_cast<float*>(smem_raw);  // [kG2Warps][V] when dbias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nrows = (int64_t)B * T * U1;
  const int n8 = V >> 3;
  if (dbias) {
    for (int j = threadIdx.x; j < kG2Warps * V; j += blockDim.x) s_cs[j] = 0.f;
    __syncthreads();
  }
  float* cs = s_cs + (size_t)warp * V;
  float4* cs4 = reinterpret_cast<float4*>(cs);
  // A warp takes 32 consecutive rows at a time: lane l derives the per-row scalars of row base + l (lattice position, the two transition
  // gradients — two expf —, the label) with coalesced reads of the [B,T,U1] planes, and the row loop fetches them by shuffle.  The first
  // version derived them per row on every lane: ~150 warp instructions per row against ~200 for the row's 1000 logits
  // (profiles/head_bf16_ncu_r02.txt: 266 M warp instructions, issue slots 55 % busy, 4.7 TB/s).
  const int64_t nchunks = (nrows + 31) >> 5;
  for (int64_t chunk = (int64_t)blockIdx.x * kG2Warps + warp; chunk < nchunks; chunk += (int64_t)gridDim.x * kG2Warps) {
    const int64_t row0 = chunk << 5;
    float nl_l = 0.f, ng_l = 0.f, gb_l = 0.f, gt_l = 0.f;
    int label_l = -2;  // -2: row outside its utterance's lattice (zero gradient); -1: inside, no truth transition
    if (row0 + lane < nrows) {
      const int64_t rl = row0 + lane;
      const uint32_t r32 = (uint32_t)rl;  // B T U1 < 2^31 (checked by the launcher): 32-bit divisions
      const int u = (int)(r32 % (uint32_t)U1);
      const int t = (int)((r32 / (uint32_t)U1) % (uint32_t)T);
      const int b = (int)(r32 / ((uint32_t)U1 * (uint32_t)T));
      int Tb, Ub;
      lattice_lens(logit_len, label_len, b, T, U1 - 1, Tb, Ub);
      if ((t < Tb) && (u <= Ub)) {
        const float logz = -costs[b];
        const float al = ws.alpha[rl];
        float gb = 0.f, gt = 0.f;
        label_l = -1;
        if (t < Tb - 1) gb = -expf(al + ws.beta[rl + U1] + ws.blank[rl] - logz);
        else if (u == Ub) gb = -1.0f;
        if (u < Ub) {
          gt = -expf(al + ws.beta[rl + 1] + ws.truth[rl] - logz);
          label_l = labels[(int64_t)b * (U1 - 1) + u];
        }
        const float sc = grad_scale ? grad_scale[b] : 1.0f;
        gb_l = gb * sc;
        gt_l = gt * sc;
        ng_l = -(gb_l + gt_l);
        nl_l = -ws.lse[rl] * kLog2e;
      }
    }
    const int nr = (int)((nrows - row0 < 32) ? (nrows - row0) : 32);
    // the NEXT row's logits are requested before the current row is processed: one row per warp in flight (2 KB) leaves the kernel at
    // the latency-bandwidth product of 24 warps per SM (4.8 TB/s measured); two rows per warp double the bytes in flight
    uint4 vn[NV8];
    float x0n = 0.f, xln = 0.f;
    auto fetch = [&](int r) {
      const int label = __shfl_sync(0xffffffffu, label_l, r);
      if (label == -2) return;
      const __nv_bfloat16* xr = logits + (row0 + r) * ld_x;
      const uint4* x8 = reinterpret_cast<const uint4*>(xr);
      if (lane == 0) x0n = __bfloat162float(xr[0]);
      if (label >= 0 && ((label >> 3) & 31) == lane) xln = __bfloat162float(xr[label]);
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        const int j = lane + 32 * i;
        if (j < n8) vn[i] = simt_ld_stream_u4(x8 + j);
      }
    };
    fetch(0);
    for (int r = 0; r < nr; ++r) {
      const int label = __shfl_sync(0xffffffffu, label_l, r);
      const float nl = __shfl_sync(0xffffffffu, nl_l, r), ng = __shfl_sync(0xffffffffu, ng_l, r);
      const float gb = __shfl_sync(0xffffffffu, gb_l, r), gt = __shfl_sync(0xffffffffu, gt_l, r);
      const int64_t row = row0 + r;
      __nv_bfloat16* gr = grads + row * ld_g;
      uint4* g8 = reinterpret_cast<uint4*>(gr);
      uint4 v[NV8];
#pragma unroll
      for (int i = 0; i < NV8; ++i) v[i] = vn[i];
      const float x0 = x0n, xl = xln;
      if (r + 1 < nr) fetch(r + 1);  // in place: row r + 1 is a different row, nothing of it is written before it is read
      if (label == -2) {
        for (int j = lane; j < n8; j += 32) g8[j] = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      const bool own_label = label >= 0 && ((label >> 3) & 31) == lane;
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        const int j = lane + 32 * i;
        if (j < n8) {
          float f[8];
          bf8_unpack(v[i], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = ng * simt_ex2(fmaf(f[k], kLog2e, nl));
          if (dbias) {
            float4 c0 = cs4[2 * j], c1 = cs4[2 * j + 1];
            c0.x += f[0];
            c0.y += f[1];
            c0.z += f[2];
            c0.w += f[3];
            c1.x += f[4];
            c1.y += f[5];
            c1.z += f[6];
            c1.w += f[7];
            cs4[2 * j] = c0;
            cs4[2 * j + 1] = c1;
          }
          g8[j] = bf8_pack(f);
        }
      }
      // the two columns with an extra term, by their owners (same thread as the vector store above: program order)
      if (lane == 0) {
        gr[0] = __float2bfloat16_rn(ng * simt_ex2(fmaf(x0, kLog2e, nl)) + gb);
        if (dbias) cs[0] += gb;
      }
      if (own_label) {
        gr[label] = __float2bfloat16_rn(ng * simt_ex2(fmaf(xl, kLog2e, nl)) + gt);
        if (dbias) cs[label] += gt;
      }
    }
  }
  if (dbias) {
    __syncthreads();
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kG2Warps; ++w) s += s_cs[(size_t)w * V + j];
      atomicAdd(dbias + j, s);
    }
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_rnnt_loss_bwd_bf16_v2(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                           int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                                           int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  if (blank != 0 || B < 0 || T < 1 || U1 < 1 || V < 2 || U1 > 1024 || (int64_t)B * T * U1 > 0x7fffffff) return TFASR_INVALID_VALUE;
  if (!logits || !logit_len || !label_len || !workspace || (U1 > 1 && !labels) || !costs || !grads || ld_grads < V || (ld_grads & 3))
    return TFASR_INVALID_VALUE;
  if ((V & 3) || V > 2048 || (reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(grads) & 7)) return TFASR_INVALID_VALUE;
  // workspace layout of rnnt_loss.cu (rnnt_ws_carve): five [B,T,U1] f32 planes, each rounded up to 256 bytes: lse, blank, truth, alpha, beta
  const size_t plane = (sizeof(float) * (size_t)B * T * U1 + 255) / 256 * 256;
  if (workspace_bytes < 5 * plane) return TFASR_WORKSPACE_TOO_SMALL;
  if (B == 0) return TFASR_OK;
  const char* wp = static_cast<const char*>(workspace);
  RnntPlanes ws;
  ws.lse = reinterpret_cast<const float*>(wp);
  ws.blank = reinterpret_cast<const float*>(wp + plane);
  ws.truth = reinterpret_cast<const float*>(wp + 2 * plane);
  ws.alpha = reinterpret_cast<const float*>(wp + 3 * plane);
  ws.beta = reinterpret_cast<const float*>(wp + 4 * plane);
  const int64_t nrows = (int64_t)B * T * U1;
  int64_t ctas = ceil_div(nrows, kG2Warps);
  if (ctas > (int64_t)kNumSMs * 12) ctas = (int64_t)kNumSMs * 12;
  const size_t smem = dbias ? sizeof(float) * (size_t)kG2Warps * V : 0;
  __nv_bfloat16* g = static_cast<__nv_bfloat16*>(grads);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
#define G2_LAUNCH(N) \
  emu::launch(dim3((unsigned)ctas), dim3(kG2Warps * 32), smem, [&] { rnnt_grad_bf16_v2_kernel<N>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias); })
#else
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
#define G2_LAUNCH(N)                                                                                                                         \
  do {                                                                                                                                       \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(rnnt_grad_bf16_v2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
      return TFASR_EXECUTION_FAILED;                                                                                                         \
    rnnt_grad_bf16_v2_kernel<N><<<(unsigned)ctas, kG2Warps * 32, smem, stream>>>(logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, \
                                                                                 g, ld_grads, dbias);                                       \
    count_launch();                                                                                                                          \
  } while (0)
#endif
  if (V <= 256) G2_LAUNCH(1);
  else if (V <= 512) G2_LAUNCH(2);
  else if (V <= 1024) G2_LAUNCH(4);
  else G2_LAUNCH(8);
#undef G2_LAUNCH
#ifdef TFASR_EMU_BUILD
  return TFASR_OK;
#else
  return check_launch();
#endif
}

/* bf16 logits (row pitch ld_logits) -> bf16 gradient (may be in place: grads == logits, ld_grads = ld_logits); else as tfasr_rnnt_loss_bwd_bf16_v2 */
extern "C" int tfasr_rnnt_loss_bwd_bf16_io(const void* logits, int64_t ld_logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                           int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                                           int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream_) {
  if (blank != 0 || B < 0 || T < 1 || U1 < 1 || V < 2 || U1 > 1024 || (int64_t)B * T * U1 > 0x7fffffff) return TFASR_INVALID_VALUE;
  if (!logits || !logit_len || !label_len || !workspace || (U1 > 1 && !labels) || !costs || !grads || ld_grads < V || (ld_grads & 3))
    return TFASR_INVALID_VALUE;
  if ((V & 3) || V > 2048 || ld_logits < V || (ld_logits & 3) || (reinterpret_cast<uintptr_t>(logits) & 7) || (reinterpret_cast<uintptr_t>(grads) & 7))
    return TFASR_INVALID_VALUE;
  const size_t plane = (sizeof(float) * (size_t)B * T * U1 + 255) / 256 * 256;
  if (workspace_bytes < 5 * plane) return TFASR_WORKSPACE_TOO_SMALL;
  if (B == 0) return TFASR_OK;
  const char* wp = static_cast<const char*>(workspace);
  RnntPlanes ws;
  ws.lse = reinterpret_cast<const float*>(wp);
  ws.blank = reinterpret_cast<const float*>(wp + plane);
  ws.truth = reinterpret_cast<const float*>(wp + 2 * plane);
  ws.alpha = reinterpret_cast<const float*>(wp + 3 * plane);
  ws.beta = reinterpret_cast<const float*>(wp + 4 * plane);
  const int64_t nrows = (int64_t)B * T * U1;
  int64_t ctas = ceil_div(nrows, kG2Warps);
  if (ctas > (int64_t)kNumSMs * 12) ctas = (int64_t)kNumSMs * 12;
  const size_t smem = dbias ? sizeof(float) * (size_t)kG2Warps * V : 0;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(logits);
  // 16-byte accesses when every row allows them
  const bool wide = !(V & 7) && !(ld_logits & 7) && !(ld_grads & 7) && !(reinterpret_cast<uintptr_t>(logits) & 15) && !(reinterpret_cast<uintptr_t>(grads) & 15);
  __nv_bfloat16* g = static_cast<__nv_bfloat16*>(grads);
#ifdef TFASR_EMU_BUILD
  (void)stream_;
#define G2_LAUNCH(KERNEL, N) \
  emu::launch(dim3((unsigned)ctas), dim3(kG2Warps * 32), smem, [&] { KERNEL<N>(x, ld_logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, ld_grads, dbias); })
#else
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
#define G2_LAUNCH(KERNEL, N)                                                                                                         \
  do {                                                                                                                               \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(KERNEL<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
      return TFASR_EXECUTION_FAILED;                                                                                                 \
    KERNEL<N><<<(unsigned)ctas, kG2Warps * 32, smem, stream>>>(x, ld_logits, labels, logit_len, label_len, B, T, U1, V, ws, costs, grad_scale, g, \
                                                               ld_grads, dbias);                                                    \
    count_launch();                                                                                                                  \
  } while (0)
#endif
  if (wide) {
    ctas = ceil_div(ceil_div(nrows, 32), kG2Warps);  // one 32-row chunk per warp and pass: CTAs retire and refill, which balances the SMs
    if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
    if (V <= 256) G2_LAUNCH(rnnt_grad_bf16_io8_kernel, 1);
    else if (V <= 512) G2_LAUNCH(rnnt_grad_bf16_io8_kernel, 2);
    else if (V <= 1024) G2_LAUNCH(rnnt_grad_bf16_io8_kernel, 4);
    else G2_LAUNCH(rnnt_grad_bf16_io8_kernel, 8);
  } else if (V <= 256) G2_LAUNCH(rnnt_grad_bf16_io_kernel, 2);
  else if (V <= 512) G2_LAUNCH(rnnt_grad_bf16_io_kernel, 4);
  else if (V <= 1024) G2_LAUNCH(rnnt_grad_bf16_io_kernel, 8);
  else G2_LAUNCH(rnnt_grad_bf16_io_kernel, 16);
#undef G2_LAUNCH
#ifdef TFASR_EMU_BUILD
  return TFASR_OK;
#else
  return check_launch();
#endif
}
