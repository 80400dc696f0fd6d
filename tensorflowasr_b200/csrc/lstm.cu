// Recurrent part of the prediction network's LSTM for sm_100a (forward and backward), one thread-block cluster per 32 utterances.
//
// Replaces the time loop of keras.layers.LSTM in TransducerPrediction (tensorflow_asr/models/transducer/base_transducer.py:74-88,
// 113-127; gate order i, f, c, o; one bias) which TF runs as a cuDNN RNN / while_loop: ~4 small kernels per label step.  The input
// projection x . W_ih + b for all steps is one tcgen05 GEMM before this kernel; what is sequential is
//     pre_t = gx_t + h_{t-1} . W_hh;   i, f, o = sigmoid,  g = tanh;   c_t = f c_{t-1} + i g;   h_t = o tanh(c_t)
// Design: a cluster of 8 CTAs keeps W_hh resident in SMEM as bf16 for the whole sequence, each CTA owning R/8 hidden units (all four
// gates of a unit, so the cell update is register-local: a thread owns the same (utterance, unit) pairs at every step and carries
// c_t / dc_t in registers).  Per step each CTA publishes its slice (h_t, resp. the gate gradients) in its own SMEM, one cluster
// barrier later every CTA pulls the 8 slices over DSMEM into a full operand tile and runs mma.sync (m16n8k16, bf16 -> f32) against
// its weight slice.  One cluster barrier + one CTA barrier per label step; no global-memory round trip on the critical path.
#include "common.cuh"
#include "attention_mma.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace tfasr {

constexpr int kLstmCluster = 8;
constexpr int kLstmRows = 32;  // utterances per cluster (two 16-row MMA tiles)

// one MUFU op per activation (tanh.approx.f32, relative error ~2^-11): h is rounded to bf16 (2^-9) right after
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

struct LstmParams {
  int B, U, R;
};

// ---------------------------------------------------------------------------------------------- forward
// gx [B, U, 4R] f32 (input projection + bias), whh [R, 4R] f32 (keras recurrent_kernel), h0 / c0 [B, R] f32 or null,
// length [B] or null (outputs at t >= length are written as zeros; the state keeps evolving, as with zero_output_for_mask)
// -> y [B, U, R] bf16, saved_g [B, U, R/2] x 16 B (i, f, g, o of a unit pair as bf16), saved_c [B, U, R] f32 (cell state after the step),
//    hT / cT [B, R] f32 (state after the last step)
__global__ void __cluster_dims__(kLstmCluster, 1, 1) __launch_bounds__(512, 1)
    lstm_rec_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const float* __restrict__ h0, const float* __restrict__ c0,
                        const int32_t* __restrict__ length, __nv_bfloat16* __restrict__ y, uint4* __restrict__ saved_g, float* __restrict__ saved_c,
                        float* __restrict__ hT,
                        float* __restrict__ cT, const LstmParams p) {
  extern __shared__ __align__(16) uint8_t lstm_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int R = p.R, U = p.U, UPC = R / kLstmCluster;  // units per CTA (multiple of 8)
  const int WS = 4 * UPC + 8;                          // sW row stride (bf16): odd multiple of 16 B
  const int HS = R + 8;                                // sH row stride
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(lstm_smem);          // [R][WS]   W_hh[:, own gate columns]
  __nv_bfloat16* sH = sW + (size_t)R * WS;                                   // [2][32][HS]  h_{t-1} (read) / h_t (written by all 8 CTAs)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g8 = lane >> 2, tq = lane & 3;
  const int b0 = blockIdx.y * kLstmRows;
  const int n_items = 2 * (UPC / 8);  // (m-tile, unit-tile) pairs: one per warp
  // ---- weights: local column gate * UPC + ul  <-  global column gate * R + rank * UPC + ul
  for (int i = threadIdx.x; i < R * 4 * UPC; i += blockDim.x) {
    const int k = i / (4 * UPC), n = i % (4 * UPC);
    const int gate = n / UPC, ul = n % UPC;
    sW[(size_t)k * WS + n] = __float2bfloat16(whh[(size_t)k * 4 * R + gate * R + rank * UPC + ul]);
  }
  for (int i = threadIdx.x; i < kLstmRows * R; i += blockDim.x) {
    const int r = i / R, k = i % R;
    const int b = b0 + r;
    sH[(size_t)r * HS + k] = __float2bfloat16((h0 && b < p.B) ? h0[(size_t)b * R + k] : 0.f);
  }
  __nv_bfloat16* peerH[kLstmCluster];  // the same buffer in every CTA of the cluster (DSMEM)
#pragma unroll
  for (int j = 0; j < kLstmCluster; ++j) peerH[j] = cluster.map_shared_rank(sH, j);
  const bool active = warp < n_items;
  const int mt = warp & 1, ut = warp >> 1;      // m-tile, unit-tile of this warp
  const int rowA = mt * 16 + g8, rowB = rowA + 8;  // rows inside the 32-row block
  const int bA = b0 + rowA, bB = b0 + rowB;
  const int ul0 = ut * 8 + 2 * tq;              // local unit of this thread's column pair
  const int ug0 = rank * UPC + ul0;             // global unit
  float cA[2] = {0.f, 0.f}, cB[2] = {0.f, 0.f};
  if (active && c0) {
    if (bA < p.B) { cA[0] = c0[(size_t)bA * R + ug0]; cA[1] = c0[(size_t)bA * R + ug0 + 1]; }
    if (bB < p.B) { cB[0] = c0[(size_t)bB * R + ug0]; cB[1] = c0[(size_t)bB * R + ug0 + 1]; }
  }
  const int lenA = (length && bA < p.B) ? length[bA] : U, lenB = (length && bB < p.B) ? length[bB] : U;
  float hA[2] = {0.f, 0.f}, hB[2] = {0.f, 0.f};
  // gx of the first step
  float2 gxa[4], gxb[4];
  auto load_gx = [&](int t) {
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) {
      gxa[gate] = gxb[gate] = make_float2(0.f, 0.f);
      if (active && t < U) {
        if (bA < p.B) gxa[gate] = *reinterpret_cast<const float2*>(gx + ((size_t)bA * U + t) * 4 * R + gate * R + ug0);
        if (bB < p.B) gxb[gate] = *reinterpret_cast<const float2*>(gx + ((size_t)bB * U + t) * 4 * R + gate * R + ug0);
      }
    }
  };
  load_gx(0);
  __syncthreads();
  cluster.sync();  // every CTA of the cluster is running (DSMEM may be touched from here on)
  for (int t = 0; t < U; ++t) {
    if (active) {
      float acc[4][4];
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) acc[gate][0] = acc[gate][1] = acc[gate][2] = acc[gate][3] = 0.f;
      const __nv_bfloat16* hcur = sH + (size_t)(t & 1) * kLstmRows * HS;
#ifdef LSTM_PROBE_NO_MMA
      for (int ks = 0; ks < 0; ++ks) {
#else
#pragma unroll 4
      for (int ks = 0; ks < R / 16; ++ks) {
#endif
        uint32_t afr[4];
        mma::ldsm_x4(afr, hcur + (size_t)(mt * 16 + (lane & 15)) * HS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          uint32_t bfr[2];
          mma::ldsm_x2_trans(bfr, sW + (size_t)(ks * 16 + (lane & 15)) * WS + gate * UPC + ut * 8);
          mma::mma_16816(acc[gate], afr, bfr);
        }
      }
      const float2 ga[4] = {gxa[0], gxa[1], gxa[2], gxa[3]}, gb[4] = {gxb[0], gxb[1], gxb[2], gxb[3]};
      load_gx(t + 1);  // next step's input projection: in flight during the cell update and the exchange
      float sv[2][2][5];  // [row A/B][unit 0/1][i f g o c]
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        {
          const float i_ = sigmoid_acc(acc[0][e] + (e ? ga[0].y : ga[0].x)), f_ = sigmoid_acc(acc[1][e] + (e ? ga[1].y : ga[1].x));
          const float g_ = tanh_fast(acc[2][e] + (e ? ga[2].y : ga[2].x)), o_ = sigmoid_acc(acc[3][e] + (e ? ga[3].y : ga[3].x));
          cA[e] = f_ * cA[e] + i_ * g_;
          hA[e] = o_ * tanh_fast(cA[e]);
          sv[0][e][0] = i_; sv[0][e][1] = f_; sv[0][e][2] = g_; sv[0][e][3] = o_; sv[0][e][4] = cA[e];
        }
        {
          const float i_ = sigmoid_acc(acc[0][2 + e] + (e ? gb[0].y : gb[0].x)), f_ = sigmoid_acc(acc[1][2 + e] + (e ? gb[1].y : gb[1].x));
          const float g_ = tanh_fast(acc[2][2 + e] + (e ? gb[2].y : gb[2].x)), o_ = sigmoid_acc(acc[3][2 + e] + (e ? gb[3].y : gb[3].x));
          cB[e] = f_ * cB[e] + i_ * g_;
          hB[e] = o_ * tanh_fast(cB[e]);
          sv[1][e][0] = i_; sv[1][e][1] = f_; sv[1][e][2] = g_; sv[1][e][3] = o_; sv[1][e][4] = cB[e];
        }
      }
      // publish h_t: straight into the next-step buffer of all 8 CTAs (fire-and-forget DSMEM stores, ordered by the cluster barrier)
      {
        const __nv_bfloat162 va = __floats2bfloat162_rn(hA[0], hA[1]), vb = __floats2bfloat162_rn(hB[0], hB[1]);
        const size_t offA = (size_t)((t + 1) & 1) * kLstmRows * HS + (size_t)rowA * HS + ug0;
        const size_t offB = (size_t)((t + 1) & 1) * kLstmRows * HS + (size_t)rowB * HS + ug0;
#pragma unroll
#ifdef LSTM_PROBE_NO_PUSH
        for (int j = rank; j <= rank; ++j) {
#else
        for (int j = 0; j < kLstmCluster; ++j) {
#endif
          *reinterpret_cast<__nv_bfloat162*>(peerH[j] + offA) = va;
          *reinterpret_cast<__nv_bfloat162*>(peerH[j] + offB) = vb;
        }
      }
      if (bA < p.B) {
        const bool on = t < lenA;
        *reinterpret_cast<__nv_bfloat162*>(y + ((size_t)bA * U + t) * R + ug0) = __floats2bfloat162_rn(on ? hA[0] : 0.f, on ? hA[1] : 0.f);
        saved_g[((size_t)bA * U + t) * (R / 2) + (ug0 >> 1)] =
            make_uint4(pack_bf2(sv[0][0][0], sv[0][1][0]), pack_bf2(sv[0][0][1], sv[0][1][1]), pack_bf2(sv[0][0][2], sv[0][1][2]), pack_bf2(sv[0][0][3], sv[0][1][3]));
        *reinterpret_cast<float2*>(saved_c + ((size_t)bA * U + t) * R + ug0) = make_float2(sv[0][0][4], sv[0][1][4]);
      }
      if (bB < p.B) {
        const bool on = t < lenB;
        *reinterpret_cast<__nv_bfloat162*>(y + ((size_t)bB * U + t) * R + ug0) = __floats2bfloat162_rn(on ? hB[0] : 0.f, on ? hB[1] : 0.f);
        saved_g[((size_t)bB * U + t) * (R / 2) + (ug0 >> 1)] =
            make_uint4(pack_bf2(sv[1][0][0], sv[1][1][0]), pack_bf2(sv[1][0][1], sv[1][1][1]), pack_bf2(sv[1][0][2], sv[1][1][2]), pack_bf2(sv[1][0][3], sv[1][1][3]));
        *reinterpret_cast<float2*>(saved_c + ((size_t)bB * U + t) * R + ug0) = make_float2(sv[1][0][4], sv[1][1][4]);
      }
    }
#ifdef LSTM_PROBE_NO_SYNC
    __syncthreads();
#else
    cluster.sync();  // h_t is complete in every CTA; nobody reads h_{t-1} any more
#endif
  }
  if (active) {
    if (bA < p.B) {
      if (hT) *reinterpret_cast<float2*>(hT + (size_t)bA * R + ug0) = make_float2(hA[0], hA[1]);
      if (cT) *reinterpret_cast<float2*>(cT + (size_t)bA * R + ug0) = make_float2(cA[0], cA[1]);
    }
    if (bB < p.B) {
      if (hT) *reinterpret_cast<float2*>(hT + (size_t)bB * R + ug0) = make_float2(hB[0], hB[1]);
      if (cT) *reinterpret_cast<float2*>(cT + (size_t)bB * R + ug0) = make_float2(cB[0], cB[1]);
    }
  }
  cluster.sync();  // no CTA leaves while a peer may still read its SMEM
}

// ---------------------------------------------------------------------------------------------- backward
// dy [B, U, R] bf16 (gradient of the masked output), saved as written by the forward, c0 [B, R] or null, whh [R, 4R] f32
// -> dgates [B, U, 4R] bf16 (gradient of the pre-activations: feeds dW_ih, dW_hh, dbias, dx as plain GEMMs), dh0 / dc0 [B, R] f32 or null
// Per step: cell backward for the CTA's own units (registers) -> its [32 x 4 R/8] gate-gradient slice in local SMEM -> partial
// dh_{t-1}[32 x R] = slice . W_hh[:, own gate columns]^T by mma.sync -> each 8-unit tile of the partial sums is pushed (DSMEM) to the CTA
// that owns those units (reduce-scatter) -> cluster barrier -> the owner adds the 8 partials.  Exchange: 5 KB per CTA pair and step.
__global__ void __cluster_dims__(kLstmCluster, 1, 1) __launch_bounds__(512, 1)
    lstm_rec_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint4* __restrict__ saved_g, const float* __restrict__ saved_c,
                        const float* __restrict__ c0,
                        const float* __restrict__ whh, const int32_t* __restrict__ length, __nv_bfloat16* __restrict__ dgates,
                        float* __restrict__ dh0, float* __restrict__ dc0, const LstmParams p) {
  extern __shared__ __align__(16) uint8_t lstm_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int R = p.R, U = p.U, UPC = R / kLstmCluster;
  const int G = 4 * R, GL = 4 * UPC;  // all gate columns / this CTA's gate columns
  const int WS = GL + 8;              // row stride of sW and sOwn (bf16)
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(lstm_smem);  // [R][WS]   W_hh[r', own gate columns]  (B operand, k = gate column)
  __nv_bfloat16* sOwn = sW + (size_t)R * WS;                         // [32][WS]  gate gradients of this CTA's units (A operand)
  float* sSlot = reinterpret_cast<float*>(sOwn + (size_t)kLstmRows * WS);  // [2][8 sources][32][UPC] partial dh for this CTA's units
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g8 = lane >> 2, tq = lane & 3;
  const int b0 = blockIdx.y * kLstmRows;
  const int n_items = 2 * (UPC / 8);
  for (int i = threadIdx.x; i < R * GL; i += blockDim.x) {
    const int k = i / GL, n = i % GL;
    const int gate = n / UPC, ul = n % UPC;
    sW[(size_t)k * WS + n] = __float2bfloat16(whh[(size_t)k * G + gate * R + rank * UPC + ul]);
  }
  const bool active = warp < n_items;
  const int mt = warp & 1, ut = warp >> 1;
  const int rowA = mt * 16 + g8, rowB = rowA + 8;
  const int bA = b0 + rowA, bB = b0 + rowB;
  const int ul0 = ut * 8 + 2 * tq, ug0 = rank * UPC + ul0;
  const int lenA = (length && bA < p.B) ? length[bA] : U, lenB = (length && bB < p.B) ? length[bB] : U;
  float dcA[2] = {0.f, 0.f}, dcB[2] = {0.f, 0.f};  // dL/dc flowing to the previous step
  float dhr[4] = {0.f, 0.f, 0.f, 0.f};             // recurrent dL/dh for this thread's (row, unit) pairs
  // MMA work split: warp w < 2 * 5 takes m-tile (w & 1) and the 8 unit tiles {w >> 1, w >> 1 + 5, ...} of R / 8 = 40 (UPC / 8 per owner)
  const int n_tiles_total = R / 8, n_warps_mma = n_items;  // same 10 warps
  __syncthreads();
  cluster.sync();
  for (int t = U - 1; t >= 0; --t) {
    if (active) {
      float dgA[2][4], dgB[2][4];  // [unit][gate] pre-activation gradients
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int b = half ? bB : bA;
        const int len = half ? lenB : lenA;
        float* dc = half ? dcB : dcA;
        float(*dg)[4] = half ? dgB : dgA;
        float2 sv[5] = {}, cp = make_float2(0.f, 0.f);
        float2 dyv = make_float2(0.f, 0.f);
        if (b < p.B) {
          const uint4 gq = saved_g[((size_t)b * U + t) * (R / 2) + (ug0 >> 1)];
          sv[0] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.x));
          sv[1] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.y));
          sv[2] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.z));
          sv[3] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.w));
          sv[4] = *reinterpret_cast<const float2*>(saved_c + ((size_t)b * U + t) * R + ug0);
          if (t > 0) cp = *reinterpret_cast<const float2*>(saved_c + ((size_t)b * U + t - 1) * R + ug0);
          else if (c0) cp = *reinterpret_cast<const float2*>(c0 + (size_t)b * R + ug0);
          if (t < len) dyv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + ((size_t)b * U + t) * R + ug0));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float i_ = e ? sv[0].y : sv[0].x, f_ = e ? sv[1].y : sv[1].x, g_ = e ? sv[2].y : sv[2].x, o_ = e ? sv[3].y : sv[3].x;
          const float c_ = e ? sv[4].y : sv[4].x, cprev = e ? cp.y : cp.x;
          const float dh = (e ? dyv.y : dyv.x) + dhr[2 * half + e];
          const float tc = tanh_fast(c_);
          const float dct = dh * o_ * (1.0f - tc * tc) + dc[e];
          dg[e][0] = dct * g_ * i_ * (1.0f - i_);
          dg[e][1] = dct * cprev * f_ * (1.0f - f_);
          dg[e][2] = dct * i_ * (1.0f - g_ * g_);
          dg[e][3] = dh * tc * o_ * (1.0f - o_);
          dc[e] = dct * f_;
        }
      }
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) {
        const __nv_bfloat162 va = __floats2bfloat162_rn(dgA[0][gate], dgA[1][gate]), vb = __floats2bfloat162_rn(dgB[0][gate], dgB[1][gate]);
        *reinterpret_cast<__nv_bfloat162*>(sOwn + (size_t)rowA * WS + gate * UPC + ul0) = va;
        *reinterpret_cast<__nv_bfloat162*>(sOwn + (size_t)rowB * WS + gate * UPC + ul0) = vb;
        if (bA < p.B) *reinterpret_cast<__nv_bfloat162*>(dgates + ((size_t)bA * U + t) * G + gate * R + ug0) = va;
        if (bB < p.B) *reinterpret_cast<__nv_bfloat162*>(dgates + ((size_t)bB * U + t) * G + gate * R + ug0) = vb;
      }
    }
    __syncthreads();  // the CTA's gate-gradient slice is complete
    if (warp < n_warps_mma) {
      // partial dh_{t-1}[row, r'] over this CTA's gate columns, for every unit r'
      float* slot_base = nullptr;
      for (int nt = ut; nt < n_tiles_total; nt += UPC / 8) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int ks = 0; ks < GL / 16; ++ks) {
          uint32_t afr[4], bfr[2];
          mma::ldsm_x4(afr, sOwn + (size_t)(mt * 16 + (lane & 15)) * WS + ks * 16 + (lane >> 4) * 8);
          mma::ldsm_x2(bfr, sW + (size_t)(nt * 8 + (lane & 7)) * WS + ks * 16 + ((lane >> 3) & 1) * 8);
          mma::mma_16816(acc, afr, bfr);
        }
        // unit tile nt belongs to CTA `owner`; this CTA's partial goes into slot [t & 1][rank] there
        const int owner = (nt * 8) / UPC, ul = (nt * 8) % UPC + 2 * tq;
        slot_base = cluster.map_shared_rank(sSlot, owner) + ((size_t)((t & 1) * kLstmCluster + rank) * kLstmRows) * UPC;
        *reinterpret_cast<float2*>(slot_base + (size_t)rowA * UPC + ul) = make_float2(acc[0], acc[1]);
        *reinterpret_cast<float2*>(slot_base + (size_t)rowB * UPC + ul) = make_float2(acc[2], acc[3]);
      }
    }
    cluster.sync();  // all partial sums of this step have arrived; sOwn may be rewritten
    if (active) {
      dhr[0] = dhr[1] = dhr[2] = dhr[3] = 0.f;
#pragma unroll
      for (int src = 0; src < kLstmCluster; ++src) {
        const float* sl = sSlot + ((size_t)((t & 1) * kLstmCluster + src) * kLstmRows) * UPC;
        const float2 a = *reinterpret_cast<const float2*>(sl + (size_t)rowA * UPC + ul0), b = *reinterpret_cast<const float2*>(sl + (size_t)rowB * UPC + ul0);
        dhr[0] += a.x;
        dhr[1] += a.y;
        dhr[2] += b.x;
        dhr[3] += b.y;
      }
    }
  }
  if (active) {
    if (bA < p.B) {
      if (dh0) *reinterpret_cast<float2*>(dh0 + (size_t)bA * R + ug0) = make_float2(dhr[0], dhr[1]);
      if (dc0) *reinterpret_cast<float2*>(dc0 + (size_t)bA * R + ug0) = make_float2(dcA[0], dcA[1]);
    }
    if (bB < p.B) {
      if (dh0) *reinterpret_cast<float2*>(dh0 + (size_t)bB * R + ug0) = make_float2(dhr[2], dhr[3]);
      if (dc0) *reinterpret_cast<float2*>(dc0 + (size_t)bB * R + ug0) = make_float2(dcB[0], dcB[1]);
    }
  }
  cluster.sync();
}

static size_t lstm_fwd_smem(int R) {
  const int UPC = R / kLstmCluster;
  return ((size_t)R * (4 * UPC + 8) + (size_t)2 * kLstmRows * (R + 8)) * 2;
}
static size_t lstm_bwd_smem(int R) {
  const int UPC = R / kLstmCluster;
  return ((size_t)R * (4 * UPC + 8) + (size_t)kLstmRows * (4 * UPC + 8)) * 2 + (size_t)2 * kLstmCluster * kLstmRows * UPC * 4;
}
static bool lstm_supported(int R) { return R > 0 && R % 64 == 0 && 2 * (R / 64) <= 16 && lstm_fwd_smem(R) <= 227 * 1024 && lstm_bwd_smem(R) <= 227 * 1024; }

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_lstm_supported(int32_t R) { return lstm_supported(R) ? 1 : 0; }

extern "C" int tfasr_lstm_rec_fwd(const float* gx, const float* whh, const float* h0, const float* c0, const int32_t* length, void* y, void* saved_g,
                                  float* saved_c, float* hT, float* cT, int32_t B, int32_t U, int32_t R, tfasr_stream_t stream) {
  if (!gx || !whh || !y || !saved_g || !saved_c || B < 0 || U < 1 || !lstm_supported(R)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const size_t smem = lstm_fwd_smem(R);
  if (cudaFuncSetAttribute(lstm_rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  LstmParams p{B, U, R};
  dim3 grid(kLstmCluster, (unsigned)ceil_div(B, kLstmRows));
  lstm_rec_fwd_kernel<<<grid, 512, smem, static_cast<cudaStream_t>(stream)>>>(gx, whh, h0, c0, length, static_cast<__nv_bfloat16*>(y), static_cast<uint4*>(saved_g),
                                                                              saved_c, hT, cT, p);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_lstm_rec_bwd(const void* dy, const void* saved_g, const float* saved_c, const float* c0, const float* whh, const int32_t* length,
                                  void* dgates, float* dh0, float* dc0, int32_t B, int32_t U, int32_t R, tfasr_stream_t stream) {
  if (!dy || !saved_g || !saved_c || !whh || !dgates || B < 0 || U < 1 || !lstm_supported(R)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  const size_t smem = lstm_bwd_smem(R);
  if (cudaFuncSetAttribute(lstm_rec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TFASR_EXECUTION_FAILED;
  LstmParams p{B, U, R};
  dim3 grid(kLstmCluster, (unsigned)ceil_div(B, kLstmRows));
  lstm_rec_bwd_kernel<<<grid, 512, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const uint4*>(saved_g), saved_c, c0, whh,
                                                                              length,
                                                                              static_cast<__nv_bfloat16*>(dgates), dh0, dc0, p);
  count_launch();
  return check_launch();
}
