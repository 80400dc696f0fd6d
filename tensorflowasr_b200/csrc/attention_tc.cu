// Relative-position multi-head self-attention core, FORWARD, on tcgen05 (5th-generation tensor cores, TMEM accumulators, TMA-fed operands).
//
// Reference: MultiHeadRelativeAttention._compute_attention (tensorflow_asr/models/layers/multihead_attention.py:543-582), rel_left_shift
// (:27-64), the masked softmax (general.py:30-41) and compute_streaming_mask (:104-143).  Same contract as tfasr_rel_attention_fwd[_ld]
// (csrc/attention.cu, the mma.sync generation this kernel replaces in the forward pass): context + the normalised probabilities in bf16.
//
// One CTA per (utterance, head, 128-query tile), 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = softmax.
//   scores[t, j] = scale * ((q_t + u) . k_j + (q_t + v) . pos[T - 1 - t + j])           (the rel-shift is an index identity)
// Keys are walked in chunks of 64.  Per chunk the MMA warp issues
//   S  [128 x  64] = Qu . K_chunk^T                       (tcgen05.mma 128 x 64 x 16, K = head dim padded to 48 / 64)
//   BD [128 x 192] = Qv . pos[rlo .. rlo + 191]^T         rlo = T - 1 - (t0 + 127) + j0: every relative position the tile x chunk can touch
// into TMEM.  Row i of the tile needs BD column (127 - i) + jj for key jj of the chunk — a per-ROW column offset, which tcgen05.ld cannot
// express (its column address is warp-uniform).  The warp-uniform part of the offset goes into the tcgen05.ld address, the per-lane rest
// (31 - lane) is resolved through a transposed per-warp SMEM scratch (bank = lane on both the write and the read: conflict-free).
// Softmax is two-pass over the chunks (the MMAs are nearly free, the [128 x T] fp32 scores do not fit anywhere): pass 0 = running row max /
// sum, pass 1 = recompute the scores, write the NORMALISED probabilities (bf16) to global (the backward reads them) and into SMEM as the A
// operand of   O [128 x HP] += P_chunk . V_chunk   (V as an MN-major B operand: no transpose).
// The head dimension 36 is not a multiple of the MMA K = 16, and a head's first column (36 n) is not always a multiple of 8 — but a TMA box must
// start on a 16-byte boundary of the row (an unaligned start raised an illegal-instruction exception on the B200).  So every box starts at
// the aligned column below the head, the head's data sits at columns dl .. dl + H - 1 of the tile (dl = 0 or 4), and Qu / Qv are written by
// this kernel at the same columns with zeros everywhere else: whatever the K / pos boxes carry outside the head is multiplied by zero.
//
// Roofline: latency / issue — 18 MFLOP per (utterance, head) at T = 250; the point of tcgen05 here is that the tensor work and the operand
// movement cost no instruction slots of the 8 softmax warps (the mma.sync generation spent 94 % of its issue slots on index arithmetic).
#include "gemm_kernel.cuh"

namespace tfasr {

constexpr int kAttTcThreads = 320;
constexpr int kKC = 64;    // keys per chunk
constexpr int kPW = 192;   // relative positions per (tile, chunk): 128 + 64 - 1, rounded up to the MMA N granularity

struct AttTcArgs {
  const float *ubias, *vbias;  // [N, H]
  const int32_t* length;       // [B] or null
  __nv_bfloat16* ctx;          // [B*T, N*H]
  __nv_bfloat16* probs;        // [B, N, T, T]
  int B, T, N, H, chunk, history;
  float scale;
};

template <int HP>  // head dimension padded to a multiple of 16: 48 (H <= 48) or 64
struct AttTcSmem {
  static constexpr int Q_BYTES = 128 * 128;          // one [128 x 64] bf16 tile
  static constexpr int K_BYTES = kKC * 128;          // [64 keys x 64]
  static constexpr int P_BYTES = kPW * 128;          // [192 positions x 64]
  static constexpr int STAGE_BYTES = 2 * K_BYTES + P_BYTES;  // K | V | pos window
  static constexpr int STAGES = 2;
  static constexpr int SCR_BYTES = 8 * 32 * 32 * 4;  // per softmax warp: [32 keys][32 lanes] f32
  static constexpr int STAT_BYTES = 128 * 2 * 2 * 4; // [row][half] (max, sum)
  static constexpr int BAR_BYTES = 32 * 8;
  // sQraw is reused as the first probability tile once Qu / Qv are built
  static constexpr int TOTAL = 4 * Q_BYTES + STAGES * STAGE_BYTES + SCR_BYTES + STAT_BYTES + BAR_BYTES + 1024;
  static_assert(TOTAL <= 227 * 1024, "smem budget");
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t S_COL = 0, BD_COL = 128, O_COL = 320;
};

template <int HP>
__global__ void __launch_bounds__(kAttTcThreads, 1)
rel_attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos, const AttTcArgs a) {
  using S = AttTcSmem<HP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = sm100::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQraw = smem;                      // q tile as loaded; later probability tile 0
  uint8_t* sQu = sQraw + S::Q_BYTES;
  uint8_t* sQv = sQu + S::Q_BYTES;
  uint8_t* sPr1 = sQv + S::Q_BYTES;           // probability tile 1
  uint8_t* sStage = sPr1 + S::Q_BYTES;        // [STAGES][K | V | pos]
  float* sScr = reinterpret_cast<float*>(sStage + S::STAGES * S::STAGE_BYTES);
  float* sStat = sScr + S::SCR_BYTES / 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + S::STAT_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_ready = bars + 1;
  uint64_t* o_full = bars + 2;
  uint64_t* bd_full = bars + 3;
  uint64_t* bd_empty = bars + 4;
  uint64_t* st_full = bars + 5;    // [2] stage loaded
  uint64_t* st_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;     // [2]
  uint64_t* s_empty = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;    // [2]
  uint64_t* p_empty = bars + 15;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, H = a.H, N = a.N;
  const int R = 2 * T - 1;
  const int bn = blockIdx.y, b = bn / N, n = bn % N;
  const int t0 = blockIdx.x * 128;
  const int dl = (n * H) & 7;        // the head starts `dl` columns into its (16-byte aligned) TMA boxes
  const int c_head = n * H - dl;     // aligned first column of the q box; k / v / pos boxes are offset likewise (N * H % 8 == 0)
  const int len_b = a.length ? min(max(a.length[b], 0), T) : T;
  const bool streaming = a.chunk > 0;
  // key chunks this tile has to visit: all of them unless the streaming window of every row in the tile (all of them valid queries) ends earlier /
  // starts later.  Rows that are masked queries attend uniformly to ALL keys (-1e9 everywhere), so such tiles visit everything.
  const int nkc_all = (T + kKC - 1) / kKC;
  int kc_lo = 0, kc_hi = nkc_all;
  if (streaming && min(t0 + 128, T) <= len_b) {
    const int tl = t0, th = min(t0 + 127, T - 1);
    const int hist = a.history < 0 ? T : a.history;
    const int lo = max((tl / a.chunk) * a.chunk - hist, 0), hi = min((th / a.chunk) * a.chunk + a.chunk, T);
    kc_lo = lo / kKC;
    kc_hi = (hi + kKC - 1) / kKC;
  }
  const int nkc = kc_hi - kc_lo;

  if (warp == 0 && lane == 0) {
    sm100::prefetch_tensormap(&tmQKV);
    sm100::prefetch_tensormap(&tmPos);
    sm100::mbar_init(q_full, 1);
    sm100::mbar_init(q_ready, 1);
    sm100::mbar_init(o_full, 1);
    sm100::mbar_init(bd_full, 1);
    sm100::mbar_init(bd_empty, 8);
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&st_full[i], 1);
      sm100::mbar_init(&st_empty[i], 1);
      sm100::mbar_init(&s_full[i], 1);
      sm100::mbar_init(&s_empty[i], 8);
      sm100::mbar_init(&p_full[i], 8);
      sm100::mbar_init(&p_empty[i], 1);
    }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc(tmem_slot, S::TMEM_COLS);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_it = 2 * nkc;  // pass 0 (statistics) + pass 1 (probabilities, context)

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      sm100::mbar_arrive_expect_tx(q_full, S::Q_BYTES);
      sm100::tma_load_2d(sQraw, &tmQKV, q_full, c_head, b * T + t0);
      sm100::tma_load_2d(sQraw + 64 * 128, &tmQKV, q_full, c_head, b * T + t0 + 64);
      for (int it = 0; it < total_it; ++it) {
        const int pass = it / nkc, kc = kc_lo + it % nkc;
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        sm100::mbar_wait(&st_empty[s], ph ^ 1);
        uint8_t* st = sStage + s * S::STAGE_BYTES;
        sm100::mbar_arrive_expect_tx(&st_full[s], S::K_BYTES + S::P_BYTES + (pass ? S::K_BYTES : 0));
        sm100::tma_load_2d(st, &tmQKV, &st_full[s], N * H + c_head, b * T + kc * kKC);                        // K chunk
        if (pass) sm100::tma_load_2d(st + S::K_BYTES, &tmQKV, &st_full[s], 2 * N * H + c_head, b * T + kc * kKC);  // V chunk
        const int rlo = T - 1 - (t0 + 127) + kc * kKC;  // may be negative / run past R: those rows are never addressed by a valid (t, j)
#pragma unroll
        for (int k = 0; k < kPW / 64; ++k) sm100::tma_load_2d(st + 2 * S::K_BYTES + k * 64 * 128, &tmPos, &st_full[s], c_head, b * R + rlo + k * 64);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = sm100::make_idesc_bf16_f32(128, kKC, 0, 0);
      constexpr uint32_t idesc_bd = sm100::make_idesc_bf16_f32(128, kPW, 0, 0);
      // B = V chunk, MN-major.  N = 64 whatever the head size: an MN-major operand is addressed in whole 64-element swizzle atoms (N = 48
      // raised an illegal-instruction exception on the B200); columns beyond the head are never read back from TMEM.
      constexpr uint32_t idesc_o = sm100::make_idesc_bf16_f32(128, 64, 0, 1);
      const uint32_t qu = sm100::smem_u32(sQu), qv = sm100::smem_u32(sQv);
      sm100::mbar_wait(q_ready, 0);
      sm100::tc_fence_after();
      int pv_done = 0;  // probability tiles consumed
      auto pv = [&](int it) {  // O += P(it) . V(it): probability tile it & 1, stage it & 1
        const int j = it - nkc;  // index within pass 1
        const uint32_t pb = j & 1, pph = (j >> 1) & 1, s = it & 1;
        sm100::mbar_wait(&p_full[pb], pph);
        sm100::tc_fence_after();
        const uint32_t pa = sm100::smem_u32(pb ? sPr1 : sQraw);
        const uint32_t vb = sm100::smem_u32(sStage + s * S::STAGE_BYTES + S::K_BYTES);
#pragma unroll
        for (int k = 0; k < kKC / 16; ++k)
          sm100::umma_bf16(tmem_base + S::O_COL, sm100::make_smem_desc(pa + k * 32, 16, 1024), sm100::make_smem_desc(vb + k * 2048, 64 * 128, 1024), idesc_o,
                           (j | k) != 0 ? 1u : 0u);
        sm100::umma_commit(&p_empty[pb]);
        sm100::umma_commit(&st_empty[s]);
        ++pv_done;
      };
      for (int it = 0; it < total_it; ++it) {
        const int pass = it / nkc;
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        sm100::mbar_wait(&st_full[s], ph);
        sm100::mbar_wait(&s_empty[s], ph ^ 1);
        sm100::mbar_wait(bd_empty, (it & 1) ^ 1);
        sm100::tc_fence_after();
        const uint32_t kb = sm100::smem_u32(sStage + s * S::STAGE_BYTES);
        const uint32_t pb = kb + 2 * S::K_BYTES;
#pragma unroll
        for (int k = 0; k < HP / 16; ++k)
          sm100::umma_bf16(tmem_base + S::S_COL + s * kKC, sm100::make_smem_desc(qu + k * 32, 16, 1024), sm100::make_smem_desc(kb + k * 32, 16, 1024), idesc_s,
                           k != 0 ? 1u : 0u);
        sm100::umma_commit(&s_full[s]);
#pragma unroll
        for (int k = 0; k < HP / 16; ++k)
          sm100::umma_bf16(tmem_base + S::BD_COL, sm100::make_smem_desc(qv + k * 32, 16, 1024), sm100::make_smem_desc(pb + k * 32, 16, 1024), idesc_bd,
                           k != 0 ? 1u : 0u);
        sm100::umma_commit(bd_full);
        if (!pass) sm100::umma_commit(&st_empty[s]);  // pass 0 is done with the stage once both products have read it
        if (pass && it > nkc) pv(it - 1);
      }
      pv(total_it - 1);
      sm100::umma_commit(o_full);
    }
  } else {
    // ===================================================================== softmax warps (2..9)
    const int q = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int row_in_tile = q * 32 + lane;
    const int t = t0 + row_in_tile;
    const bool row_ok = t < T;
    const bool q_valid = t < len_b;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float* scr = sScr + ew * (32 * 32);
    if (ew < 4) {
      // ---- Qu = q + u, Qv = q + v (bf16), zero beyond the head dimension; one row per thread
      sm100::mbar_wait(q_full, 0);
      const float* ub = a.ubias + n * H;
      const float* vb = a.vbias + n * H;
#pragma unroll
      for (int c = 0; c < HP / 8; ++c) {
        const uint4 u4 = *reinterpret_cast<const uint4*>(sQraw + swz_off(row_in_tile, c));
        const uint32_t w[4] = {u4.x, u4.y, u4.z, u4.w};
        uint32_t ou[4], ov[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h0 = c * 8 + 2 * e - dl;  // head column of this pair (dl is even: pairs never straddle the head's edge)
          const bool in = h0 >= 0 && h0 + 1 < H;
          const float q0 = __uint_as_float(w[e] << 16), q1 = __uint_as_float(w[e] & 0xFFFF0000u);
          const float u0 = in ? q0 + __ldg(ub + h0) : 0.f, u1 = in ? q1 + __ldg(ub + h0 + 1) : 0.f;
          const float v0 = in ? q0 + __ldg(vb + h0) : 0.f, v1 = in ? q1 + __ldg(vb + h0 + 1) : 0.f;
          ou[e] = pack2(u0, u1);
          ov[e] = pack2(v0, v1);
        }
        *reinterpret_cast<uint4*>(sQu + swz_off(row_in_tile, c)) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<uint4*>(sQv + swz_off(row_in_tile, c)) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      }
      sm100::fence_proxy_async();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (ew == 0 && lane == 0) sm100::mbar_arrive(q_ready);
    }
    // streaming window of this query
    int win_lo = 0, win_hi = T;
    if (streaming && row_ok) {
      const int hist = a.history < 0 ? T : a.history;
      const int qc = (t / a.chunk) * a.chunk;
      win_lo = max(qc - hist, 0);
      win_hi = min(qc + a.chunk, T);
    }
    float m_run = -CUDART_INF_F, l_run = 0.f, m_fin = 0.f, inv_l = 0.f;
    __nv_bfloat16* prow = a.probs + (((size_t)b * N + n) * T + (row_ok ? t : 0)) * T;
    const bool t_even = (T & 1) == 0;
    // pass 1 writes zeros for the key chunks a streaming tile skips
    if (row_ok && nkc < nkc_all) {
      for (int kc = 0; kc < nkc_all; ++kc) {
        if (kc >= kc_lo && kc < kc_hi) continue;
        for (int jj = half * 32; jj < half * 32 + 32; ++jj)
          if (kc * kKC + jj < T) prow[kc * kKC + jj] = __float2bfloat16(0.f);
      }
    }
    for (int it = 0; it < total_it; ++it) {
      const int pass = it / nkc, kc = kc_lo + it % nkc;
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      // ---- BD: the 64 columns this warp's 32 rows x 32 keys can touch; resolve the per-lane shift through the transposed scratch
      sm100::mbar_wait(bd_full, it & 1);
      sm100::tc_fence_after();
      {
        const uint32_t base = S::BD_COL + (96 - 32 * q + 32 * half);
        uint32_t r[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) sm100::tmem_ld_32x32b_x16(tmem_row + base + g * 16, r[g]);
        sm100::tmem_ld_wait();
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(bd_empty);
        const int sh = 31 - lane;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int jj = g * 16 + e - sh;  // key (within this warp's 32) that column g*16+e serves for this lane
            if (jj >= 0 && jj < 32) scr[jj * 32 + lane] = __uint_as_float(r[g][e]);
          }
      }
      // ---- S
      sm100::mbar_wait(&s_full[s], ph);
      sm100::tc_fence_after();
      uint32_t sr[2][16];
      sm100::tmem_ld_32x32b_x16(tmem_row + S::S_COL + s * kKC + half * 32, sr[0]);
      sm100::tmem_ld_32x32b_x16(tmem_row + S::S_COL + s * kKC + half * 32 + 16, sr[1]);
      sm100::tmem_ld_wait();
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&s_empty[s]);
      float sc[32];
      const int j0 = kc * kKC + half * 32;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = j0 + jj;
        float v = (__uint_as_float(sr[jj >> 4][jj & 15]) + scr[jj * 32 + lane]) * a.scale;
        if (!(q_valid && j >= win_lo && j < win_hi)) v = -1e9f;  // masked softmax: general.py:30-41 fills with -1e9
        sc[jj] = (j < T) ? v : -CUDART_INF_F;                    // keys beyond the utterance tensor do not exist
      }
      if (pass == 0) {
        float mc = -CUDART_INF_F;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) mc = fmaxf(mc, sc[jj]);
        if (mc > -CUDART_INF_F) {
          const float mn = fmaxf(m_run, mc);
          float acc = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) acc += __expf(sc[jj] - mn);
          l_run = fmaf(l_run, __expf(m_run - mn), acc);
          m_run = mn;
        }
        if (it == nkc - 1) {
          // combine the two halves of every row: (max, sum) through SMEM
          sStat[(row_in_tile * 2 + half) * 2] = m_run;
          sStat[(row_in_tile * 2 + half) * 2 + 1] = l_run;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float mo = sStat[(row_in_tile * 2 + (half ^ 1)) * 2], lo = sStat[(row_in_tile * 2 + (half ^ 1)) * 2 + 1];
          m_fin = fmaxf(m_run, mo);
          const float lt = (m_run > -CUDART_INF_F ? l_run * __expf(m_run - m_fin) : 0.f) + (mo > -CUDART_INF_F ? lo * __expf(mo - m_fin) : 0.f);
          inv_l = lt > 0.f ? 1.0f / lt : 0.f;
        }
      } else {
        const int jp = it - nkc;
        const uint32_t pb = jp & 1, pph = (jp >> 1) & 1;
        uint32_t pk[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) pk[jj >> 1] = pack2(__expf(sc[jj] - m_fin) * inv_l, __expf(sc[jj + 1] - m_fin) * inv_l);
        if (row_ok) {
          if (t_even) {
            uint32_t* pg = reinterpret_cast<uint32_t*>(prow + j0);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (j0 + 2 * jj + 1 < T) pg[jj] = pk[jj];
              else if (j0 + 2 * jj < T) reinterpret_cast<unsigned short*>(prow)[j0 + 2 * jj] = (unsigned short)(pk[jj] & 0xFFFFu);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (j0 + jj < T) reinterpret_cast<unsigned short*>(prow)[j0 + jj] = (unsigned short)((pk[jj >> 1] >> ((jj & 1) * 16)) & 0xFFFFu);
          }
        }
        sm100::mbar_wait(&p_empty[pb], pph ^ 1);
        uint8_t* pt = pb ? sPr1 : sQraw;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(pt + swz_off(row_in_tile, half * 4 + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        sm100::fence_proxy_async();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&p_full[pb]);
      }
    }
    // ---- context
    sm100::mbar_wait(o_full, 0);
    sm100::tc_fence_after();
    if (half == 0) {
      uint32_t o[HP / 16][16];
#pragma unroll
      for (int g = 0; g < HP / 16; ++g) sm100::tmem_ld_32x32b_x16(tmem_row + S::O_COL + g * 16, o[g]);
      sm100::tmem_ld_wait();
      if (row_ok) {
        __nv_bfloat16* cp = a.ctx + ((size_t)b * T + t) * (N * H) + n * H;
#pragma unroll
        for (int g = 0; g < HP / 16; ++g)
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const int h0 = g * 16 + e - dl;  // accumulator column -> head column
            if (h0 >= 0 && h0 + 1 < H) *reinterpret_cast<uint32_t*>(cp + h0) = pack2(__uint_as_float(o[g][e]), __uint_as_float(o[g][e + 1]));
          }
      }
    }
    sm100::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int HP>
static int launch_att_tc(const CUtensorMap& tmQKV, const CUtensorMap& tmPos, const AttTcArgs& a, cudaStream_t stream) {
  using S = AttTcSmem<HP>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(rel_attention_fwd_tc_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  const dim3 grid((unsigned)((a.T + 127) / 128), (unsigned)(a.B * a.N));
  rel_attention_fwd_tc_kernel<HP><<<grid, kAttTcThreads, S::TOTAL, stream>>>(tmQKV, tmPos, a);
  count_launch();
  return check_launch();
}

}  // namespace tfasr

using namespace tfasr;

// 1 when the tcgen05 forward covers the shape: head dimension even and <= 64, 16-byte aligned row pitches.
extern "C" int tfasr_rel_attention_tc_supported(int32_t T, int32_t N, int32_t H, int32_t ld_qkv, int32_t ld_pos) {
  if (!(T >= 1 && N >= 1 && H >= 2 && (H & 1) == 0 && (ld_qkv % 8) == 0 && (ld_pos % 8) == 0 && ((N * H) % 8) == 0)) return 0;
  int dl_max = 0;  // a head starts up to 6 columns into its aligned box: the padded head dimension must hold dl + H
  for (int n = 0; n < N; ++n) dl_max = (n * H) % 8 > dl_max ? (n * H) % 8 : dl_max;
  return H + dl_max <= 64 ? 1 : 0;
}

extern "C" int tfasr_rel_attention_fwd_tc(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                                          const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx,
                                          void* probs, tfasr_stream_t stream) {
  if (!qkv || !pos || !ubias || !vbias || !ctx || !probs || B < 0) return TFASR_INVALID_VALUE;
  if (!tfasr_rel_attention_tc_supported(T, N, H, ld_qkv, ld_pos)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  CUtensorMap tmQKV, tmPos;
  // the q / k / v column groups of one head start at arbitrary (even) columns: TMA takes element coordinates, only the base and the row pitch
  // must be 16-byte aligned.  The box is 64 columns wide; what lies beyond the head's H columns is neutralised by the zeros in Qu / Qv.
  bool ok = make_tmap_bf16_2d(&tmQKV, qkv, (uint64_t)B * T, (uint64_t)ld_qkv, (uint64_t)ld_qkv, 64, 64);
  ok = ok && make_tmap_bf16_2d(&tmPos, pos, (uint64_t)B * (2 * T - 1), (uint64_t)ld_pos, (uint64_t)ld_pos, 64, 64);
  if (!ok) return TFASR_INVALID_VALUE;
  AttTcArgs a{};
  a.ubias = ubias;
  a.vbias = vbias;
  a.length = length;
  a.ctx = static_cast<__nv_bfloat16*>(ctx);
  a.probs = static_cast<__nv_bfloat16*>(probs);
  a.B = B, a.T = T, a.N = N, a.H = H, a.chunk = chunk, a.history = history;
  a.scale = 1.0f / sqrtf((float)H);
  int dl_max = 0;
  for (int n = 0; n < N; ++n) dl_max = (n * H) % 8 > dl_max ? (n * H) % 8 : dl_max;
  if (H + dl_max <= 48) return launch_att_tc<48>(tmQKV, tmPos, a, static_cast<cudaStream_t>(stream));
  return launch_att_tc<64>(tmQKV, tmPos, a, static_cast<cudaStream_t>(stream));
}
