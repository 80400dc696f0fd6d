// Fused log-mel frontend for sm_100a.
//
// Replaces tensorflow_asr/models/layers/feature_extraction.py:255-303 (FeatureExtraction.call):
//   preemphasis_signal :170-175 -> tf.signal.stft(pad_end) :192-212 -> |.|^2 -> mel matmul :220-231 (or first-F bins
//   "spectrogram" :233-235) -> log(x + eps) :214-218.  The reference runs this as ~10 TF kernels (cuFFT, cuBLAS SGEMM
//   on a 97 %-sparse filterbank, several elementwise passes); here it is ONE kernel that reads each sample once
//   (+ a 240-sample halo per 16-frame tile) and writes each feature once.
//
// Roofline: HBM-bound.  Algorithmic bytes per 10 s utterance = 4*160000 (read) + 4*1000*80 (write) = 960 000 B.
//
// Layout per CTA (256 threads): a tile of FR consecutive frames of one utterance.
//   1. coalesced float4 load of the tile's samples -> pre-emphasis -> SMEM
//   2. frames packed in pairs into complex sequences z = f_even + i f_odd (window applied, zero padded to NFFT)
//   3. Stockham autosort FFT in SMEM, radix-8 passes (radix 4 / 2 first pass for NFFT 256 / 1024), twiddles from SMEM
//   4. split the pair spectrum, power -> SMEM
//   5. banded mel filterbank from SMEM (<= 2 non-zero weights per FFT bin), + eps, log -> coalesced store
#include "common.cuh"

#include <math.h>
#include <string.h>

#include <vector>

namespace tfasr {

struct LogmelLayout {
  int window_off, tw_off, mel_start_off, mel_count_off, mel_woff_off, mel_w_off, total_words;
};

static bool logmel_cfg_ok(const tfasr_logmel_cfg* c) {
  if (!c) return false;
  if (c->nfft != 256 && c->nfft != 512 && c->nfft != 1024) return false;
  if (c->frame_length < 2 || c->frame_length > c->nfft) return false;
  if (c->frame_step < 1 || c->frame_step > c->frame_length) return false;
  if (c->num_feature_bins < 1) return false;
  if (c->feature_type == 1 && c->num_feature_bins > c->nfft / 2 + 1) return false;
  if (c->feature_type != 0 && c->feature_type != 1) return false;
  if (c->feature_type == 0 && c->num_feature_bins > 512) return false;
  if (c->padding < 0) return false;
  return true;
}

static LogmelLayout logmel_layout(const tfasr_logmel_cfg* c) {
  LogmelLayout l;
  int F = c->num_feature_bins;
  int nbins = c->nfft / 2 + 1;
  l.window_off = 0;
  l.tw_off = (int)align_up(c->frame_length, 4);
  l.mel_start_off = l.tw_off + 2 * c->nfft;
  l.mel_count_off = l.mel_start_off + F;
  l.mel_woff_off = l.mel_count_off + F;
  l.mel_w_off = l.mel_woff_off + F;
  l.total_words = (int)align_up(l.mel_w_off + 2 * nbins + 8, 4);
  return l;
}

// tf.signal.linear_to_mel_weight_matrix restated in fp32 (HTK mel, DC row zero), then band-compressed.
static void build_mel(const tfasr_logmel_cfg* c, std::vector<int>& start, std::vector<int>& count, std::vector<int>& woff,
                      std::vector<float>& w) {
  const int F = c->num_feature_bins, nb = c->nfft / 2 + 1;
  auto mel = [](float f) { return 1127.0f * logf(1.0f + f / 700.0f); };
  const float nyq = (float)c->sample_rate / 2.0f;
  std::vector<float> spec_mel(nb), edges(F + 2);
  // np.linspace / tf.linspace: start + i * step with step = (stop - start) / (n - 1)
  const float step_hz = nyq / (float)(nb - 1);
  for (int k = 0; k < nb; ++k) spec_mel[k] = mel(k == nb - 1 ? nyq : (float)k * step_hz);
  const float mlo = mel(c->lower_edge_hertz), mhi = mel(c->upper_edge_hertz);
  const float step_mel = (mhi - mlo) / (float)(F + 1);
  for (int i = 0; i < F + 2; ++i) edges[i] = (i == F + 1) ? mhi : mlo + (float)i * step_mel;
  start.assign(F, 0), count.assign(F, 0), woff.assign(F, 0), w.clear();
  for (int m = 0; m < F; ++m) {
    const float lo = edges[m], ce = edges[m + 1], hi = edges[m + 2];
    int first = -1, last = -1;
    std::vector<float> col(nb, 0.f);
    for (int k = 1; k < nb; ++k) {  // DC bin has zero weight (bands_to_zero = 1)
      const float ls = (spec_mel[k] - lo) / (ce - lo);
      const float us = (hi - spec_mel[k]) / (hi - ce);
      const float v = fmaxf(0.f, fminf(ls, us));
      col[k] = v;
      if (v > 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    woff[m] = (int)w.size();
    if (first >= 0) {
      start[m] = first;
      count[m] = last - first + 1;
      for (int k = first; k <= last; ++k) w.push_back(col[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ device
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
  float2 t = a;
  a = make_float2(t.x + b.x, t.y + b.y);
  b = make_float2(t.x - b.x, t.y - b.y);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

template <int R>
__device__ __forceinline__ void dft_r(float2* v);  // natural-order output in v
template <>
__device__ __forceinline__ void dft_r<2>(float2* v) {
  bfly2(v[0], v[1]);
}
template <>
__device__ __forceinline__ void dft_r<4>(float2* v) {
  bfly2(v[0], v[2]);
  bfly2(v[1], v[3]);
  v[3] = mul_mi(v[3]);
  bfly2(v[0], v[1]);
  bfly2(v[2], v[3]);
  float2 t = v[1];
  v[1] = v[2];
  v[2] = t;
}
template <>
__device__ __forceinline__ void dft_r<8>(float2* v) {
  const float s = 0.70710678118654752440f;
  bfly2(v[0], v[4]);
  bfly2(v[1], v[5]);
  bfly2(v[2], v[6]);
  bfly2(v[3], v[7]);
  v[5] = make_float2(s * (v[5].x + v[5].y), s * (v[5].y - v[5].x));    // * (1 - i)/sqrt2
  v[6] = mul_mi(v[6]);                                                  // * (-i)
  v[7] = make_float2(s * (v[7].y - v[7].x), s * (-v[7].x - v[7].y));   // * (-1 - i)/sqrt2
  bfly2(v[0], v[2]);
  bfly2(v[1], v[3]);
  bfly2(v[4], v[6]);
  bfly2(v[5], v[7]);
  v[3] = mul_mi(v[3]);
  v[7] = mul_mi(v[7]);
  bfly2(v[0], v[1]);
  bfly2(v[2], v[3]);
  bfly2(v[4], v[5]);
  bfly2(v[6], v[7]);
  // bit-reversed -> natural: X = {v0, v4, v2, v6, v1, v5, v3, v7}
  float2 t1 = v[1], t3 = v[3];
  v[1] = v[4];
  v[3] = v[6];
  v[4] = t1;
  v[6] = t3;
}

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }

// One Stockham pass of radix R over NF independent N-point FFTs held in (s_re, s_im) (padded index).
// Every thread first gathers all of its butterflies' inputs, then the CTA syncs, then scatters (in-place safe).
template <int N, int R, int NF, int THREADS>
__device__ __forceinline__ void stockham_pass(float* s_re, float* s_im, const float2* __restrict__ s_tw, int Ns, int tid) {
  constexpr int PER_FFT = N / R;
  constexpr int ITEMS = NF * PER_FFT;
  constexpr int ITERS = (ITEMS + THREADS - 1) / THREADS;
  constexpr int STRIDE = N + N / 32;  // padded length of one FFT buffer
  float2 v[ITERS][R];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int item = tid + it * THREADS;
    if (ITEMS % THREADS == 0 || item < ITEMS) {
      const int f = item / PER_FFT, j = item % PER_FFT;
      const int k = j % Ns;
      const int tw_step = k * (N / (Ns * R));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int idx = f * STRIDE + pad_idx(j + r * PER_FFT);
        float2 x = make_float2(s_re[idx], s_im[idx]);
        v[it][r] = (r == 0) ? x : cmul(x, s_tw[(r * tw_step) & (N - 1)]);
      }
      dft_r<R>(v[it]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int item = tid + it * THREADS;
    if (ITEMS % THREADS == 0 || item < ITEMS) {
      const int f = item / PER_FFT, j = item % PER_FFT;
      const int k = j % Ns;
      const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int idx = f * STRIDE + pad_idx(j0 + r * Ns);
        s_re[idx] = v[it][r].x;
        s_im[idx] = v[it][r].y;
      }
    }
  }
  __syncthreads();
}

template <int N, int NF, int THREADS>
__device__ __forceinline__ void fft_all(float* s_re, float* s_im, const float2* s_tw, int tid) {
  if constexpr (N == 512) {
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 1, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 8, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 64, tid);
  } else if constexpr (N == 256) {
    stockham_pass<N, 4, NF, THREADS>(s_re, s_im, s_tw, 1, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 4, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 32, tid);
  } else {
    stockham_pass<N, 2, NF, THREADS>(s_re, s_im, s_tw, 1, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 2, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 16, tid);
    stockham_pass<N, 8, NF, THREADS>(s_re, s_im, s_tw, 128, tid);
  }
}

struct LogmelParams {
  const float* signal;
  const float* table;
  float* out;
  int B, N, Ntot, T, F;
  int frame_length, frame_step;
  int feature_type, log_base10;
  float preemph, eps;
  LogmelLayout lay;
};

constexpr int kLogmelThreads = 256;

template <int N>
struct LogmelTile {
  static constexpr int NF = (N == 1024) ? 4 : 8;  // complex FFTs per CTA
  static constexpr int FR = 2 * NF;               // frames per CTA
  static constexpr int NB = N / 2 + 1;
  static constexpr int STRIDE = N + N / 32;
};

template <int N>
__global__ void __launch_bounds__(kLogmelThreads) logmel_kernel(const LogmelParams p) {
  using TL = LogmelTile<N>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TL::FR;
  const int nfr = min(TL::FR, p.T - t0);
  const int span = (TL::FR - 1) * p.frame_step + p.frame_length;  // samples covered by a full tile

  // smem carve-up
  float* s_tab = smem;                                       // table copy (window | twiddles | mel)
  float* s_sig = s_tab + p.lay.total_words;                  // span floats (pre-emphasised)
  float* s_re = s_sig + (int)align_up(span, 4);              // NF * STRIDE
  float* s_im = s_re + TL::NF * TL::STRIDE;                  // NF * STRIDE
  float* s_pow = s_im + TL::NF * TL::STRIDE;                 // FR * (NB + 1)
  const float* s_win = s_tab + p.lay.window_off;
  const float2* s_tw = reinterpret_cast<const float2*>(s_tab + p.lay.tw_off);

  // 0. constant table -> smem (L2-resident, 128-bit loads)
  {
    const float4* src = reinterpret_cast<const float4*>(p.table);
    float4* dst = reinterpret_cast<float4*>(s_tab);
    for (int i = tid; i < p.lay.total_words / 4; i += kLogmelThreads) dst[i] = __ldg(src + i);
  }
  // 1. samples [n0, n0 + span) of utterance b, pre-emphasised: y[n] = x[n] - c x[n-1], y[0] = x[0];
  //    x is zero beyond N (the `padding` zeros and the pad_end tail), y is zero beyond Ntot.
  {
    const float* x = p.signal + (size_t)b * p.N;
    const int n0 = t0 * p.frame_step;
    const bool vec_ok = ((p.N & 3) == 0) && ((n0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.signal) & 15) == 0);
    const float c = p.preemph;
    if (vec_ok) {
      for (int i4 = tid; i4 < (span + 3) / 4; i4 += kLogmelThreads) {
        const int n = n0 + i4 * 4;
        float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n + 3 < p.N) {
          cur = __ldg(reinterpret_cast<const float4*>(x + n));
        } else {
          if (n < p.N) cur.x = x[n];
          if (n + 1 < p.N) cur.y = x[n + 1];
          if (n + 2 < p.N) cur.z = x[n + 2];
        }
        const float prev = (n > 0 && n - 1 < p.N) ? __ldg(x + n - 1) : 0.f;
        float4 y;
        y.x = (n == 0) ? cur.x : cur.x - c * prev;
        y.y = cur.y - c * cur.x;
        y.z = cur.z - c * cur.y;
        y.w = cur.w - c * cur.z;
        if (n >= p.Ntot) y.x = 0.f;
        if (n + 1 >= p.Ntot) y.y = 0.f;
        if (n + 2 >= p.Ntot) y.z = 0.f;
        if (n + 3 >= p.Ntot) y.w = 0.f;
        const int i = i4 * 4;
        if (i + 3 < (int)align_up(span, 4)) {
          *reinterpret_cast<float4*>(s_sig + i) = y;
        }
      }
    } else {
      for (int i = tid; i < span; i += kLogmelThreads) {
        const int n = n0 + i;
        const float cur = (n < p.N) ? x[n] : 0.f;
        const float prev = (n > 0 && n - 1 < p.N) ? x[n - 1] : 0.f;
        float y = (n == 0) ? cur : cur - c * prev;
        if (n >= p.Ntot) y = 0.f;
        s_sig[i] = y;
      }
    }
  }
  __syncthreads();
  // 2. windowed frame pairs -> complex FFT input (re = even frame, im = odd frame), zero padded to N
  for (int i = tid; i < TL::NF * N; i += kLogmelThreads) {
    const int f = i / N, n = i % N;
    float re = 0.f, im = 0.f;
    if (n < p.frame_length) {
      const float w = s_win[n];
      const int fe = 2 * f, fo = 2 * f + 1;
      if (fe < nfr) re = s_sig[fe * p.frame_step + n] * w;
      if (fo < nfr) im = s_sig[fo * p.frame_step + n] * w;
    }
    const int idx = f * TL::STRIDE + pad_idx(n);
    s_re[idx] = re;
    s_im[idx] = im;
  }
  __syncthreads();
  // 3. FFT
  fft_all<N, TL::NF, kLogmelThreads>(s_re, s_im, s_tw, tid);
  // 4. split the packed pair and take |.|^2 (abs then square in the reference: same value up to 1 ulp)
  for (int i = tid; i < TL::NF * TL::NB; i += kLogmelThreads) {
    const int f = i / TL::NB, k = i % TL::NB;
    const int ik = f * TL::STRIDE + pad_idx(k);
    const int in = f * TL::STRIDE + pad_idx((N - k) & (N - 1));
    const float a = s_re[ik], bb = s_im[ik], cc = s_re[in], d = s_im[in];
    const float p1 = 0.25f * ((a + cc) * (a + cc) + (bb - d) * (bb - d));
    const float p2 = 0.25f * ((a - cc) * (a - cc) + (bb + d) * (bb + d));
    s_pow[(2 * f) * (TL::NB + 1) + k] = p1;
    s_pow[(2 * f + 1) * (TL::NB + 1) + k] = p2;
  }
  __syncthreads();
  // 5. mel / spectrogram + log, coalesced store of [nfr, F]
  const float inv_ln10 = 0.43429448190325176f;
  float* out = p.out + ((size_t)b * p.T + t0) * p.F;
  if (p.feature_type == 0) {
    const int* s_start = reinterpret_cast<const int*>(s_tab + p.lay.mel_start_off);
    const int* s_count = reinterpret_cast<const int*>(s_tab + p.lay.mel_count_off);
    const int* s_woff = reinterpret_cast<const int*>(s_tab + p.lay.mel_woff_off);
    const float* s_w = s_tab + p.lay.mel_w_off;
    for (int i = tid; i < nfr * p.F; i += kLogmelThreads) {
      const int fr = i / p.F, m = i % p.F;
      const float* pw = s_pow + fr * (TL::NB + 1) + s_start[m];
      const float* w = s_w + s_woff[m];
      const int cnt = s_count[m];
      float acc = 0.f;
      for (int k = 0; k < cnt; ++k) acc = fmaf(pw[k], w[k], acc);
      float v = logf(acc + p.eps);
      if (p.log_base10) v *= inv_ln10;
      out[i] = v;
    }
  } else {
    for (int i = tid; i < nfr * p.F; i += kLogmelThreads) {
      const int fr = i / p.F, m = i % p.F;
      float v = logf(s_pow[fr * (TL::NB + 1) + m] + p.eps);
      if (p.log_base10) v *= inv_ln10;
      out[i] = v;
    }
  }
}

__global__ void logmel_len_kernel(const int32_t* in_len, int32_t* out_len, int B, int frame_length, int frame_step, int pad_end) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const int n = in_len[i];
    // floor division like Python's // (feature_extraction.py:305-313)
    int v;
    if (pad_end) {
      v = (n + frame_step - 1) / frame_step;
      if (n <= 0) v = -((-n) / frame_step);
    } else {
      const int d = n - frame_length;
      v = 1 + (d >= 0 ? d / frame_step : -((-d + frame_step - 1) / frame_step));
    }
    out_len[i] = v;
  }
}

// per-utterance gain 1 / (max|x| + eps), in place (feature_extraction.py:164-168)
__global__ void signal_normalize_kernel(float* signal, int N, float eps) {
  float* x = signal + (size_t)blockIdx.x * N;
  __shared__ float s_red[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) s_red[0] = v;
  }
  __syncthreads();
  const float gain = 1.0f / (s_red[0] + eps);
  for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] *= gain;
}

// per-(utterance, feature bin) normalisation over time (feature_extraction.py:177-190); block = (b, f-chunk)
__global__ void feature_normalize_kernel(float* feat, int T, int F, int mode, float eps, float log_eps_value) {
  const int b = blockIdx.x;
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float* x = feat + (size_t)b * T * F + f;
  if (mode == 0) {
    float mean = 0.f;
    for (int t = 0; t < T; ++t) mean += x[(size_t)t * F];
    mean /= (float)T;
    float var = 0.f;
    for (int t = 0; t < T; ++t) {
      const float d = x[(size_t)t * F] - mean;
      var += d * d;
    }
    var /= (float)T;
    const float inv = 1.0f / sqrtf(var + eps);
    for (int t = 0; t < T; ++t) x[(size_t)t * F] = (x[(size_t)t * F] - mean) * inv;
  } else {
    float mx = -CUDART_INF_F;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, x[(size_t)t * F]);
    const float mn = log_eps_value;
    const float inv = 1.0f / (mx - mn);
    for (int t = 0; t < T; ++t) x[(size_t)t * F] = (x[(size_t)t * F] - mn) * inv;
  }
}

template <int N>
static int launch_logmel(const LogmelParams& p, cudaStream_t stream) {
  using TL = LogmelTile<N>;
  const int span = (TL::FR - 1) * p.frame_step + p.frame_length;
  const size_t smem = sizeof(float) * ((size_t)p.lay.total_words + align_up(span, 4) + 2 * TL::NF * TL::STRIDE + TL::FR * (TL::NB + 1));
  if (smem > 227 * 1024) return TFASR_INVALID_VALUE;
  static bool attr_set = false;  // idempotent; benign if raced
  if (!attr_set) {
    if (cudaFuncSetAttribute(logmel_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return TFASR_EXECUTION_FAILED;
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(p.T, TL::FR), (unsigned)p.B);
  logmel_kernel<N><<<grid, kLogmelThreads, smem, stream>>>(p);
  count_launch();
  return check_launch();
}

}  // namespace tfasr

using namespace tfasr;

extern "C" size_t tfasr_logmel_table_bytes(const tfasr_logmel_cfg* cfg) {
  if (!logmel_cfg_ok(cfg)) return 0;
  return sizeof(float) * (size_t)logmel_layout(cfg).total_words;
}

extern "C" int tfasr_logmel_build_table(const tfasr_logmel_cfg* cfg, void* table_host) {
  if (!logmel_cfg_ok(cfg) || !table_host) return TFASR_INVALID_VALUE;
  const LogmelLayout l = logmel_layout(cfg);
  float* t = static_cast<float*>(table_host);
  memset(t, 0, sizeof(float) * l.total_words);
  // tf.signal.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi k / (n + even - 1))
  const int n = cfg->frame_length;
  const int denom = n + (1 - n % 2) - 1;
  for (int k = 0; k < n; ++k) t[l.window_off + k] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)k / (double)denom));
  for (int m = 0; m < cfg->nfft; ++m) {
    const double a = 2.0 * M_PI * (double)m / (double)cfg->nfft;
    t[l.tw_off + 2 * m] = (float)cos(a);
    t[l.tw_off + 2 * m + 1] = (float)(-sin(a));
  }
  if (cfg->feature_type == 0) {
    std::vector<int> start, count, woff;
    std::vector<float> w;
    build_mel(cfg, start, count, woff, w);
    if ((int)w.size() > 2 * (cfg->nfft / 2 + 1) + 8) return TFASR_INVALID_VALUE;
    int* ti = reinterpret_cast<int*>(t);
    for (int m = 0; m < cfg->num_feature_bins; ++m) {
      ti[l.mel_start_off + m] = start[m];
      ti[l.mel_count_off + m] = count[m];
      ti[l.mel_woff_off + m] = woff[m];
    }
    for (size_t i = 0; i < w.size(); ++i) t[l.mel_w_off + i] = w[i];
  }
  return TFASR_OK;
}

extern "C" int32_t tfasr_logmel_nframes(const tfasr_logmel_cfg* cfg, int32_t nsamples) {
  const int n = nsamples + cfg->padding;
  if (cfg->pad_end) return (n + cfg->frame_step - 1) / cfg->frame_step;
  if (n < cfg->frame_length) return 0;
  return 1 + (n - cfg->frame_length) / cfg->frame_step;
}

extern "C" int tfasr_logmel_f32(const float* signal, const int32_t* signal_len, int32_t B, int32_t N, const tfasr_logmel_cfg* cfg,
                                const void* table, float* out, int32_t* out_len, tfasr_stream_t stream_) {
  if (!logmel_cfg_ok(cfg) || !signal || !table || !out || B < 0 || N < 1) return TFASR_INVALID_VALUE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = tfasr_logmel_nframes(cfg, N);
  if (out_len) {
    if (!signal_len) return TFASR_INVALID_VALUE;
    if (B > 0) {
      logmel_len_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, stream>>>(signal_len, out_len, B, cfg->frame_length, cfg->frame_step, cfg->pad_end);
      count_launch();
    }
  }
  if (B == 0 || T <= 0) return check_launch();
  if (B > 65535) return TFASR_INVALID_VALUE;
  LogmelParams p;
  p.signal = signal;
  p.table = static_cast<const float*>(table);
  p.out = out;
  p.B = B;
  p.N = N;
  p.Ntot = N + cfg->padding;
  p.T = T;
  p.F = cfg->num_feature_bins;
  p.frame_length = cfg->frame_length;
  p.frame_step = cfg->frame_step;
  p.feature_type = cfg->feature_type;
  p.log_base10 = cfg->log_base10;
  p.preemph = cfg->preemphasis > 0.f ? cfg->preemphasis : 0.f;
  p.eps = cfg->epsilon;
  p.lay = logmel_layout(cfg);
  switch (cfg->nfft) {
    case 256: return launch_logmel<256>(p, stream);
    case 512: return launch_logmel<512>(p, stream);
    case 1024: return launch_logmel<1024>(p, stream);
  }
  return TFASR_INVALID_VALUE;
}

extern "C" int tfasr_signal_normalize_f32(float* signal, int32_t B, int32_t N, float epsilon, tfasr_stream_t stream) {
  if (!signal || B < 0 || N < 1) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  signal_normalize_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(signal, N, epsilon);
  count_launch();
  return check_launch();
}

extern "C" int tfasr_feature_normalize_f32(float* feat, int32_t B, int32_t T, int32_t F, int32_t mode, float epsilon,
                                           float log_eps_value, tfasr_stream_t stream) {
  if (!feat || B < 0 || T < 1 || F < 1 || (mode != 0 && mode != 1)) return TFASR_INVALID_VALUE;
  if (B == 0) return TFASR_OK;
  dim3 grid(B, (unsigned)ceil_div(F, 128));
  feature_normalize_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(feat, T, F, mode, epsilon, log_eps_value);
  count_launch();
  return check_launch();
}
