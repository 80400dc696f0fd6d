// SpecAugment (time / frequency masking) on the log-mel features, in place, one kernel, device-side RNG (CUDA-graph safe).
//
// Replaces Augmentation.feature_augment (tensorflow_asr/augmentations/augmentation.py:67-90: tf.map_fn over utterances) with
// FreqMasking.augment / TimeMasking.augment (augmentations/methods/specaugment.py:58-137: per mask a tf.where over the spectrogram).
// Per utterance, per mask:   frequency: f ~ U{0..F-1}, f = min(f, nbins), f0 ~ U{0..nbins-f-1}          -> bins [f0, f0+f) = value
//                            time:      Tm = floor(len * p_upperbound), t ~ U{0..Tm-1}, t = min(t, len), t0 ~ U{0..len-t-1}
// each applied with probability `prob`.  Frequency masks are drawn first (the reference sorts augmentations by name).
// The random stream is a counter hash of (*seed_ptr + salt, utterance, mask index): the distribution matches the reference, the
// individual draws cannot (TensorFlow's Philox stream) — parity for this row is statistical (tests/test_gpu_specaugment.py).
#include "common.cuh"

namespace tfasr {

__device__ __forceinline__ uint32_t sa_hash(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c) {
  unsigned long long k = seed + (unsigned long long)a * 0x9E3779B97F4A7C15ull + (unsigned long long)b * 0xC2B2AE3D27D4EB4Full + c * 0x165667B19E3779F9ull;
  k ^= k >> 33;
  k *= 0xFF51AFD7ED558CCDull;
  k ^= k >> 33;
  k *= 0xC4CEB9FE1A85EC53ull;
  k ^= k >> 33;
  return (uint32_t)(k >> 16);
}
__device__ __forceinline__ float sa_uniform(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c) {
  return (float)(sa_hash(seed, a, b, c) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ int sa_randint(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c, int n) {  // U{0..n-1}, 0 if n <= 0
  if (n <= 0) return 0;
  const int v = (int)(sa_uniform(seed, a, b, c) * (float)n);
  return v >= n ? n - 1 : v;
}

constexpr int kSaMaxMasks = 32;

// grid (ceil(T / 32), B); block 256
__global__ void __launch_bounds__(256) specaugment_kernel(float* __restrict__ feat, const int32_t* __restrict__ length, int T, int F, int n_time,
                                                          float p_upperbound, float time_prob, int n_freq, int freq_factor, float freq_prob,
                                                          float mask_value, const unsigned long long* __restrict__ seed_ptr,
                                                          unsigned long long salt) {
  __shared__ int s_t0[kSaMaxMasks], s_t1[kSaMaxMasks], s_f0[kSaMaxMasks], s_f1[kSaMaxMasks];
  const int b = blockIdx.y;
  const unsigned long long seed = (seed_ptr ? *seed_ptr : 0ull) + salt;
  const int len = min(length ? length[b] : T, T);
  if (threadIdx.x < n_freq) {
    const int m = threadIdx.x;
    const bool apply = sa_uniform(seed, b, m, 0) <= freq_prob;
    int f = min(sa_randint(seed, b, m, 1, freq_factor), F);
    int f0 = sa_randint(seed, b, m, 2, F - f);
    if (!apply) f = 0, f0 = 0;
    s_f0[m] = f0;
    s_f1[m] = f0 + f;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + n_time) {
    const int m = threadIdx.x - 32;
    const bool apply = sa_uniform(seed, b, 64 + m, 0) <= time_prob;
    const int Tm = (int)floorf((float)len * p_upperbound);
    int t = min(sa_randint(seed, b, 64 + m, 1, Tm), len);
    int t0 = sa_randint(seed, b, 64 + m, 2, len - t);
    if (!apply) t = 0, t0 = 0;
    s_t0[m] = t0;
    s_t1[m] = t0 + t;
  }
  __syncthreads();
  const int t_lo = blockIdx.x * 32;
  float* fb = feat + (size_t)b * T * F;
  for (int i = threadIdx.x; i < 32 * F; i += blockDim.x) {
    const int t = t_lo + i / F, f = i % F;
    if (t >= T) break;
    bool masked = false;
    for (int m = 0; m < n_freq; ++m) masked |= (f >= s_f0[m] && f < s_f1[m]);
    for (int m = 0; m < n_time; ++m) masked |= (t >= s_t0[m] && t < s_t1[m]);
    if (masked) fb[(size_t)t * F + f] = mask_value;
  }
}

}  // namespace tfasr

using namespace tfasr;

extern "C" int tfasr_specaugment_f32(float* feat, const int32_t* length, int32_t B, int32_t T, int32_t F, int32_t n_time, float p_upperbound,
                                     float time_prob, int32_t n_freq, int32_t freq_factor, float freq_prob, float mask_value,
                                     const unsigned long long* seed_ptr, unsigned long long salt, tfasr_stream_t stream) {
  if (!feat || B < 0 || T < 1 || F < 1 || n_time < 0 || n_freq < 0 || n_time > kSaMaxMasks || n_freq > kSaMaxMasks || B > 65535)
    return TFASR_INVALID_VALUE;
  if (B == 0 || (n_time == 0 && n_freq == 0)) return TFASR_OK;
  specaugment_kernel<<<dim3((unsigned)ceil_div(T, 32), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      feat, length, T, F, n_time, p_upperbound, time_prob, n_freq, freq_factor, freq_prob, mask_value, seed_ptr, salt);
  count_launch();
  return check_launch();
}
