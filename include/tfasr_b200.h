/*
 * libtfasr_b200 — C ABI of the B200-native TensorFlowASR training hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and enqueues work on the
 * caller's CUDA stream only (asynchronous; no implicit synchronisation, no global mutable state,
 * re-entrant).  All data pointers are DEVICE pointers to row-major contiguous buffers owned by the
 * caller (the library never allocates, frees or retains them), except where a parameter is marked
 * HOST.  Return value: tfasr_status_t (0 = ok).
 *
 * What each entry point replaces in the reference (paths relative to the TensorFlowASR tree):
 *   tfasr_logmel_*        tensorflow_asr/models/layers/feature_extraction.py:255-303 (FeatureExtraction.call:
 *                         preemphasis :170-175, stft :192-212, log_mel_spectrogram :220-231, spectrogram :233-235)
 *   tfasr_rnnt_loss_*     the warp-transducer FFI the reference sits behind, tensorflow_asr/losses/impl/rnnt.py:8,55
 *                         (`warprnnt_tensorflow.rnnt_loss(acts, labels, input_lengths, label_lengths, blank_label)`)
 *                         and its pure-TF twin compute_rnnt_loss_and_grad_helper, impl/rnnt.py:181-278
 *   tfasr_ctc_loss_*      tf.nn.ctc_loss as called at tensorflow_asr/losses/ctc_loss.py:57-66
 *   (further groups are declared below next to the reference lines they replace)
 */
#ifndef TFASR_B200_H_
#define TFASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tfasr_stream_t; /* a cudaStream_t / CUstream */

typedef enum {
  TFASR_OK = 0,
  TFASR_INVALID_VALUE = 1,      /* bad shape / unsupported configuration / blank != 0 */
  TFASR_EXECUTION_FAILED = 2,   /* kernel launch or CUDA runtime failure */
  TFASR_WORKSPACE_TOO_SMALL = 3 /* caller-provided workspace is smaller than *_workspace_bytes() */
} tfasr_status_t;

int tfasr_version(void);                 /* major*10000 + minor*100 + patch */
const char* tfasr_status_string(int st); /* static string */
/* Number of CUDA kernels this library has launched since load / the last reset (bench.py's `gpu_launches`). */
int tfasr_launch_count(void);
void tfasr_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Log-mel frontend (FeatureExtraction).  One fused kernel: pre-emphasis -> framing -> periodic Hann
 * -> zero-pad -> real FFT -> |.|^2 -> mel (or first-F-bins "spectrogram") -> log(x + eps).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t sample_rate;      /* 16000 */
  int32_t frame_length;     /* round(sample_rate * frame_ms / 1000)  = 400 */
  int32_t frame_step;       /* round(sample_rate * stride_ms / 1000) = 160 */
  int32_t nfft;             /* 256 | 512 | 1024 */
  int32_t num_feature_bins; /* 80 */
  int32_t feature_type;     /* 0 = log_mel_spectrogram, 1 = spectrogram (feature_extraction.py:233-235) */
  int32_t pad_end;          /* 1: ceil(N/step) frames (tf.signal.frame pad_end=True) */
  int32_t log_base10;       /* 0: natural log, 1: log10 (math_util.log10) */
  int32_t padding;          /* zero samples appended before feature extraction (:277-278) */
  float preemphasis;        /* 0.97; <= 0 disables */
  float epsilon;            /* 1e-6 */
  float lower_edge_hertz;   /* 0 */
  float upper_edge_hertz;   /* 8000 */
} tfasr_logmel_cfg;

/* Size in bytes of the constant table (window, twiddles, banded mel filterbank) for `cfg`. 0 if cfg is unsupported. */
size_t tfasr_logmel_table_bytes(const tfasr_logmel_cfg* cfg);
/* Fills `table_host` (HOST, tfasr_logmel_table_bytes bytes).  The caller copies it to the device once. */
int tfasr_logmel_build_table(const tfasr_logmel_cfg* cfg, void* table_host);
/* Number of output frames for `nsamples` input samples (feature_extraction.py:305-313, + cfg->padding). */
int32_t tfasr_logmel_nframes(const tfasr_logmel_cfg* cfg, int32_t nsamples);
/*
 * signal [B, N] f32, signal_len [B] i32 (only used for out_len, like the reference),
 * table = device copy of the table, out [B, T, F] f32 with T = tfasr_logmel_nframes(cfg, N), out_len [B] i32 (may be NULL).
 */
int tfasr_logmel_f32(const float* signal, const int32_t* signal_len, int32_t B, int32_t N, const tfasr_logmel_cfg* cfg,
                     const void* table, float* out, int32_t* out_len, tfasr_stream_t stream);
/* Optional per-utterance normalisations (feature_extraction.py:164-168 and :177-190); in place. */
int tfasr_signal_normalize_f32(float* signal, int32_t B, int32_t N, float epsilon, tfasr_stream_t stream);
int tfasr_feature_normalize_f32(float* feat, int32_t B, int32_t T, int32_t F, int32_t mode /*0 zscore, 1 minmax*/,
                                float epsilon, float log_eps_value, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * RNN-T loss + gradient (warp-transducer-shaped).  logits are RAW activations (log-softmax is
 * computed inside, like warp-transducer's GPU path, impl/rnnt.py:53-55).  blank must be 0
 * (losses/base_loss.py:24).  The caller applies the length clamp of base_loss.py:36 (the Python
 * RnntLoss mirror does).  labels [B, U1-1] i32; logit_len / label_len [B] i32.
 *   costs [B] f32                 = -log P(labels | logits)
 *   grads [B, T, U1, V] f32       = d costs[b] / d logits   (may alias `logits`: in-place)
 * Workspace (caller-owned, tfasr_rnnt_workspace_bytes) holds lse / blank / truth / alpha / beta.
 * ---------------------------------------------------------------------------------------------- */
size_t tfasr_rnnt_workspace_bytes(int32_t B, int32_t T, int32_t U1, int32_t V);
/* forward + gradient in one call (the warp-transducer contract); grads may be NULL (loss only). */
int tfasr_rnnt_loss_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                        int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs, float* grads,
                        void* workspace, size_t workspace_bytes, tfasr_stream_t stream);
/* two-phase variant used by the autograd wrapper: fwd fills costs + workspace; bwd (later, same workspace) writes
 * grads scaled by grad_scale[b] (NULL = 1), i.e. the custom-gradient closure of impl/rnnt.py:321-324 fused in. */
int tfasr_rnnt_loss_fwd_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                            int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs, void* workspace,
                            size_t workspace_bytes, tfasr_stream_t stream);
int tfasr_rnnt_loss_bwd_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                            int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs,
                            const float* grad_scale, float* grads, const void* workspace, size_t workspace_bytes,
                            tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CTC loss + gradient (tf.nn.ctc_loss, dense labels, batch-major logits [B, T, V], blank_index 0).
 *   costs [B] f32; grads [B, T, V] f32 = d costs[b] / d logits (may alias logits).  Infeasible
 *   alignments give +inf cost and zero gradient.
 * ---------------------------------------------------------------------------------------------- */
size_t tfasr_ctc_workspace_bytes(int32_t B, int32_t T, int32_t L, int32_t V);
int tfasr_ctc_loss_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                       int32_t B, int32_t T, int32_t L, int32_t V, int32_t blank, float* costs, float* grads,
                       const float* grad_scale, void* workspace, size_t workspace_bytes, tfasr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TFASR_B200_H_ */
