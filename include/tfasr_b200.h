/*
 * libtfasr_b200 — C ABI of the B200-native TensorFlowASR training hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and enqueues work on the
 * caller's CUDA stream only (asynchronous; no implicit synchronisation, no global mutable state,
 * re-entrant).  All data pointers are DEVICE pointers to row-major contiguous buffers owned by the
 * caller (the library never allocates, frees or retains them), except where a parameter is marked
 * HOST.  Return value: tfasr_status_t (0 = ok).
 *
 * Launches: the frequent kernels are enqueued as programmatic dependent launches (cudaLaunchAttributeProgrammaticStreamSerialization);
 * each of them executes griddepcontrol.wait before its first global-memory access, so the stream order the caller sees is unchanged, and a
 * stream capture records programmatic edges.  TFASR_PDL=0 in the environment (read once) restores ordinary launches.
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

/* Forward without the log-sum-exp pass over the logits (first step): `rowstats` [B*T*U1, parts, 2] f32 are the per-slice
 * (max, sum exp(v - max)) pairs left by tfasr_gemm_bf16 with epilogue->rowstats set (parts = 2 * ceil(V / 256)).
 * tfasr_rnnt_lse_from_rowstats_f32 is the combining step alone (lse / blank / truth planes [B,T,U1] f32, rows outside the lattice untouched);
 * tfasr_rnnt_loss_fwd_rowstats_f32 = that step + the alpha / beta wavefronts, same outputs and workspace layout as tfasr_rnnt_loss_fwd_f32. */
int tfasr_rnnt_lse_from_rowstats_f32(const float* logits, const float* rowstats, int32_t parts, const int32_t* labels, const int32_t* logit_len,
                                     const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V, float* lse, float* blank,
                                     float* truth, tfasr_stream_t stream);
int tfasr_rnnt_loss_fwd_rowstats_f32(const float* logits, const float* rowstats, int32_t parts, const int32_t* labels, const int32_t* logit_len,
                                     const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs,
                                     void* workspace, size_t workspace_bytes, tfasr_stream_t stream);

/* Training-head variant of the backward: the gradient is written as bf16 with row pitch ld_grads (elements, multiple of 4) — the
 * operand format of the joint's backward GEMMs — and may be written IN PLACE over the fp32 logits (grads == logits, ld_grads = 2V).
 * dbias [V] (f32, may be NULL) accumulates the column sums (the vocab Dense bias gradient).  Requires V % 4 == 0, V <= 2048. */
int tfasr_rnnt_loss_bwd_bf16(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                             int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                             int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream);
/* second generation of the same kernel (csrc/rnnt_grad_v2.cu: rows processed in two halves, bias partial sums in warp-private shared memory:
 * 72 registers instead of 126, three CTAs per SM); identical contract.  Opt-in until measured on a B200. */
int tfasr_rnnt_loss_bwd_bf16_v2(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                                int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream);

/* bf16-logits variants (ops.bf16_logits; SURVEY 8f N1, second step): the joint GEMM stores bf16 logits and takes its row statistics over the
 * ROUNDED values, so the head computes the exact RNN-T loss and gradient of the stored logits; the [B,T,U1,V] tensor is 2 bytes per entry
 * and its gradient replaces it in place (grads == logits, ld_grads = ld_logits).  ld_logits = row pitch in elements (>= V, multiple of 4;
 * a multiple of 64 keeps every row on a 128-byte boundary for the TMA loads of the backward GEMMs).  Same workspace layout; V % 4 == 0,
 * V <= 2048; 16-byte accesses when V, both pitches are multiples of 8 and the bases are 16-byte aligned. */
int tfasr_rnnt_lse_from_rowstats_bf16(const void* logits, int64_t ld_logits, const float* rowstats, int32_t parts, const int32_t* labels, const int32_t* logit_len,
                                      const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V, float* lse, float* blank,
                                      float* truth, tfasr_stream_t stream);
int tfasr_rnnt_loss_fwd_rowstats_bf16(const void* logits, int64_t ld_logits, const float* rowstats, int32_t parts, const int32_t* labels, const int32_t* logit_len,
                                      const int32_t* label_len, int32_t B, int32_t T, int32_t U1, int32_t V, int32_t blank, float* costs,
                                      void* workspace, size_t workspace_bytes, tfasr_stream_t stream);
int tfasr_rnnt_loss_bwd_bf16_io(const void* logits, int64_t ld_logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len, int32_t B,
                                int32_t T, int32_t U1, int32_t V, int32_t blank, const float* costs, const float* grad_scale, void* grads,
                                int64_t ld_grads, float* dbias, const void* workspace, size_t workspace_bytes, tfasr_stream_t stream);

/* Alpha / beta wavefronts alone, second generation (csrc/rnnt_lattice_v2.cu; ops.lattice_v2): the utterance's blank / truth planes staged
 * in shared memory, direction as a template parameter.  blank / truth / alpha / beta are the [B,T,U1] f32 planes of the workspace
 * (tfasr_rnnt_lse_from_rowstats_* fills the first two); costs [B] = -beta[0,0] (+inf for an empty utterance).  Needs 8 T U1 + 256 bytes
 * of shared memory: tfasr_rnnt_lattice_smem_supported(T, U1) says whether the shape fits (else TFASR_INVALID_VALUE). */
int tfasr_rnnt_lattice_smem_supported(int32_t T, int32_t U1);
int tfasr_rnnt_lattice_smem_f32(const float* blank, const float* truth, const int32_t* logit_len, const int32_t* label_len, int32_t B, int32_t T,
                                int32_t U1, float* alpha, float* beta, float* costs, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CTC loss + gradient (tf.nn.ctc_loss, dense labels, batch-major logits [B, T, V], blank_index 0).
 *   costs [B] f32; grads [B, T, V] f32 = d costs[b] / d logits (may alias logits).  Infeasible
 *   alignments give +inf cost and zero gradient.
 * ---------------------------------------------------------------------------------------------- */
size_t tfasr_ctc_workspace_bytes(int32_t B, int32_t T, int32_t L, int32_t V);
int tfasr_ctc_loss_f32(const float* logits, const int32_t* labels, const int32_t* logit_len, const int32_t* label_len,
                       int32_t B, int32_t T, int32_t L, int32_t V, int32_t blank, float* costs, float* grads,
                       const float* grad_scale, void* workspace, size_t workspace_bytes, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * bf16 tensor-core GEMM (tcgen05 / TMEM / TMA) with fused epilogues:  C[M,N] = epilogue(alpha * A . B^T)
 * Replaces the cuBLAS calls behind keras Dense / EinsumDense / 1x1 Conv1D in the reference
 * (conformer.py:72-87,293-352,591; multihead_attention.py:628-654; base_transducer.py:285-292) and their backward.
 *   default ("TN"): A [M,K] bf16 row-major (lda), B [N,K] bf16 row-major (ldb) — both K-major
 *   a_mn_major = b_mn_major = 1: A source is [K,M] row-major, B source is [K,N] row-major (C = A^T . B; used for weight
 *   gradients dW = X^T . dY without materialising transposes)
 * All leading dimensions in elements; bf16 pointers 16-byte aligned, ld * 2 bytes a multiple of 16.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* bias;      /* [N] f32, added after alpha * acc; NULL = none */
  void* aux;              /* optional bf16 [M, ld_aux]: copy of (alpha * acc + bias) before activation / GLU (kept for backward) */
  const void* residual;   /* optional bf16 [M, ld_residual]: out = residual + residual_scale * value */
  const void* mul;        /* optional bf16 [M, ld_mul]: value *= g(mul), g per mul_mode (backward epilogues) */
  float* colsum;          /* optional f32 [N]: += column sums of the final value (bias gradients); must be zeroed by the caller */
  int32_t ld_aux, ld_residual, ld_mul;
  int32_t c_f32;          /* 0: C is bf16, 1: C is f32 */
  int32_t activation;     /* 0 none, 1 swish, 2 relu, 3 tanh */
  int32_t glu;            /* 1: columns are interleaved (a_j, b_j) pairs; C [M, N/2] = a * sigmoid(b)  (activations/glu.py:25-28) */
  int32_t mul_mode;       /* 0 none, 1: d swish(mul)/d mul, 2: 1 - mul^2 (tanh'), 3: mul */
  float residual_scale;   /* e.g. 0.5 for the macaron FFN (models/layers/residual.py:58-62) */
  float alpha;            /* accumulator scale, normally 1 */
  float dropout_p;        /* 0 = off; keep mask = hash(*dropout_seed_ptr + dropout_salt, row * N + col) >= p, kept values * 1/(1-p) */
  int32_t atomic_out;     /* 1: atomicAdd into the f32 C (split-K partial sums, gradient accumulation) */
  int32_t k_splits;       /* > 1: split the contraction over gridDim.z CTAs (requires atomic_out) */
  const unsigned long long* dropout_seed_ptr; /* DEVICE counter (NULL = 0): bumped once per step, so replayed CUDA graphs differ */
  unsigned long long dropout_salt;            /* per-call-site constant */
  float* rowstats; /* optional f32 [M, P, 2], P = 2 * ceil(N / 256): per row and column slice s the pair (max_s, sum exp(v - max_s)) of the
                      final values, so that logsumexp(row) = M + log sum_s exp(max_s - M) sum_s' with M = max_s max_s — the log-softmax
                      normaliser of the joint logits without a second pass over them (tfasr_rnnt_loss_fwd_rowstats_f32).  Empty slices
                      hold (-inf, 0).  Lean epilogue only (bias / alpha; no activation, residual, dropout, mul, colsum, split-K). */
} tfasr_gemm_epilogue;

int tfasr_gemm_bf16(const void* A, const void* B, void* C, int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb, int32_t ldc,
                    int32_t a_mn_major, int32_t b_mn_major, const tfasr_gemm_epilogue* epilogue /* NULL = plain bf16 store */,
                    tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise fused kernels of the Conformer block (bf16 activations, fp32 statistics).
 *   layernorm: keras LayerNormalization(eps) over the last axis (conformer.py:63,152,284,424-499); x is bf16 or f32, y bf16.
 *     bwd: dx = LN'(dy) (+ dres, the gradient arriving through the residual branch: models/layers/residual.py:58-62),
 *          dgamma / dbeta accumulate (+=) into f32 [D] buffers (may be NULL).
 *   scale_mask_colsum: out = scale * dropout_mask * in (+ add); colsum (+=) column sums of out — the module-entry step of
 *     backward (dropout backward + bias gradient), mask regenerated from hash(*seed_ptr + salt, row * N + col).
 *   glu_bwd: backward of a * sigmoid(b) on interleaved (a, b) column pairs (activations/glu.py:25-28).
 * ---------------------------------------------------------------------------------------------- */
int tfasr_layernorm_fwd(const void* x, int32_t x_is_f32, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                        int32_t M, int32_t D, float eps, tfasr_stream_t stream);
int tfasr_layernorm_bwd(const void* dy, const void* x, int32_t x_is_f32, const float* mean, const float* rstd, const float* gamma,
                        const void* dres, void* dx, float* dgamma, float* dbeta, int32_t M, int32_t D, tfasr_stream_t stream);
int tfasr_scale_mask_colsum(const void* in, const void* add, void* out, float* colsum, int32_t M, int32_t N, float scale, float dropout_p,
                            const unsigned long long* seed_ptr, unsigned long long salt, tfasr_stream_t stream);
int tfasr_glu_bwd(const void* aux, const void* dg, void* daux, float* colsum, int32_t M, int32_t D, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Macaron feed-forward module as one kernel (csrc/ffn_fused.cu; conformer.py:101-109 FFModule.call):
 *   out = x + factor * Drop2(W2 . Drop1(swish(W1 . LN(x) + b1)) + b2)
 * x [M, D] bf16; w1t = W1^T [H, D] bf16; w2t = W2^T [D, H] bf16; biases / LayerNorm parameters f32.  Saved for the backward (all written):
 * xn = LN(x) [M, D] bf16, z = pre-activation [M, H] bf16, h = Drop1(swish(z)) [M, H] bf16, mean / rstd [M] f32.  The dropout masks are the
 * counter hashes of the GEMM epilogues: hash(*seed_ptr + salt, row, col), salt1 for the hidden activation, salt2 for the output.
 * Supported: tfasr_ffn_supported(D, H) (D = 144, H a multiple of 64); anything else returns TFASR_INVALID_VALUE.
 * Backward gradient chain in one launch (the weight / bias gradients and the LayerNorm backward stay separate kernels):
 *   dy = factor * Drop2'(dout) [M, D];  dz = (dy . W2^T) * Drop1' * swish'(z) [M, H];  da = dz . W1^T [M, D]   (all bf16, all written)
 * w2 = W2 [H, D] bf16, w1 = W1 [D, H] bf16 (the untransposed operand copies), z the saved pre-activation.
 * ---------------------------------------------------------------------------------------------- */
int tfasr_ffn_supported(int32_t D, int32_t H);
int tfasr_ffn_bwd_bf16(const void* dout, const void* z, const void* w2, const void* w1, void* dy, void* dz, void* da, int32_t M, int32_t D, int32_t H,
                       float factor, float drop_p, const unsigned long long* seed_ptr, unsigned long long salt1, unsigned long long salt2,
                       tfasr_stream_t stream);
int tfasr_ffn_fwd_bf16(const void* x, const float* ln_g, const float* ln_b, const void* w1t, const float* b1, const void* w2t, const float* b2,
                       void* out, void* xn, void* z, void* h, float* mean, float* rstd, int32_t M, int32_t D, int32_t H, float eps, float factor,
                       float drop_p, const unsigned long long* seed_ptr, unsigned long long salt1, unsigned long long salt2,
                       tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Transducer joint, training path (base_transducer.py:199-207,291): the elementwise / reduction halves around the GEMMs.
 *   joint_tanh_fwd:   h[b,t,u,:] = tanh(e[b,t,:] + p[b,u,:])         e [B,T,J], p [B,U1,J], h [B,T,U1,J]  (all bf16)
 *   joint_reduce_bwd: de[b,t,:] = sum_u dz[b,t,u,:] (written), dp[b,u,:] += sum_t dz[b,t,u,:] (accumulated; zero it first)
 * ---------------------------------------------------------------------------------------------- */
int tfasr_joint_tanh_fwd(const void* e, const void* p, void* h, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream);
int tfasr_joint_reduce_bwd(const void* dz, float* de, float* dp, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream);
/* second generation of the two (csrc/joint_v2.cu: 16-byte accesses, column vector fixed per thread, de reduced by warp shuffles — no CTA
 * barrier in the frame loop); same contracts (J % 8 == 0, U1 <= 72 for the reduction).  Opt-in until measured on a B200. */
int tfasr_joint_tanh_fwd_v2(const void* e, const void* p, void* h, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream);
int tfasr_joint_reduce_bwd_v2(const void* dz, float* de, float* dp, int32_t B, int32_t T, int32_t U1, int32_t J, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Relative-position multi-head self-attention core (multihead_attention.py:543-582; rel_left_shift :27-64 as index arithmetic;
 * masked softmax general.py:30-41; streaming mask :104-143 when chunk > 0).
 *   qkv  [B*T, ld_qkv] bf16: q at column n*H, k at N*H + n*H, v at 2*N*H + n*H      pos [B*(2T-1), N*H] bf16 (projected relpe)
 *   ubias / vbias [N, H] f32 (content / positional attention bias)                   length [B] i32 or NULL (query-side mask)
 *   fwd:  ctx [B*T, N*H] bf16, probs [B, N, T, T] bf16 (kept for backward)
 *   bwd:  ctx = the forward's output (delta = dctx . ctx; NULL selects the FP32 kernels that sum P dP instead);
 *         dqkv [B*T, ld_qkv] bf16, dpos [B*(2T-1), N*H] bf16 (written), du / dvb [N, H] f32 (accumulated); ds_ws [B,N,T,T] bf16 scratch
 * ---------------------------------------------------------------------------------------------- */
int tfasr_rel_attention_fwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias, const int32_t* length,
                            int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx, void* probs,
                            tfasr_stream_t stream);
int tfasr_rel_attention_bwd(const void* qkv, int32_t ld_qkv, const void* pos, const float* ubias, const float* vbias, const int32_t* length,
                            int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, const void* probs, const void* dctx,
                            const void* ctx, void* ds_ws, void* dqkv, void* dpos, float* du, float* dvb, tfasr_stream_t stream);
/* Forward on tcgen05 (csrc/attention_tc.cu: TMA-fed operands, S / BD / context accumulators in TMEM, rel-shift through a per-warp SMEM
 * transpose); same contract and outputs as tfasr_rel_attention_fwd_ld, for the shapes tfasr_rel_attention_tc_supported accepts (even head
 * dimension <= 64, 16-byte aligned row pitches); TFASR_INVALID_VALUE otherwise.  Any T (keys walked in chunks of 64, queries in tiles of 128). */
int tfasr_rel_attention_tc_supported(int32_t T, int32_t N, int32_t H, int32_t ld_qkv, int32_t ld_pos);
int tfasr_rel_attention_fwd_tc(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                               const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx,
                               void* probs, tfasr_stream_t stream);
/* The same with explicit row strides (elements) for pos and dpos: every Conformer block projects the SAME relative position table
 * (conformer.py:672-701 computes relpe once and hands it to all blocks), so the position projections of all blocks can be ONE GEMM
 * [B*(2T-1), D] x [D, blocks*N*H]; block b then reads its N*H-wide column slice (pos = base + b*N*H, ld_pos = blocks*N*H) and writes its
 * slice of the shared gradient.  ld_pos % 4 == 0 and an 8-byte aligned pos keep the tensor-core path (FP32 kernels otherwise). */
int tfasr_rel_attention_fwd_ld(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                               const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history, void* ctx,
                               void* probs, tfasr_stream_t stream);
int tfasr_rel_attention_bwd_ld(const void* qkv, int32_t ld_qkv, const void* pos, int32_t ld_pos, const float* ubias, const float* vbias,
                               const int32_t* length, int32_t B, int32_t T, int32_t N, int32_t H, int32_t chunk, int32_t history,
                               const void* probs, const void* dctx, const void* ctx, void* ds_ws, void* dqkv, void* dpos, int32_t ld_dpos,
                               float* du, float* dvb, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2dSubsampling pieces and BatchNorm (+ swish), channels-last bf16 (subsampling.py:163-230, convolution.py:25-37,132-144;
 * keras BatchNormalization eps 1e-3 with batch statistics over all leading axes, as in conformer.py:326-333 / subsampling.py:199-206).
 *   conv1_3x3s2_fwd / wgrad : first conv (Cin = 1): x [B,T,F] f32, w [3,3,1,C] f32 -> y [B,ceil(T/2),ceil(F/2),C] bf16; causal pad
 *   im2col_3x3s2 / col2im_3x3s2 : a [B,T1,F1,C] <-> col [B*T2*F2, 9C] (column order kh,kw,c) feeding the tcgen05 GEMM of conv 2
 *   bn_stats        : sums[0][c] += sum x, sums[1][c] += sum x^2 over the R rows of x [R,C]
 *   bn_apply        : y = act(x * scale[c] + shift[c]), act 0 none / 1 swish
 *   bn_bwd_stats    : d = dy * act'(.), sums[0][c] += sum d, sums[1][c] += sum d * xhat
 *   bn_bwd_apply    : dx = scale * (d - sums[0]/count - xhat * sums[1]/count)   (count <= 0: fixed statistics, dx = scale * d)
 *   count == TFASR_BN_COUNT_ON_DEVICE (bn_finalize, bn_bwd_apply): the row count is the float stored at sums[2 * C] — synchronized
 *   BatchNorm over replicas whose row counts differ: every replica stores its own R there before the all-reduce of the sums, so the reduced
 *   vector carries the global count (keras BatchNormalization(synchronized=True) reduces the count the same way).
 * ---------------------------------------------------------------------------------------------- */
#define TFASR_BN_COUNT_ON_DEVICE (-2.0f)
int tfasr_conv1_3x3s2_fwd(const float* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t F, int32_t C,
                          tfasr_stream_t stream);
int tfasr_conv1_3x3s2_wgrad(const float* x, const void* dy, float* dw, int32_t B, int32_t T, int32_t F, int32_t C, tfasr_stream_t stream);
/* second generation of the two (csrc/conv1.cu: x rows staged in SMEM, thread = 8 channels, 16-byte accesses; C % 8 == 0, C <= 256); same
 * contract, wgrad accumulates (+=) into dw.  Opt-in until measured on a B200. */
int tfasr_conv1_3x3s2_fwd_v2(const float* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t F, int32_t C,
                             tfasr_stream_t stream);
int tfasr_conv1_3x3s2_wgrad_v2(const float* x, const void* dy, float* dw, int32_t B, int32_t T, int32_t F, int32_t C, tfasr_stream_t stream);
/* second generation of im2col_3x3s2 / col2im_3x3s2 (csrc/subsample_v2.cu: one CTA per output / input row, 32-bit indexing, 16-byte accesses);
 * same contracts.  Opt-in until measured on a B200. */
int tfasr_im2col_3x3s2_v2(const void* a, void* col, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream);
int tfasr_col2im_3x3s2_v2(const void* dcol, void* da, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream);
int tfasr_im2col_3x3s2(const void* a, void* col, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream);
int tfasr_col2im_3x3s2(const void* dcol, void* da, int32_t B, int32_t T1, int32_t F1, int32_t C, tfasr_stream_t stream);
int tfasr_bn_stats(const void* x, float* sums, int64_t R, int32_t C, tfasr_stream_t stream);
int tfasr_bn_apply(const void* x, const float* scale, const float* shift, void* y, int64_t R, int32_t C, int32_t act, tfasr_stream_t stream);
int tfasr_bn_bwd_stats(const void* dy, const void* x, const float* scale, const float* shift, const float* mean, const float* rstd,
                       float* sums, int64_t R, int32_t C, int32_t act, tfasr_stream_t stream);
/* per-channel bookkeeping in one launch: out [4][C] = scale, shift, mean, rstd from sums [2][C] (sum x, sum x^2) over `count` elements;
 * moving statistics updated in place (may be NULL); bwd: dgamma += sums[1], dbeta += sums[0]. */
int tfasr_bn_finalize(const float* sums, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* out, int32_t C,
                      float count, float eps, float momentum, tfasr_stream_t stream);
int tfasr_bn_bwd_finalize(const float* sums, float* dgamma, float* dbeta, int32_t C, tfasr_stream_t stream);
int tfasr_bn_bwd_apply(const void* dy, const void* x, const float* scale, const float* shift, const float* mean, const float* rstd,
                       const float* sums, float count, void* dx, int64_t R, int32_t C, int32_t act, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Recurrent part of keras.layers.LSTM in TransducerPrediction (base_transducer.py:74-88, 113-127; gate order i, f, c, o): the time
 * loop of TF's cuDNN RNN as ONE persistent kernel per direction (an 8-CTA cluster per 32 utterances keeps W_hh in SMEM).
 *   fwd: gx [B, U, 4R] f32 = x . kernel + bias (a GEMM before this call), whh [R, 4R] f32 = recurrent_kernel, h0 / c0 [B, R] f32 or
 *        NULL (zeros), length [B] or NULL (outputs at t >= length are zeros, state keeps evolving: zero_output_for_mask)
 *        -> y [B, U, R] bf16, saved_g [B, U, 4R] bf16 (i, f, g, o interleaved per unit pair), saved_c [B, U, R] f32, hT / cT [B, R] f32 or NULL
 *   bwd: dy [B, U, R] bf16 -> dgates [B, U, 4R] bf16 (pre-activation gradients; dW_ih, dW_hh, dbias, dx follow as plain GEMMs /
 *        column sums), dh0 / dc0 [B, R] f32 or NULL
 * tfasr_lstm_supported(R): 1 when R % 64 == 0 and the weight slices fit in SMEM (R <= 384); else the calls return 1 (invalid value).
 * ---------------------------------------------------------------------------------------------- */
int tfasr_lstm_supported(int32_t R);
int tfasr_lstm_rec_fwd(const float* gx, const float* whh, const float* h0, const float* c0, const int32_t* length, void* y, void* saved_g,
                       float* saved_c, float* hT, float* cT, int32_t B, int32_t U, int32_t R, tfasr_stream_t stream);
int tfasr_lstm_rec_bwd(const void* dy, const void* saved_g, const float* saved_c, const float* c0, const float* whh, const int32_t* length,
                       void* dgates, float* dh0, float* dc0, int32_t B, int32_t U, int32_t R, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Parameter-side kernels: the whole optimizer update and all bf16 weight-operand copies are one launch each.
 *   tfasr_adamw_flat_f32 : AdamW (Keras Adam + decoupled weight decay, small.yml.j2:73-87; per-variable `optimizer.apply`,
 *                          base_model.py:196-215) over flat fp32 buffers p, g, m, v [n] (n % 4 == 0, 16-byte aligned);
 *                          g' = g * grad_scale + 2 * l2 * p where l2_mask[i] != 0 (Keras l2 regularizers, conformer.py:27);
 *                          lr / step are device scalars (step = 1, 2, ...); zero_grad != 0 clears g for the next step.
 *   tfasr_pack_weights   : table (device memory) of n_entries tfasr_pack_entry; entry = fp32 [K, N] source scattered to
 *                          dst_w[k, c0 + n * cs] (bf16, or fp32 when flags & 1) and dst_wt[c0 + n * cs, k] (bf16); NULL = skip.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tfasr_pack_entry {
  const float* src;
  void* dst_w;
  void* dst_wt;
  int32_t K, N, src_ld, ldw, ldwt, c0, cs, flags;
} tfasr_pack_entry;
int tfasr_adamw_flat_f32(float* p, float* g, float* m, float* v, const uint8_t* l2_mask, int64_t n, const float* lr, const float* step,
                         float beta1, float beta2, float eps, float weight_decay, float l2, float grad_scale, int32_t zero_grad,
                         tfasr_stream_t stream);
int tfasr_pack_weights(const void* table, int32_t n_entries, tfasr_stream_t stream);
/* inverse for gradients: entry.src = parameter gradient (fp32 [K, N], row stride src_ld), entry.dst_w = packed fp32 buffer:
 * grad[k, n] += packed[k, c0 + n * cs]; packed[...] = 0.  One launch per step for all fused Q|K|V / GLU projections. */
int tfasr_unpack_add(const void* table, int32_t n_entries, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Causal depthwise 1-D convolution of the ConvModule (convolution.py:159-216, conformer.py:311-323,370): channels-last bf16 [B,T,C],
 * w [K,C] f32 (keras depthwise kernel [K,C,1]), bias [C] f32 or NULL.  bwd writes dx and accumulates (+=) dw [K,C] / dbias [C] (or NULL).
 * ---------------------------------------------------------------------------------------------- */
int tfasr_dwconv_causal_fwd(const void* x, const float* w, const float* bias, void* y, int32_t B, int32_t T, int32_t C, int32_t K,
                            tfasr_stream_t stream);
int tfasr_dwconv_causal_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int32_t B, int32_t T, int32_t C,
                            int32_t K, tfasr_stream_t stream);
/* the two halves separately: bwd_dx = the data gradient of tfasr_dwconv_causal_bwd alone; wgrad_v2 (csrc/dwconv_wgrad_v2.cu) = the weight /
 * bias gradient with one thread per (tap, channel pair) walking the frames from L1/L2 (no SMEM staging, no barriers).  Opt-in pair until
 * measured on a B200. */
int tfasr_dwconv_causal_bwd_dx(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t K, tfasr_stream_t stream);
int tfasr_dwconv_causal_wgrad_v2(const void* dy, const void* x, float* dw, float* dbias, int32_t B, int32_t T, int32_t C, int32_t K,
                                 tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SpecAugment on the features, in place (augmentations/augmentation.py:67-90, methods/specaugment.py:58-137): n_freq frequency masks
 * (width < freq_factor) then n_time time masks (width < floor(len * p_upperbound)) per utterance, each applied with its probability,
 * masked cells set to mask_value.  Random stream = counter hash of (*seed_ptr + salt, utterance, mask).
 * ---------------------------------------------------------------------------------------------- */
int tfasr_specaugment_f32(float* feat /*[B,T,F]*/, const int32_t* length /*[B] frames, or NULL*/, int32_t B, int32_t T, int32_t F, int32_t n_time,
                          float p_upperbound, float time_prob, int32_t n_freq, int32_t freq_factor, float freq_prob, float mask_value,
                          const unsigned long long* seed_ptr, unsigned long long salt, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ContextNet encoder body (tensorflow_asr/models/encoders/contextnet.py).  Activations channels-last bf16 [B,T,C]
 * (C % 8 == 0, 16-byte aligned), parameters / statistics fp32.  act codes: 0 linear, 1 swish (silu), 2 relu (contextnet.py:29-36).
 *   tfasr_seq_dwconv_fwd/bwd : depthwise half of SeparableConv1D(padding="causal", strides=s) (contextnet.py:53-63, causal pad
 *                         convolution.py:25-37): y [B, ceil(T/s), C] = sum_k w[k,c] x[b, j s - (K-1) + k, c]; w [K,C] f32 (keras
 *                         depthwise_kernel [K,C,1]).  bwd writes dx [B,T,C] (NULL = skip) and accumulates (+=) dw [K,C]; K <= 32.
 *                         (the pointwise half + bias is tfasr_gemm_bf16, the BatchNorm + activation tfasr_bn_*)
 *   tfasr_seq_masked_mean  : GlobalAveragePooling1D with the sequence mask (contextnet.py:137,156): m [B,C] f32 = mean over t < length[b]
 *   tfasr_se_fc_fwd/bwd    : SEModule fc1 -> activation -> fc2 -> sigmoid (contextnet.py:157-160).  W1 [C,H], b1 [H], W2 [H,C], b2 [C];
 *                         fwd -> z1 [B,H] (pre-activation, saved), s [B,C].  bwd: ds [B,C] -> dm [B,C] (gradient of the masked mean's
 *                         INPUT frames, i.e. already divided by length), accumulates (+=) dW1 / db1 / dW2 / db2; dz2 [B,C], dz1 [B,H] scratch.
 *   tfasr_se_apply_fwd     : out = act(y * s[b,:] + res)   (SEModule.call :162-163 + ConvBlock.call :258-261; res NULL = no residual)
 *   tfasr_se_apply_bwd     : dres [B,T,C] bf16 = dout * act'(y s + res) (also the residual branch's gradient); ds [B,C] = sum_t dres * y
 *   tfasr_se_dy            : dy = dres * s[b,:] + (t < length[b] ? dm[b,:] : 0)
 * ---------------------------------------------------------------------------------------------- */
int tfasr_seq_dwconv_fwd(const void* x, const float* w, void* y, int32_t B, int32_t T, int32_t C, int32_t K, int32_t stride, tfasr_stream_t stream);
int tfasr_seq_dwconv_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, int32_t B, int32_t T, int32_t C, int32_t K,
                         int32_t stride, tfasr_stream_t stream);
int tfasr_seq_masked_mean(const void* y, const int32_t* length, float* m, int32_t B, int32_t T, int32_t C, tfasr_stream_t stream);
int tfasr_se_fc_fwd(const float* m, const float* W1, const float* b1, const float* W2, const float* b2, float* z1, float* s, int32_t B, int32_t C,
                    int32_t H, int32_t act, tfasr_stream_t stream);
int tfasr_se_fc_bwd(const float* ds, const float* s, const float* z1, const float* m, const float* W1, const float* W2, const int32_t* length,
                    float* dz2, float* dz1, float* dm, float* dW1, float* db1, float* dW2, float* db2, int32_t B, int32_t T, int32_t C, int32_t H,
                    int32_t act, tfasr_stream_t stream);
int tfasr_se_apply_fwd(const void* y, const float* s, const void* res, void* out, int32_t B, int32_t T, int32_t C, int32_t act,
                       tfasr_stream_t stream);
int tfasr_se_apply_bwd(const void* dout, const void* y, const float* s, const void* res, void* dres, float* ds, int32_t B, int32_t T, int32_t C,
                       int32_t act, tfasr_stream_t stream);
int tfasr_se_dy(const void* dres, const float* s, const float* dm, const int32_t* length, void* dy, int32_t B, int32_t T, int32_t C,
                tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2D lowering for the DeepSpeech2 conv module (models/encoders/deepspeech2.py:56-60,105-117): general kernel
 * (kh, kw), strides (sh, sw), zero padding of pt rows / pl columns in front ("same": total // 2, causal: k - 1).  The convolution is
 * col [B*To*Fo, Kp] . W[Kp, Cout] on tfasr_gemm_bf16 with Kp = round_up(kh*kw*Cin, 8) (zero columns beyond kh*kw*Cin).
 *   tfasr_conv_im2col : x [B,T,F,Cin] (fp32 when x_f32 != 0 — the feature tensor —, else bf16) -> col bf16, column (i*kw + j)*Cin + ci
 *   tfasr_conv_col2im : dcol -> da [B,T,F,Cin] bf16 (Cin % 8 == 0), the data gradient
 * ---------------------------------------------------------------------------------------------- */
typedef struct tfasr_conv_geom {
  int32_t B, T, F, Cin, To, Fo, kh, kw, sh, sw, pt, pl, Kp;
} tfasr_conv_geom;
int tfasr_conv_im2col(const void* x, int32_t x_f32, void* col, const tfasr_conv_geom* geom, tfasr_stream_t stream);
int tfasr_conv_col2im(const void* dcol, void* da, const tfasr_conv_geom* geom, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Persistent greedy transducer decode: the `tf.while_loop` of Transducer.recognize_batch / recognize_single
 * (models/transducer/base_transducer.py:496-575, :577-712; one call_next :437-464 per iteration) as ONE launch, one CTA per utterance.
 * All buffers fp32 / int32 device memory.  Weights in their Keras layouts: emb [V,E], w_ih [E,4R], w_hh [R,4R], b [4R] (gate order
 * i,f,c,o), ln_g / ln_b [R] or both NULL, wp [R,J], bp [J], wv [J,V], bv [V]; e_proj [B,T,J] = encoder output . joint.enc kernel + bias.
 * State (in/out): prev_tok [B], states [B,2,R] (h, c), frame_idx [B], tok_idx [B], tokens [B,max_tokens].
 *   mode 0  recognize_batch trajectory of each utterance until ITS frame index reaches nframes - 1 (-> iters[b] = iterations run, or -1
 *           when max_iters was hit first; cap_iter[b] = first iteration whose loop test saw tok_idx >= max_tokens - 1, or -1)
 *   mode 1  the same trajectory for exactly iters[b] further iterations (the loop stops when ALL utterances are done: the host takes
 *           L = max_b iters[b] after mode 0 and runs mode 1 with L - iters[b])
 *   mode 2  recognize_single: a frame is left on blank or after per_frame_cap symbols; tok_idx = index of the last written token (-1)
 * `threads`: CTA size, a multiple of 32 in [32, 1024].
 * ---------------------------------------------------------------------------------------------- */
typedef struct tfasr_greedy_params {
  const float* e_proj;
  const int32_t* nframes;
  const float *emb, *w_ih, *w_hh, *b, *ln_g, *ln_b, *wp, *bp, *wv, *bv;
  int32_t* prev_tok;
  float* states;
  int32_t *frame_idx, *tok_idx, *tokens, *iters, *cap_iter;
  int32_t T, E, R, J, V, max_tokens, blank, mode, max_iters, per_frame_cap;
} tfasr_greedy_params;
int tfasr_greedy_decode_f32(const tfasr_greedy_params* params, int32_t B, int32_t threads, tfasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small all-reduce over peer memory: the sync-BatchNorm statistics exchange of keras BatchNormalization(synchronized=True)
 * (models/encoders/conformer.py:326-333, models/layers/subsampling.py norms "batch"), 2 x C floats per BatchNorm, forward and backward —
 * in place of one NCCL all-reduce per call.  One launch, one CTA: push into every rank's window, flag (release.sys), bounded wait
 * (acquire.sys), sum in rank order (bit-identical on every rank).
 *   windows   device array [world] of window base addresses AS MAPPED IN THIS PROCESS (cudaIpcOpenMemHandle / symmetric memory); a window
 *             is tfasr_peer_window_bytes(world, capacity) bytes, zero-initialised before the first call on any rank
 *   state     2 uint32 of LOCAL device memory, zero-initialised once: [0] calls completed (advanced by the kernel: graph-replay safe),
 *             [1] becomes 1 when a peer did not arrive within timeout_ms (the call then returns garbage instead of hanging the GPU)
 *   data      [n] fp32 in/out, n <= capacity; every rank makes the same sequence of calls on one stream
 * ---------------------------------------------------------------------------------------------- */
typedef struct tfasr_peer_comm {
  const uint64_t* windows;
  uint32_t* state;
  int32_t rank, world, capacity;
  int32_t timeout_ms; /* 0 = default (30 s) */
} tfasr_peer_comm;
size_t tfasr_peer_window_bytes(int32_t world, int32_t capacity);
/* Windows over plain CUDA IPC, on the calling thread's current device: alloc = cudaMalloc + zero-fill + cudaIpcGetMemHandle (the 64-byte
 * handle goes to the peers through the host-side process group), open = cudaIpcOpenMemHandle with lazy peer access, close = cudaFree
 * (owned) / cudaIpcCloseMemHandle. */
#define TFASR_PEER_IPC_HANDLE_BYTES 64
int tfasr_peer_window_alloc(size_t bytes, void** window, void* ipc_handle);
int tfasr_peer_window_open(const void* ipc_handle, void** window);
int tfasr_peer_window_close(void* window, int32_t owned);
int tfasr_peer_allreduce_f32(const tfasr_peer_comm* comm, float* data, int32_t n, tfasr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TFASR_B200_H_ */
