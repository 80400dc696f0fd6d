"""CPU oracle: the log-mel frontend (TEST INFRASTRUCTURE, never the product path).

Restates, in numpy, ``tensorflow_asr/models/layers/feature_extraction.py``:
  * ``preemphasis_signal``            :170-175
  * ``stft``                          :192-212  (tf.signal.stft, pad_end=True, periodic Hann)
  * ``logarithm``                     :214-218
  * ``log_mel_spectrogram``           :220-231  (tf.signal.linear_to_mel_weight_matrix, HTK)
  * ``spectrogram``                   :233-235
  * ``normalize_signal``              :164-168
  * ``normalize_audio_features``      :177-190
  * ``get_nframes`` / ``compute_mask``:305-331
  * ``FeatureExtraction.call``        :255-303

The spectral arithmetic itself lives in a third-party dependency that is absent
from /root/reference: ``tensorflow~=2.18.0`` (requirements.gpu.txt:1) —
``tf.signal.hann_window`` (periodic raised cosine, n = window_length for even
lengths), ``tf.signal.frame(pad_end=True)`` (ceil(N/step) frames, zero tail),
``tf.signal.rfft`` (frame zero-padded on the right to fft_length) and
``tf.signal.linear_to_mel_weight_matrix`` (HTK mel 1127*ln(1+f/700), computed in
the requested dtype, DC row zeroed).  Their published algorithms are restated here.

PARITY PINNING: the reference's own tests do not pin any log-mel value
(tests/test_layers.py:33-78 only plots) and TensorFlow cannot be installed in the
build container => **parity unpinned against the reference**.  The restatement is
cross-checked in tests/test_oracle_logmel.py against an independent
implementation (torch.stft + torchaudio HTK mel filterbank) and against a direct
O(N^2) DFT in float64.
"""
from __future__ import annotations

import numpy as np


def hann_window_periodic(n: int, dtype=np.float32) -> np.ndarray:
    """tf.signal.hann_window(n, periodic=True): 0.5 - 0.5 cos(2 pi k / (n + even - 1))."""
    even = 1 - n % 2
    denom = n + even - 1
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / denom)).astype(dtype)


def hertz_to_mel(f, dtype):
    return (np.asarray(1127.0, dtype) * np.log(np.asarray(1.0, dtype) + np.asarray(f, dtype) / np.asarray(700.0, dtype))).astype(dtype)


def linear_to_mel_weight_matrix(num_mel_bins=80, num_spectrogram_bins=257, sample_rate=16000,
                                lower_edge_hertz=0.0, upper_edge_hertz=8000.0, dtype=np.float32) -> np.ndarray:
    """tf.signal.linear_to_mel_weight_matrix restated; all arithmetic in `dtype` like TF."""
    dt = dtype
    nyquist = np.asarray(sample_rate, dt) / np.asarray(2.0, dt)
    lin = np.linspace(np.asarray(0.0, dt), nyquist, num_spectrogram_bins, dtype=dt)[1:]
    spec_mel = hertz_to_mel(lin, dt)[:, None]
    edges = np.linspace(hertz_to_mel(lower_edge_hertz, dt), hertz_to_mel(upper_edge_hertz, dt), num_mel_bins + 2, dtype=dt)
    lower = edges[None, :-2]
    center = edges[None, 1:-1]
    upper = edges[None, 2:]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(np.asarray(0.0, dt), np.minimum(lower_slopes, upper_slopes)).astype(dt)
    return np.concatenate([np.zeros((1, num_mel_bins), dt), w], axis=0)


def get_nframes(nsamples, frame_length=400, frame_step=160, pad_end=True):
    """feature_extraction.py:305-313."""
    nsamples = np.asarray(nsamples)
    if pad_end:
        return -(-nsamples // frame_step)
    return 1 + (nsamples - frame_length) // frame_step


def preemphasis(signal: np.ndarray, coef: float) -> np.ndarray:
    """feature_extraction.py:170-175 (first sample kept; filter crosses any zero padding)."""
    if not coef or coef <= 0.0:
        return signal
    out = np.empty_like(signal)
    out[:, 0] = signal[:, 0]
    out[:, 1:] = signal[:, 1:] - np.asarray(coef, signal.dtype) * signal[:, :-1]
    return out


def stft_power(signal: np.ndarray, frame_length=400, frame_step=160, nfft=512, pad_end=True, dtype=np.float32) -> np.ndarray:
    """|rfft(frame * hann)|^2, feature_extraction.py:205-209; [B, T, nfft/2+1]."""
    B, N = signal.shape
    T = int(get_nframes(N, frame_length, frame_step, pad_end))
    if pad_end:
        need = (T - 1) * frame_step + frame_length
        if need > N:
            signal = np.concatenate([signal, np.zeros((B, need - N), signal.dtype)], axis=1)
    idx = np.arange(T)[:, None] * frame_step + np.arange(frame_length)[None, :]
    frames = signal[:, idx] * hann_window_periodic(frame_length, dtype)[None, None, :]
    spec = np.fft.rfft(frames.astype(np.float64) if dtype == np.float64 else frames, n=nfft, axis=-1)
    mag = np.abs(spec).astype(dtype)
    return np.square(mag)


def log_features(S: np.ndarray, epsilon=1e-6, log_base="e") -> np.ndarray:
    S = S + np.asarray(epsilon, S.dtype)
    if log_base == "10":
        return (np.log(S) / np.log(np.asarray(10.0, S.dtype))).astype(S.dtype)
    return np.log(S)


def feature_extraction(signals: np.ndarray, signals_length: np.ndarray, *, sample_rate=16000, frame_ms=25, stride_ms=10,
                       num_feature_bins=80, feature_type="log_mel_spectrogram", preemphasis_coef=0.97, pad_end=True,
                       epsilon=1e-6, lower_edge_hertz=0.0, upper_edge_hertz=8000.0, log_base="e", nfft=512,
                       normalize_signal=False, normalize_zscore=False, normalize_min_max=False, padding=0,
                       dtype=np.float32):
    """FeatureExtraction.call(training=False), feature_extraction.py:255-303.

    Returns (features [B, T, F, 1], features_length [B] int32).
    """
    frame_length = int(round(sample_rate * frame_ms / 1000.0))
    frame_step = int(round(sample_rate * stride_ms / 1000.0))
    x = np.asarray(signals, dtype=dtype)
    if padding > 0:
        x = np.concatenate([x, np.zeros((x.shape[0], padding), dtype)], axis=1)
    if normalize_signal:
        gain = 1.0 / (np.max(np.abs(x), axis=1, keepdims=True) + np.asarray(epsilon, dtype))
        x = x * gain.astype(dtype)
    x = preemphasis(x, preemphasis_coef)
    S = stft_power(x, frame_length, frame_step, nfft, pad_end, dtype)
    if feature_type == "spectrogram":
        feat = log_features(S, epsilon, log_base)[:, :, :num_feature_bins]
    elif feature_type == "log_mel_spectrogram":
        W = linear_to_mel_weight_matrix(num_feature_bins, S.shape[-1], sample_rate, lower_edge_hertz, upper_edge_hertz, dtype)
        feat = log_features(np.matmul(S, W).astype(dtype), epsilon, log_base)
    else:
        raise NotImplementedError(feature_type)
    if normalize_zscore:
        mean = feat.mean(axis=1, keepdims=True)
        std = np.sqrt(feat.var(axis=1, keepdims=True) + np.asarray(epsilon, dtype))
        feat = (feat - mean) / std
    elif normalize_min_max:
        mn = log_features(np.asarray(epsilon, dtype), epsilon, log_base)  # logarithm(eps) adds eps again (:185)
        feat = (feat - mn) / (feat.max(axis=1, keepdims=True) - mn)
    feat = feat[..., None].astype(dtype)
    out_len = get_nframes(np.asarray(signals_length, np.int32), frame_length, frame_step, pad_end).astype(np.int32)
    return feat, out_len


def compute_mask(signals_shape, signals_length, padding=0, frame_length=400, frame_step=160, pad_end=True):
    """feature_extraction.py:315-322."""
    n = np.minimum(np.asarray(signals_length), signals_shape[1] + padding)
    nframes = get_nframes(n, frame_length, frame_step, pad_end)
    padded = int(get_nframes(signals_shape[1] + padding, frame_length, frame_step, pad_end))
    return np.arange(padded)[None, :] < nframes[:, None]


def get_signal_chunk_size_and_step(nframes, frame_length=400, frame_step=160):
    """feature_extraction.py:141-162."""
    return (nframes - 1) * frame_step + frame_length, nframes * frame_step
