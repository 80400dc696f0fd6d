"""FeatureExtraction / SpeechFeaturizer — host-side mirror of the reference's log-mel frontend layer.

Mirrors `tensorflow_asr/models/layers/feature_extraction.py:33-331` (same constructor kwargs, `call` contract,
`get_nframes`, `compute_mask`, `get_signal_chunk_size_and_step`, `compute_output_shape`), with torch CUDA tensors in
place of tf.Tensors.  All arithmetic runs in ONE fused sm_100a kernel behind the C ABI (`tfasr_logmel_f32`,
include/tfasr_b200.h); there is no CPU path.

`SpeechFeaturizer` (the name BASELINE.json uses; it no longer exists at the reference commit) is an alias.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib

FEATURE_TYPES = ("spectrogram", "log_mel_spectrogram", "mfcc", "log_gammatone_spectrogram")
_SUPPORTED = {"log_mel_spectrogram": 0, "spectrogram": 1}


class FeatureExtraction(torch.nn.Module):
    def __init__(
        self,
        sample_rate=16000,
        frame_ms=25,
        stride_ms=10,
        num_feature_bins=80,
        feature_type="log_mel_spectrogram",
        preemphasis=0.97,
        pad_end=True,
        use_librosa_like_stft=False,
        epsilon=1e-6,
        lower_edge_hertz=0.0,
        upper_edge_hertz=8000.0,
        log_base="e",
        nfft=512,
        normalize_signal=False,
        normalize_zscore=False,
        normalize_min_max=False,
        padding=0,
        augmentation_config={},
        **kwargs,
    ):
        super().__init__()
        assert feature_type in FEATURE_TYPES, f"feature_type must be in {FEATURE_TYPES}"
        if feature_type not in _SUPPORTED:
            # out of the hot-path scope (BASELINE configs use log-mel / spectrogram only)
            raise NotImplementedError(f"feature_type={feature_type!r} is outside the B200 hot path (log_mel_spectrogram, spectrogram)")
        if use_librosa_like_stft:
            raise NotImplementedError("use_librosa_like_stft=True is outside the B200 hot path")
        self.name = kwargs.get("name", feature_type)
        self.sample_rate = sample_rate
        self.frame_ms = frame_ms
        self.frame_length = int(round(self.sample_rate * self.frame_ms / 1000.0))
        self.stride_ms = stride_ms
        self.frame_step = int(round(self.sample_rate * self.stride_ms / 1000.0))
        self.num_feature_bins = num_feature_bins
        self.feature_type = feature_type
        self.preemphasis = preemphasis
        self.pad_end = pad_end
        self.use_librosa_like_stft = use_librosa_like_stft
        self.epsilon = epsilon
        assert self.epsilon > 1e-9 and self.epsilon <= 0.001, "epsilon must be in (1e-9, 0.001]"
        self.lower_edge_hertz = lower_edge_hertz
        self.upper_edge_hertz = upper_edge_hertz
        self.log_base = log_base
        assert self.log_base in ("10", "e"), "log_base must be '10' or 'e'"
        self._normalize_signal = normalize_signal
        self._normalize_zscore = normalize_zscore
        self._normalize_min_max = normalize_min_max
        self.padding = padding
        self.nfft = self.frame_length if nfft is None else nfft
        self.augmentation_config = dict(augmentation_config or {})
        self._tables = {}  # (cfg key, device) -> device table

    # ------------------------------------------------------------------ C-ABI plumbing
    def _cfg(self) -> _lib.LogmelCfg:
        return _lib.LogmelCfg(
            sample_rate=int(self.sample_rate), frame_length=int(self.frame_length), frame_step=int(self.frame_step),
            nfft=int(self.nfft), num_feature_bins=int(self.num_feature_bins), feature_type=_SUPPORTED[self.feature_type],
            pad_end=int(bool(self.pad_end)), log_base10=int(self.log_base == "10"), padding=int(self.padding),
            preemphasis=float(self.preemphasis or 0.0), epsilon=float(self.epsilon),
            lower_edge_hertz=float(self.lower_edge_hertz), upper_edge_hertz=float(self.upper_edge_hertz),
        )

    def _table(self, cfg: _lib.LogmelCfg, device) -> torch.Tensor:
        key = (bytes(cfg), str(device))
        tab = self._tables.get(key)
        if tab is None:
            lib = _lib.load()
            nbytes = lib.tfasr_logmel_table_bytes(ctypes.byref(cfg))
            if nbytes == 0:
                raise ValueError("FeatureExtraction: configuration not supported by the fused kernel "
                                 "(nfft must be 256/512/1024, frame_length <= nfft, frame_step <= frame_length)")
            host = torch.empty(nbytes // 4, dtype=torch.float32)
            _lib.check(lib.tfasr_logmel_build_table(ctypes.byref(cfg), host.data_ptr()), "tfasr_logmel_build_table")
            tab = host.to(device)
            self._tables[key] = tab
        return tab

    # ------------------------------------------------------------------ reference API
    def get_signal_chunk_size_and_step(self, nframes):
        chunk_size = (nframes - 1) * self.frame_step + self.frame_length
        chunk_step = nframes * self.frame_step
        return chunk_size, chunk_step

    def get_nframes(self, nsamples):
        if self.pad_end:
            return -(-nsamples // self.frame_step)
        return 1 + (nsamples - self.frame_length) // self.frame_step

    def forward(self, inputs, training=False):
        return self.call(inputs, training=training)

    def call(self, inputs, training=False):
        """(signals [B, N] float, signals_length [B] int) -> (features [B, T, F, 1] f32, features_length [B] i32)."""
        signals, signals_length = inputs
        _lib.require_cuda(signals, "FeatureExtraction")
        if training and self.augmentation_config:
            from . import augment  # SpecAugment lives next to the frontend (reference: augmentations/)

            signals, signals_length = augment.signal_augment(self.augmentation_config, signals, signals_length)
        lib = _lib.load()
        orig_dtype = signals.dtype
        x = signals.to(torch.float32).contiguous()  # stft casts f16/bf16 up to f32 (feature_extraction.py:193-195)
        B, N = x.shape
        cfg = self._cfg()
        table = self._table(cfg, x.device)
        stream = _lib.stream_ptr(x.device)
        if self._normalize_signal:
            if x.data_ptr() == signals.data_ptr():
                x = x.clone()
            _lib.check(lib.tfasr_signal_normalize_f32(_lib.ptr(x), B, N, float(self.epsilon), stream), "tfasr_signal_normalize_f32")
        T = lib.tfasr_logmel_nframes(ctypes.byref(cfg), N)
        feats = torch.empty((B, max(T, 0), self.num_feature_bins), dtype=torch.float32, device=x.device)
        lens_in = signals_length.to(device=x.device, dtype=torch.int32).contiguous()
        lens_out = torch.empty_like(lens_in)
        with _lib.TIMERS.region("logmel"):
            _lib.check(
                lib.tfasr_logmel_f32(_lib.ptr(x), _lib.ptr(lens_in), B, N, ctypes.byref(cfg), _lib.ptr(table), _lib.ptr(feats),
                                     _lib.ptr(lens_out), stream),
                "tfasr_logmel_f32",
            )
        if self._normalize_zscore or self._normalize_min_max:
            mode = 0 if self._normalize_zscore else 1
            # min_value = self.logarithm(self.epsilon): logarithm() adds epsilon again (feature_extraction.py:185,215)
            log_eps = math.log(2.0 * self.epsilon) / (math.log(10.0) if self.log_base == "10" else 1.0)
            _lib.check(lib.tfasr_feature_normalize_f32(_lib.ptr(feats), B, T, self.num_feature_bins, mode, float(self.epsilon),
                                                       float(log_eps), stream), "tfasr_feature_normalize_f32")
        if orig_dtype in (torch.float16, torch.bfloat16):
            feats = feats.to(orig_dtype)
        feats = feats.unsqueeze(-1)
        if training and self.augmentation_config:
            from . import augment

            feats, lens_out = augment.feature_augment(self.augmentation_config, feats, lens_out)
        return feats, lens_out

    def compute_mask(self, inputs, mask=None):
        signals, signals_length = inputs
        maxlen = signals.shape[1] + self.padding
        nsamples = torch.clamp(signals_length.to(torch.int64), max=maxlen)
        nframes = self.get_nframes(nsamples)
        padded_nframes = self.get_nframes(maxlen)
        return torch.arange(padded_nframes, device=signals.device)[None, :] < nframes[:, None], None

    def compute_output_shape(self, input_shape):
        signal_shape, signal_length_shape = input_shape
        B, nsamples = signal_shape
        if nsamples is None:
            output_shape = [B, None, self.num_feature_bins, 1]
        else:
            output_shape = [B, self.get_nframes(nsamples + self.padding), self.num_feature_bins, 1]
        return tuple(output_shape), tuple(signal_length_shape)


SpeechFeaturizer = FeatureExtraction
