"""Augmentation — host-side mirror of tensorflow_asr/augmentations/augmentation.py for the feature-level SpecAugment
(`feature_augment: {time_masking: {...}, freq_masking: {...}}`, e.g. conformer small.yml.j2:13-24), run by ONE in-place CUDA kernel
(`tfasr_specaugment_f32`), and for the signal-level `gauss_noise`."""
from __future__ import annotations

import torch

from . import _lib

_SALT = 0x5EC0A06


def signal_augment(config, signals, signals_length, generator=None):
    """`signal_augment: {gauss_noise: {mean, stddev, prob}}` (augmentations/methods/gaussnoise.py:19-40, applied per utterance by
    Augmentation.signal_augment's map_fn, augmentation.py:43-65): with probability `prob` per utterance add N(mean, stddev) noise to the
    samples inside the utterance's length.  Device-side random draws (torch's generator: the stream differs from TF's, like SpecAugment's)."""
    sa = config.get("signal_augment") or {}
    unknown = set(sa) - {"gauss_noise"}
    if unknown:
        raise KeyError(f"No augmentation named: {sorted(unknown)}; available: ['gauss_noise'] (signal), ['freq_masking', 'time_masking'] (feature)")
    if "gauss_noise" not in sa:
        return signals, signals_length
    gn = sa["gauss_noise"] or {}
    mean, stddev, prob = float(gn.get("mean", 0.0)), float(gn.get("stddev", 0.075)), float(gn.get("prob", 0.5))
    B, N = signals.shape
    dev = signals.device
    apply = (torch.rand(B, device=dev, generator=generator) <= prob).to(signals.dtype)[:, None]
    noise = torch.randn(B, N, device=dev, dtype=signals.dtype, generator=generator) * stddev + mean
    inside = (torch.arange(N, device=dev)[None, :] < signals_length.to(dev)[:, None]).to(signals.dtype)
    return signals + noise * inside * apply, signals_length


def feature_augment(config, features, features_length, seed=None):
    """features [B,T,F,1] f32 (modified in place), features_length [B] -> same tensors."""
    fa = config.get("feature_augment") or {}
    tm, fm = fa.get("time_masking"), fa.get("freq_masking")
    if not tm and not fm:
        return features, features_length
    for m in (tm, fm):
        if m and not isinstance(m.get("mask_value", 0), (int, float)) and m.get("mask_value") != "zero":
            raise NotImplementedError("SpecAugment mask_value mean/min/max is outside the B200 hot path (zero or a number)")
    if features.dtype != torch.float32:
        raise NotImplementedError("feature_augment expects f32 features")
    from . import fused

    lib = _lib.load()
    B, T, F = features.shape[0], features.shape[1], features.shape[2]
    seed_t = seed if seed is not None else fused.STATE.seed_tensor(features.device)

    def num(v, default):
        return default if v is None else v

    mv = 0.0
    for m in (tm, fm):
        if m and isinstance(m.get("mask_value", 0), (int, float)):
            mv = float(m.get("mask_value", 0))
    _lib.check(
        lib.tfasr_specaugment_f32(features.data_ptr(), _lib.ptr(features_length.to(torch.int32).contiguous()), B, T, F,
                                  int(num(tm and tm.get("num_masks"), 1)) if tm else 0, float(num(tm and tm.get("p_upperbound"), 1.0)) if tm else 0.0,
                                  float(num(tm and tm.get("prob"), 1.0)) if tm else 0.0,
                                  int(num(fm and fm.get("num_masks"), 1)) if fm else 0, int(num(fm and fm.get("mask_factor"), 27)) if fm else 0,
                                  float(num(fm and fm.get("prob"), 1.0)) if fm else 0.0, mv, seed_t.data_ptr(), _SALT,
                                  _lib.stream_ptr(features.device)),
        "tfasr_specaugment_f32",
    )
    return features, features_length
