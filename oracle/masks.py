"""CPU oracle: attention masks and the Transformer-XL relative shift (TEST INFRASTRUCTURE).

Restates ``tensorflow_asr/models/layers/multihead_attention.py``:
  * ``rel_left_shift``          :27-64   literal pad / reshape / slice sequence (both branches)
  * ``compute_causal_mask``     :79-101
  * ``compute_streaming_mask``  :104-143
and the masked softmax of ``models/layers/general.py:30-41`` with the fill value of
``utils/math_util.py:240-246`` (-1e9, or float16.min for fp16).

PARITY PINNING: ``compute_streaming_mask`` is pinned bit-exactly by the reference's own golden
matrices (tests/test_mask.py:6-55, committed as tests/golden/streaming_mask.json by
tests/golden/make_golden.py); ``rel_left_shift`` is pinned by the causal example in its docstring
(multihead_attention.py:31-41) on the lower triangle — the code as written leaves wrap-around values in the
upper triangle (row 0 is [3, 0, 4], not the docstring's [3, 0, 0]); those positions are masked downstream.
"""
from __future__ import annotations

import numpy as np


def rel_left_shift(x: np.ndarray, causal: bool = False) -> np.ndarray:
    """x [B,N,T,R] -> shifted [B,N,T,R'] exactly as the reference's pad/reshape/slice chain."""
    b, n, t, r = x.shape
    if causal:
        x = np.pad(x, [(0, 0), (0, 0), (0, 0), (1, 0)])
        x = x.reshape(b, n, -1)
        x = np.pad(x, [(0, 0), (0, 0), (r - t, 0)])
        x = x.reshape(b, n, 1 + t, r)
        return x[:, :, 1:, :]
    x = np.pad(x, [(0, 0), (0, 0), (0, 0), (0, 1)])
    x = x.reshape(b, n, -1)
    x = np.pad(x, [(0, 0), (0, 0), (0, r - t)])
    x = x.reshape(b, n, 1 + t, r)
    return x[:, :, :t, (t - 1):]


def rel_shift_closed_form(x: np.ndarray, s: int) -> np.ndarray:
    """Non-causal closed form used by the fused kernel: after the shift and the 'last S columns'
    slice (multihead_attention.py:563-567), out[t, j] = x[t, (R - S - T + 1) + T - 1 - t + j]
    for R = 2T + M - 1, S = T + M:  out[t, j] = x[t, T - 1 - t + j]."""
    b, n, t, r = x.shape
    tt = np.arange(t)[:, None]
    jj = np.arange(s)[None, :]
    idx = t - 1 - tt + jj
    return np.take_along_axis(x, np.broadcast_to(idx, (b, n, t, s)), axis=3)


def compute_causal_mask(t: int, s: int | None = None) -> np.ndarray:
    s = t if s is None else s
    return np.tril(np.ones((1, t, s), bool))


def compute_streaming_mask(chunk_size: int, history_size: int, t: int, s: int | None = None) -> np.ndarray:
    s = t if s is None else s
    hist = s if history_size < 0 else history_size
    q = np.arange(t) // chunk_size
    start = np.maximum(0, q * chunk_size - hist)
    end = np.minimum(s, q * chunk_size + chunk_size)
    k = np.arange(s)[None, :]
    return ((k < end[:, None]) & (k >= start[:, None]))[None]


def masked_softmax(scores: np.ndarray, mask: np.ndarray | None, fill: float = -1e9) -> np.ndarray:
    if mask is not None:
        scores = np.where(np.broadcast_to(mask, scores.shape), scores, np.asarray(fill, scores.dtype))
    m = scores.max(-1, keepdims=True)
    e = np.exp(scores - m)
    return e / e.sum(-1, keepdims=True)
