"""CPU oracle: RNN-T loss and gradient (TEST INFRASTRUCTURE, never the product path).

Restates ``tensorflow_asr/losses/impl/rnnt.py`` (the pure-TF path, which is what the
reference runs when warp-transducer is not installed, impl/rnnt.py:30):
  * ``transition_probs``                  :94-105   blank / truth log-probs
  * ``forward_dp``                        :108-137  alpha
  * ``backward_dp``                       :140-178  beta, seeded at (T_b-1, U_b) with the blank log-prob
  * ``compute_rnnt_loss_and_grad_helper`` :181-278  loss = -beta[0,0]; g_blank / g_truth; softmax Jacobian
  * ``rnnt_loss_tf`` custom gradient      :321-324  upstream gradient scales per utterance
and the length clamp of ``tensorflow_asr/losses/base_loss.py:33-37``.

The reference evaluates the recurrences one anti-diagonal at a time with tf.scan;
the recurrences themselves are element-wise
    alpha[t,u] = logaddexp(alpha[t-1,u] + blank[t-1,u], alpha[t,u-1] + truth[t,u-1])
    beta[t,u]  = logaddexp(beta[t+1,u] + blank[t,u],   beta[t,u+1] + truth[t,u])
and are restated here row by row (same values, different evaluation order).

The alternative FFI backend (warp-transducer, impl/rnnt.py:41-58) is an un-vendored,
un-pinned external (scripts/install_rnnt_loss.sh:13, `git clone --depth 1` HEAD of
github.com/nglehuy/warp-transducer); it computes the same loss and the same
gradient w.r.t. the logits.

PARITY PINNING: the reference's own test (tests/test_rnnt_loss.py) asserts nothing,
and TF is not installable here.  The oracle is pinned against (1) the
warp-transducer known-answer vector shipped with that library's tests
(B=1,T=2,U+1=3,V=5, labels [1,2] -> cost 4.495666, plus its published gradient) and
(2) ``torchaudio.functional.rnnt_loss`` (an independent implementation) on random
cases, loss and gradient — see tests/test_oracle_rnnt.py.
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def log_softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def clamp_lengths(logit_length, label_length):
    """base_loss.py:33-37: logit_length = max(logit_length, label_length)."""
    logit_length = np.asarray(logit_length, np.int32)
    label_length = np.asarray(label_length, np.int32)
    return np.where(logit_length < label_length, label_length, logit_length), label_length


def rnnt_loss_and_grad(logits, labels, logit_length, label_length, blank=0, dtype=np.float64):
    """Returns (loss [B], grads wrt logits [B,T,U+1,V], alpha, beta).

    logits [B,T,U+1,V]; labels [B,U] int; lengths [B] int.  blank must be 0 (base_loss.py:24).
    """
    assert blank == 0, "Only support blank=0"
    logits = np.asarray(logits, dtype)
    B, T, U1, V = logits.shape
    labels = np.asarray(labels)
    assert labels.shape == (B, U1 - 1)
    lp = log_softmax(logits)  # impl/rnnt.py:211
    blank_lp = lp[..., 0]  # [B,T,U1]
    truth_lp = np.take_along_axis(lp[:, :, :-1, :], np.broadcast_to(labels[:, None, :, None], (B, T, U1 - 1, 1)), axis=3)[..., 0]
    loss = np.zeros(B, dtype)
    grads = np.zeros_like(logits)
    alphas = np.full((B, T, U1), NEG_INF, dtype)
    betas = np.full((B, T, U1), NEG_INF, dtype)
    for b in range(B):
        Tb, Ub = int(logit_length[b]), int(label_length[b])
        bl, tr = blank_lp[b], truth_lp[b]
        a = alphas[b]
        a[0, 0] = 0.0
        for u in range(1, Ub + 1):
            a[0, u] = a[0, u - 1] + tr[0, u - 1]
        for t in range(1, Tb):
            a[t, 0] = a[t - 1, 0] + bl[t - 1, 0]
            # the u-recurrence within a row is sequential
            for u in range(1, Ub + 1):
                a[t, u] = np.logaddexp(a[t - 1, u] + bl[t - 1, u], a[t, u - 1] + tr[t, u - 1])
        be = betas[b]
        be[Tb - 1, Ub] = bl[Tb - 1, Ub]
        for u in range(Ub - 1, -1, -1):
            be[Tb - 1, u] = be[Tb - 1, u + 1] + tr[Tb - 1, u]
        for t in range(Tb - 2, -1, -1):
            be[t, Ub] = be[t + 1, Ub] + bl[t, Ub]
            for u in range(Ub - 1, -1, -1):
                be[t, u] = np.logaddexp(be[t + 1, u] + bl[t, u], be[t, u + 1] + tr[t, u])
        logz = be[0, 0]
        loss[b] = -logz
        # gradients w.r.t. log-probs (impl/rnnt.py:233-253)
        g_blank = np.zeros((T, U1), dtype)
        g_truth = np.zeros((T, U1), dtype)
        if Tb > 1:
            g_blank[: Tb - 1, : Ub + 1] = -np.exp(a[: Tb - 1, : Ub + 1] + be[1:Tb, : Ub + 1] - logz + bl[: Tb - 1, : Ub + 1])
        g_blank[Tb - 1, Ub] = -1.0
        if Ub > 0:
            g_truth[:Tb, :Ub] = -np.exp(a[:Tb, :Ub] + be[:Tb, 1 : Ub + 1] - logz + tr[:Tb, :Ub])
        g = np.zeros((T, U1, V), dtype)
        g[:, :, 0] = g_blank
        tt, uu = np.meshgrid(np.arange(Tb), np.arange(Ub), indexing="ij")
        if Ub > 0:
            np.add.at(g, (tt, uu, labels[b][uu]), g_truth[:Tb, :Ub])
        # softmax Jacobian (impl/rnnt.py:275)
        grads[b] = g - np.exp(lp[b]) * g.sum(axis=-1, keepdims=True)
    return loss, grads, alphas, betas


def rnnt_loss(logits, labels, logit_length, label_length, blank=0, dtype=np.float64):
    logit_length, label_length = clamp_lengths(logit_length, label_length)
    loss, grads, _, _ = rnnt_loss_and_grad(logits, labels, logit_length, label_length, blank, dtype)
    return loss, grads


# warp-transducer's shipped known-answer test (tests/test_cpu.cpp `small_test`, also used by
# torchaudio's test-suite): acts are already-normalised probabilities-like activations.
WARP_TRANSDUCER_KAT = dict(
    acts=np.array(
        [0.1, 0.6, 0.1, 0.1, 0.1, 0.1, 0.1, 0.6, 0.1, 0.1, 0.1, 0.1, 0.2, 0.8, 0.1,
         0.1, 0.6, 0.1, 0.1, 0.1, 0.1, 0.1, 0.2, 0.1, 0.1, 0.7, 0.1, 0.2, 0.1, 0.1], dtype=np.float64
    ).reshape(1, 2, 3, 5),
    labels=np.array([[1, 2]], dtype=np.int32),
    cost=4.495666,
    grads=np.array(
        [-0.13116688, -0.3999269, 0.17703125, 0.17703125, 0.17703125,
         -0.18572757, 0.12247056, -0.18168412, 0.12247056, 0.12247056,
         -0.32091254, 0.06269141, 0.06928472, 0.12624499, 0.06269141,
         0.05456069, -0.21824276, 0.05456069, 0.05456069, 0.05456069,
         0.12073959, 0.12073959, -0.48295835, 0.12073959, 0.12073959,
         -0.6925882, 0.16871116, 0.18645467, 0.16871116, 0.16871116], dtype=np.float64
    ).reshape(1, 2, 3, 5),
)
