"""CPU oracle: CTC loss and gradient (TEST INFRASTRUCTURE, never the product path).

Reference call site: ``tensorflow_asr/losses/ctc_loss.py:47-66`` — on GPU the reference calls
``tf.nn.ctc_loss(logits [B,T,V], labels [B,L] dense, label_length, logit_length,
logits_time_major=False, blank_index=0)`` after the length clamp of
``tensorflow_asr/losses/base_loss.py:33-37``.  ``tf.nn.ctc_loss`` lives in the absent
third-party dependency ``tensorflow~=2.18.0`` (requirements.gpu.txt:1); its published algorithm
(Graves et al. 2006; TF ``ctc_loss_dense``/``ctc_loss_and_grad``) is restated here:
  * logits -> log_softmax over V;
  * states s = 0..2L: even = blank, odd = labels[(s-1)/2];
  * alpha_t(s) = lp_t(l_s) + logsumexp(alpha_{t-1}(s), alpha_{t-1}(s-1), [alpha_{t-1}(s-2) if l_s != blank and l_s != l_{s-2}])
  * loss = -logsumexp(alpha_{T-1}(2L), alpha_{T-1}(2L-1));
  * d loss / d logits[t,k] = softmax[t,k] - exp(logsumexp_{s: l_s = k}(alpha_t(s) + beta_t(s)) - lp_t(k) - logZ)
    for t < logit_length, 0 beyond it.
The in-tree TPU variant ``tensorflow_asr/losses/impl/ctc_tpu.py:1295-1314`` computes the same
quantity with a (label index, open/closed) state parametrisation; an infeasible alignment
(T < L + repeats) gives loss=+inf and zero gradient there (:547-551); this oracle returns
+inf and a zero gradient in that case as well.

PARITY PINNING: no reference test covers CTC => **parity unpinned against the reference**;
cross-checked against ``torch.nn.functional.ctc_loss`` (independent implementation) in
tests/test_oracle_ctc.py.
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def _lse(*xs):
    m = np.maximum.reduce(xs)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    s = sum(np.exp(x - m_safe) for x in xs)
    with np.errstate(divide="ignore"):
        return np.where(np.isfinite(m), np.log(s) + m_safe, NEG_INF)


def ctc_loss_and_grad(logits, labels, logit_length, label_length, blank=0, dtype=np.float64):
    """logits [B,T,V]; labels [B,L]; returns (loss [B], grad wrt logits [B,T,V])."""
    logits = np.asarray(logits, dtype)
    B, T, V = logits.shape
    m = logits.max(-1, keepdims=True)
    lp = logits - m - np.log(np.exp(logits - m).sum(-1, keepdims=True))
    loss = np.zeros(B, dtype)
    grad = np.zeros_like(logits)
    for b in range(B):
        Tb, L = int(logit_length[b]), int(label_length[b])
        S = 2 * L + 1
        ext = np.full(S, blank, np.int64)
        ext[1::2] = np.asarray(labels[b][:L])
        skip = np.zeros(S, bool)
        if S > 2:
            skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        alpha = np.full((Tb, S), NEG_INF, dtype)
        alpha[0, 0] = lp[b, 0, blank]
        if S > 1:
            alpha[0, 1] = lp[b, 0, ext[1]]
        for t in range(1, Tb):
            prev = alpha[t - 1]
            pad = np.concatenate([[NEG_INF, NEG_INF], prev])
            p1 = pad[1 : S + 1]
            p2 = np.where(skip, pad[:S], NEG_INF)
            alpha[t] = _lse(prev, p1, p2) + lp[b, t, ext]
        beta = np.full((Tb, S), NEG_INF, dtype)
        beta[Tb - 1, S - 1] = lp[b, Tb - 1, blank]
        if S > 1:
            beta[Tb - 1, S - 2] = lp[b, Tb - 1, ext[S - 2]]
        skip_f = np.zeros(S, bool)  # can s jump to s+2
        if S > 2:
            skip_f[:-2] = skip[2:]
        for t in range(Tb - 2, -1, -1):
            nxt = beta[t + 1]
            pad = np.concatenate([nxt, [NEG_INF, NEG_INF]])
            n1 = pad[1 : S + 1]
            n2 = np.where(skip_f, pad[2 : S + 2], NEG_INF)
            beta[t] = _lse(nxt, n1, n2) + lp[b, t, ext]
        logz = _lse(alpha[Tb - 1, S - 1], alpha[Tb - 1, S - 2] if S > 1 else np.asarray(NEG_INF))
        if not np.isfinite(logz):
            loss[b] = np.inf
            continue
        loss[b] = -logz
        ab = alpha + beta  # includes lp twice
        g = np.exp(lp[b, :Tb])
        occ = np.full((Tb, V), NEG_INF, dtype)
        for s in range(S):
            occ[:, ext[s]] = _lse(occ[:, ext[s]], ab[:, s])
        with np.errstate(invalid="ignore"):
            g = g - np.where(np.isfinite(occ), np.exp(occ - lp[b, :Tb] - logz), 0.0)
        grad[b, :Tb] = g
    return loss, grad


def ctc_loss(logits, labels, logit_length, label_length, blank=0, dtype=np.float64):
    logit_length = np.asarray(logit_length, np.int32)
    label_length = np.asarray(label_length, np.int32)
    logit_length = np.where(logit_length < label_length, label_length, logit_length)  # base_loss.py:36
    return ctc_loss_and_grad(logits, labels, logit_length, label_length, blank, dtype)
