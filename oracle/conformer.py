"""CPU oracle: Conformer transducer forward (and, via torch autograd, backward) — TEST INFRASTRUCTURE ONLY.

A functional torch-CPU restatement (float32 or float64) of the reference modules below; it never runs in the product
path.  It is the parity checker for the CUDA encoder / prediction / joint kernels, the gradient checker (autograd), and
the "port" CPU arm of bench.py (`--impl reference` / `cpu_baseline`), because TensorFlow/Keras are not installable here.

Restated reference code (paths relative to tensorflow_asr/):
  * models/layers/subsampling.py:163-230        Conv2dSubsampling (causal 3x3 stride-2 conv -> BN|LN -> swish) x2, merge (F, C)
  * models/layers/convolution.py:25-37,132-144  causal padding = (kernel-1) on the top/left of each spatial dim, then VALID
  * utils/math_util.py:282-305                  conv_output_length ("causal": ceil(len / stride))
  * models/encoders/conformer.py:672-701        ConformerEncoder.call: subsampling -> linear -> (dropout) -> relpe -> blocks
  * models/layers/positional_encoding.py:31-52,114-174  interleaved sinusoid table over positions [T-1..-(T-1)], per-utterance
                                                roll by -(T - len_b) and zero mask beyond 2*len_b - 1 (not added to x)
  * models/encoders/conformer.py:101-109        FFModule      x + 0.5 * W2 swish(W1 LN(x))
  * models/encoders/conformer.py:209-239        MHSAModule    x + MHSA(LN(x))
  * models/layers/multihead_attention.py:543-667 relative MHSA: (q+u)k^T, (q+v)p^T, rel_left_shift, QUERY-side mask only,
                                                masked softmax with -1e9 fill (models/layers/general.py:30-41)
  * models/encoders/conformer.py:366-377        ConvModule    x + PW2 swish(BN|LN(DW31 causal(GLU(PW1 LN(x)))))
  * models/activations/glu.py:25-28             GLU: a * sigmoid(b), a = first half
  * models/encoders/conformer.py:504-535        ConformerBlock: FFN/2 -> MHSA -> Conv -> FFN/2 -> LN
  * models/transducer/base_transducer.py:123-132  prediction net: Embedding -> LSTM (Keras gate order i,f,c,o; masked steps carry
                                                state and output zeros) -> LN
  * models/transducer/base_transducer.py:199-207,280-293  joint: Wv tanh(We enc[:, :, None] + Wp pred[:, None])
  * models/transducer/base_transducer.py:427-435  Transducer.call glue
Keras layer semantics that live in the absent dependency keras-nightly~=3.9 (requirements.txt:11) are restated from its
published behaviour: LayerNormalization eps=1e-3 over the last axis; BatchNormalization eps=1e-3, momentum 0.99, biased
batch variance over all leading axes in training (padded frames included: Conv/Dense layers drop the Keras mask before BN),
moving statistics in inference; Dense = x @ kernel + bias; EinsumDense projections [D,N,H] / [N,H,D].

PARITY PINNING: no reference test pins any of these outputs and TF/Keras cannot run here => **parity unpinned against
the reference**; the component formulas that the reference does pin (streaming mask, rel-shift) are covered in oracle/masks.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from . import rnnt as _rnnt

LN_EPS = 1e-3
BN_EPS = 1e-3
BN_MOMENTUM = 0.99


@dataclass
class ConformerConfig:
    """examples/models/transducer/conformer/small.yml.j2:5-65 (Conformer-S)."""

    vocab_size: int = 1000
    num_feature_bins: int = 80
    dmodel: int = 144
    num_blocks: int = 16
    head_size: int = 36
    num_heads: int = 4
    kernel_size: int = 31
    ffm_scale: int = 4
    subsampling_filters: tuple = (144, 144)
    subsampling_norms: tuple = ("batch", "batch")
    dw_norm_type: str = "batch"
    ffm_residual_factor: float = 0.5
    mhsam_residual_factor: float = 1.0
    convm_residual_factor: float = 1.0
    embed_dim: int = 320
    rnn_units: int = 320
    joint_dim: int = 320
    history_size: int | None = None
    chunk_size: int | None = None
    extra: dict = field(default_factory=dict)


def conv_output_length(length, stride=2):
    return (length + stride - 1) // stride


# ------------------------------------------------------------------------------------------------ parameters
def init_params(cfg: ConformerConfig, seed: int = 42, dtype=torch.float32, randomize_all: bool = False) -> dict:
    """Keras default initialisers restated (glorot_uniform kernels, zero biases, uniform(-0.05, 0.05) embedding, LSTM
    unit_forget_bias); `randomize_all` perturbs every tensor (norm scales/offsets, biases, moving stats) so that parity
    tests exercise every term."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def glorot(*shape, fan_in=None, fan_out=None):
        if fan_in is None:
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1).mul(lim).to(dtype)

    def rnd(*shape, scale=0.1, base=0.0):
        if randomize_all:
            return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale + base).to(dtype)
        return torch.full(shape, base, dtype=dtype)

    def norm(prefix, n, bn=False):
        p[prefix + ".gamma"] = rnd(n, base=1.0)
        p[prefix + ".beta"] = rnd(n)
        if bn:
            p[prefix + ".moving_mean"] = rnd(n)
            p[prefix + ".moving_variance"] = rnd(n, scale=0.05, base=1.0).abs()

    D, N, H = cfg.dmodel, cfg.num_heads, cfg.head_size
    cin = 1
    f = cfg.num_feature_bins
    for i, (co, nm) in enumerate(zip(cfg.subsampling_filters, cfg.subsampling_norms)):
        p[f"encoder.subsampling.conv_{i}.kernel"] = glorot(3, 3, cin, co)
        p[f"encoder.subsampling.conv_{i}.bias"] = rnd(co)
        if nm == "batch":
            norm(f"encoder.subsampling.bn_{i}", co, bn=True)
        elif nm == "layer":
            norm(f"encoder.subsampling.ln_{i}", co)
        cin = co
        f = conv_output_length(f)
    p["encoder.linear.kernel"] = glorot(f * cin, D)
    p["encoder.linear.bias"] = rnd(D)
    p["encoder.content_attention_bias"] = rnd(N, H)
    p["encoder.positional_attention_bias"] = rnd(N, H)
    for b in range(cfg.num_blocks):
        pre = f"encoder.block_{b}"
        for ff in ("ff_module_1", "ff_module_2"):
            norm(f"{pre}.{ff}.ln", D)
            p[f"{pre}.{ff}.dense_1.kernel"] = glorot(D, cfg.ffm_scale * D)
            p[f"{pre}.{ff}.dense_1.bias"] = rnd(cfg.ffm_scale * D)
            p[f"{pre}.{ff}.dense_2.kernel"] = glorot(cfg.ffm_scale * D, D)
            p[f"{pre}.{ff}.dense_2.bias"] = rnd(D)
        norm(f"{pre}.mhsa_module.ln", D)
        for nm in ("query", "key", "value", "encoding"):
            p[f"{pre}.mhsa_module.mhsa.{nm}.kernel"] = glorot(D, N, H, fan_in=D, fan_out=N * H)
            p[f"{pre}.mhsa_module.mhsa.{nm}.bias"] = rnd(N, H)
        p[f"{pre}.mhsa_module.mhsa.attention_output.kernel"] = glorot(N, H, D, fan_in=N * H, fan_out=D)
        p[f"{pre}.mhsa_module.mhsa.attention_output.bias"] = rnd(D)
        norm(f"{pre}.conv_module.ln", D)
        p[f"{pre}.conv_module.pw_conv_1.kernel"] = glorot(D, 2 * D)
        p[f"{pre}.conv_module.pw_conv_1.bias"] = rnd(2 * D)
        p[f"{pre}.conv_module.dw_conv.kernel"] = glorot(cfg.kernel_size, D, fan_in=cfg.kernel_size, fan_out=cfg.kernel_size)
        p[f"{pre}.conv_module.dw_conv.bias"] = rnd(D)
        if cfg.dw_norm_type == "batch":
            norm(f"{pre}.conv_module.dw_bn", D, bn=True)
        else:
            norm(f"{pre}.conv_module.dw_ln", D)
        p[f"{pre}.conv_module.pw_conv_2.kernel"] = glorot(D, D)
        p[f"{pre}.conv_module.pw_conv_2.bias"] = rnd(D)
        norm(f"{pre}.ln", D)
    E, R, J, V = cfg.embed_dim, cfg.rnn_units, cfg.joint_dim, cfg.vocab_size
    p["prediction.embedding.embeddings"] = (torch.rand(V, E, generator=g, dtype=torch.float64) * 0.1 - 0.05).to(dtype)
    p["prediction.lstm_0.kernel"] = glorot(E, 4 * R)
    q, _ = torch.linalg.qr(torch.randn(4 * R, R, generator=g, dtype=torch.float64))
    p["prediction.lstm_0.recurrent_kernel"] = q.T.contiguous().to(dtype)  # orthogonal [R, 4R]
    bias = rnd(4 * R)
    bias[R : 2 * R] += 1.0  # unit_forget_bias
    p["prediction.lstm_0.bias"] = bias
    norm("prediction.ln_0", R)
    p["joint.enc.kernel"] = glorot(D, J)
    p["joint.enc.bias"] = rnd(J)
    p["joint.pred.kernel"] = glorot(R, J)
    p["joint.pred.bias"] = rnd(J)
    p["joint.vocab.kernel"] = glorot(J, V)
    p["joint.vocab.bias"] = rnd(V)
    return p


# ------------------------------------------------------------------------------------------------ layers
def layer_norm(x, p, prefix):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + LN_EPS) * p[prefix + ".gamma"] + p[prefix + ".beta"]


def batch_norm(x, p, prefix, training, stats_out=None):
    """Normalises over every axis but the last.  training: biased batch statistics (+ moving-average update reported through
    `stats_out`); inference: moving statistics."""
    if training:
        dims = tuple(range(x.dim() - 1))
        mean = x.mean(dims)
        var = ((x - mean) ** 2).mean(dims)
        if stats_out is not None:
            stats_out[prefix] = (mean.detach(), var.detach())
    else:
        mean, var = p[prefix + ".moving_mean"], p[prefix + ".moving_variance"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * p[prefix + ".gamma"] + p[prefix + ".beta"]


def swish(x):
    return x * torch.sigmoid(x)


def conv2d_subsampling(x, length, p, cfg, training, stats_out=None):
    """x [B,T,F,1] -> ([B,T',F'*C], len')."""
    y = x.permute(0, 3, 1, 2)  # NCHW, H = time, W = freq
    for i, nm in enumerate(cfg.subsampling_norms):
        w = p[f"encoder.subsampling.conv_{i}.kernel"].permute(3, 2, 0, 1)  # [kh,kw,cin,cout] -> [cout,cin,kh,kw]
        y = F.pad(y, (2, 0, 2, 0))  # causal: kernel-1 on the left (freq) and top (time)
        y = F.conv2d(y, w, p[f"encoder.subsampling.conv_{i}.bias"], stride=2)
        y = y.permute(0, 2, 3, 1)  # NHWC
        if nm == "batch":
            y = batch_norm(y, p, f"encoder.subsampling.bn_{i}", training, stats_out)
        elif nm == "layer":
            y = layer_norm(y, p, f"encoder.subsampling.ln_{i}")
        y = swish(y).permute(0, 3, 1, 2)
        length = conv_output_length(length)
    y = y.permute(0, 2, 3, 1)  # [B,T',F',C]
    return y.reshape(y.shape[0], y.shape[1], -1), length


def relative_position_encoding(T, length, dmodel, dtype, memory_length=0):
    """[B, 2T+M-1, D]; positional_encoding.py:114-174 (non-causal branch, interleave=True)."""
    pos = torch.cat([torch.arange(T + memory_length - 1, 0, -1), torch.arange(0, -T, -1)]).to(dtype)
    d = torch.arange(dmodel)
    timescales = torch.pow(torch.tensor(1.0 / 10000.0, dtype=dtype), (2 * (d // 2)).to(dtype) / dmodel)
    ang = pos[:, None] * timescales[None, :]
    pe = torch.where((d % 2 == 0)[None, :], torch.sin(ang), torch.cos(ang))  # even dims sine, odd cosine
    R = pe.shape[0]
    out = []
    for lb in length.tolist():
        rolled = torch.roll(pe, shifts=-(T - lb), dims=0)
        mask = (torch.arange(R) < 2 * lb + memory_length - 1).to(dtype)[:, None]
        out.append(rolled * mask)
    return torch.stack(out)


def streaming_mask(chunk, history, T):
    q = torch.arange(T) // chunk
    hist = T if history < 0 else history
    start = torch.clamp(q * chunk - hist, min=0)
    end = torch.clamp(q * chunk + chunk, max=T)
    k = torch.arange(T)[None, :]
    return (k < end[:, None]) & (k >= start[:, None])


def rel_mhsa(x, relpe, length, p, prefix, cfg, ubias, vbias):
    B, T, D = x.shape
    N, H = cfg.num_heads, cfg.head_size
    q = torch.einsum("btd,dnh->btnh", x, p[prefix + ".query.kernel"]) + p[prefix + ".query.bias"]
    k = torch.einsum("btd,dnh->btnh", x, p[prefix + ".key.kernel"]) + p[prefix + ".key.bias"]
    v = torch.einsum("btd,dnh->btnh", x, p[prefix + ".value.kernel"]) + p[prefix + ".value.bias"]
    pos = torch.einsum("brd,dnh->brnh", relpe, p[prefix + ".encoding.kernel"]) + p[prefix + ".encoding.bias"]
    scale = 1.0 / math.sqrt(H)
    content = torch.einsum("bsnh,btnh->bnts", k, (q + ubias) * scale)
    positional = torch.einsum("brnh,btnh->bntr", pos, (q + vbias) * scale)
    # rel_left_shift (non-causal) + last-S slice == out[t, j] = positional[t, T-1-t+j]   (oracle/masks.py proves the identity)
    idx = (T - 1 - torch.arange(T)[:, None] + torch.arange(T)[None, :]).expand(B, N, T, T)
    positional = torch.gather(positional, 3, idx)
    scores = content + positional
    mask = (torch.arange(T)[None, :] < length[:, None])[:, None, :, None]  # QUERY-side mask only (multihead_attention.py:614-623)
    if cfg.chunk_size is not None and cfg.history_size is not None:
        mask = mask & streaming_mask(cfg.chunk_size, cfg.history_size, T)[None, None]
    scores = torch.where(mask.expand_as(scores), scores, torch.tensor(-1e9, dtype=scores.dtype))
    attn = torch.softmax(scores, dim=-1)
    out = torch.einsum("bnts,bsnh->btnh", attn, v)
    return torch.einsum("btnh,nhd->btd", out, p[prefix + ".attention_output.kernel"]) + p[prefix + ".attention_output.bias"]


def ff_module(x, p, prefix, cfg):
    y = layer_norm(x, p, prefix + ".ln")
    y = swish(y @ p[prefix + ".dense_1.kernel"] + p[prefix + ".dense_1.bias"])
    y = y @ p[prefix + ".dense_2.kernel"] + p[prefix + ".dense_2.bias"]
    return x + cfg.ffm_residual_factor * y


def conv_module(x, p, prefix, cfg, training, stats_out=None):
    y = layer_norm(x, p, prefix + ".ln")
    y = y @ p[prefix + ".pw_conv_1.kernel"] + p[prefix + ".pw_conv_1.bias"]
    a, b = y.chunk(2, dim=-1)
    y = a * torch.sigmoid(b)
    K = cfg.kernel_size
    w = p[prefix + ".dw_conv.kernel"]  # [K, C]
    yp = F.pad(y.transpose(1, 2), (K - 1, 0))  # causal
    y = F.conv1d(yp, w.t().unsqueeze(1), p[prefix + ".dw_conv.bias"], groups=w.shape[1]).transpose(1, 2)
    if cfg.dw_norm_type == "batch":
        y = batch_norm(y, p, prefix + ".dw_bn", training, stats_out)
    else:
        y = layer_norm(y, p, prefix + ".dw_ln")
    y = swish(y)
    y = y @ p[prefix + ".pw_conv_2.kernel"] + p[prefix + ".pw_conv_2.bias"]
    return x + cfg.convm_residual_factor * y


def conformer_encoder(features, length, p, cfg, training=False, stats_out=None, return_intermediates=False):
    """features [B,T,F,1], length [B] int64 -> (enc [B,T',D], len')."""
    inter = {}
    x, length = conv2d_subsampling(features, length, p, cfg, training, stats_out)
    inter["subsampling"] = x
    x = x @ p["encoder.linear.kernel"] + p["encoder.linear.bias"]
    inter["linear"] = x
    relpe = relative_position_encoding(x.shape[1], length, cfg.dmodel, x.dtype)
    inter["relpe"] = relpe
    u, v = p["encoder.content_attention_bias"], p["encoder.positional_attention_bias"]
    for b in range(cfg.num_blocks):
        pre = f"encoder.block_{b}"
        x = ff_module(x, p, pre + ".ff_module_1", cfg)
        y = layer_norm(x, p, pre + ".mhsa_module.ln")
        x = x + cfg.mhsam_residual_factor * rel_mhsa(y, relpe, length, p, pre + ".mhsa_module.mhsa", cfg, u, v)
        x = conv_module(x, p, pre + ".conv_module", cfg, training, stats_out)
        x = ff_module(x, p, pre + ".ff_module_2", cfg)
        x = layer_norm(x, p, pre + ".ln")
        inter[f"block_{b}"] = x
    if return_intermediates:
        return x, length, inter
    return x, length


def lstm(x, length, p, prefix, state=None):
    """Keras LSTM (gate order i, f, c, o), masked steps carry state and emit zeros (zero_output_for_mask=True)."""
    B, U, _ = x.shape
    R = p[prefix + ".recurrent_kernel"].shape[0]
    h = x.new_zeros(B, R) if state is None else state[0]
    c = x.new_zeros(B, R) if state is None else state[1]
    xs = x @ p[prefix + ".kernel"] + p[prefix + ".bias"]
    outs = []
    for t in range(U):
        z = xs[:, t] + h @ p[prefix + ".recurrent_kernel"]
        i, f, g, o = z.chunk(4, dim=-1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        if length is not None:
            m = (t < length)[:, None]
            outs.append(torch.where(m, h_new, torch.zeros_like(h_new)))
            h, c = torch.where(m, h_new, h), torch.where(m, c_new, c)
        else:
            outs.append(h_new)
            h, c = h_new, c_new
    return torch.stack(outs, dim=1), (h, c)


def prediction_net(predictions, predictions_length, p, cfg):
    emb = p["prediction.embedding.embeddings"][predictions.long()]
    y, _ = lstm(emb, predictions_length, p, "prediction.lstm_0")
    return layer_norm(y, p, "prediction.ln_0")


def joint_net(enc, pred, p):
    e = enc @ p["joint.enc.kernel"] + p["joint.enc.bias"]
    q = pred @ p["joint.pred.kernel"] + p["joint.pred.bias"]
    h = torch.tanh(e[:, :, None, :] + q[:, None, :, :])
    return h @ p["joint.vocab.kernel"] + p["joint.vocab.bias"]


def transducer_forward(p, cfg, features, features_length, predictions, predictions_length, training=False, stats_out=None):
    """Transducer.call after the frontend: -> (logits [B,T',U+1,V], logits_length)."""
    enc, length = conformer_encoder(features, features_length, p, cfg, training, stats_out)
    pred = prediction_net(predictions, predictions_length, p, cfg)
    return joint_net(enc, pred, p), length


class _RnntLossOracleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, logit_length, label_length):
        loss, grads = _rnnt.rnnt_loss(logits.detach().numpy(), labels.numpy(), logit_length.numpy(), label_length.numpy(),
                                      dtype=np.float64 if logits.dtype == torch.float64 else np.float32)
        ctx.save_for_backward(torch.from_numpy(grads).to(logits.dtype))
        return torch.from_numpy(loss).to(logits.dtype)

    @staticmethod
    def backward(ctx, g):
        (grads,) = ctx.saved_tensors
        return grads * g[:, None, None, None], None, None, None


def rnnt_loss_torch(logits, labels, logit_length, label_length, fast=False):
    """Differentiable per-utterance RNN-T loss on CPU.  fast=True uses torchaudio's C++ kernel (the CPU baseline arm's
    timing path); otherwise the numpy oracle (oracle/rnnt.py)."""
    if fast:
        import torchaudio

        ll = torch.maximum(logit_length.int(), label_length.int())
        return torchaudio.functional.rnnt_loss(logits.float(), labels.int(), ll, label_length.int(), blank=0, reduction="none")
    return _RnntLossOracleFn.apply(logits, labels, logit_length, label_length)


def _call_next(frame, tok, h, c, p):
    """Transducer.call_next (base_transducer.py:437-464) for a batch: frame [B,1,D], tok [B,1] -> (log-probs [B,V], h, c)."""
    emb = p["prediction.embedding.embeddings"][tok.long()]
    y, (h2, c2) = lstm(emb, None, p, "prediction.lstm_0", state=(h, c))
    y = layer_norm(y, p, "prediction.ln_0")
    return torch.log_softmax(joint_net(frame, y, p)[:, 0, 0], -1), h2, c2


def _margin(lp):
    top = torch.topk(lp, 2, dim=-1).values
    return float((top[..., 0] - top[..., 1]).min())


def greedy_single(enc, nframes, p, cfg, blank=0, max_tokens_per_frame=3):
    """Transducer.recognize_single restated (base_transducer.py:577-712); enc [1,T,D].  Returns (tokens list incl. blank padding,
    smallest top-2 margin seen: tells a test whether an argmax was a near-tie)."""
    R = cfg.rnn_units
    h, c = enc.new_zeros(1, R), enc.new_zeros(1, R)
    tok = torch.tensor([[blank]])
    toks = [blank] * (nframes * max_tokens_per_frame)
    frame, token_index, per_frame, margin = 0, -1, 0, float("inf")
    while frame < nframes:
        lp, h2, c2 = _call_next(enc[:, frame : frame + 1], tok, h, c, p)
        margin = min(margin, _margin(lp))
        k = int(lp[0].argmax())
        if k == blank:
            frame, per_frame = frame + 1, 0
        else:
            per_frame += 1
            token_index += 1
            toks[token_index] = k
            tok = torch.tensor([[k]])
            h, c = h2, c2
            if per_frame >= max_tokens_per_frame:
                frame, per_frame = frame + 1, 0
    return toks, margin


def greedy_batch(enc, enc_len, p, cfg, blank=0):
    """Transducer.recognize_batch restated verbatim (base_transducer.py:496-575), quirks included.  enc [B,T,D]."""
    B, T, _ = enc.shape
    R = cfg.rnn_units
    h, c = enc.new_zeros(B, R), enc.new_zeros(B, R)
    prev = torch.full((B, 1), blank, dtype=torch.long)
    nframes = enc_len.long()[:, None]
    frame_idx = torch.zeros(B, 1, dtype=torch.long)
    max_tokens = 2 * T + 1
    tokens = torch.full((B, max_tokens), blank, dtype=torch.long)
    tok_idx = torch.ones(B, 1, dtype=torch.long)
    rows = torch.arange(B)
    margin = float("inf")
    while not (bool((frame_idx >= nframes - 1).all()) or bool((tok_idx >= max_tokens - 1).all())):
        cur_frames = enc[rows, torch.minimum(frame_idx, nframes - 1)[:, 0]].unsqueeze(1)
        lp, h2, c2 = _call_next(cur_frames, prev, h, c, p)
        cur = lp.argmax(-1)[:, None]
        eq_blank = (cur == blank) | (tok_idx >= max_tokens) | (frame_idx > nframes)
        live = ~((tok_idx >= max_tokens) | (frame_idx > nframes))[:, 0]
        if bool(live.any()):
            margin = min(margin, _margin(lp[live]))
        upd_tok = torch.where(eq_blank, torch.full_like(cur, blank), cur)[:, 0]
        upd_idx = torch.where(eq_blank, torch.zeros_like(tok_idx), torch.clamp(tok_idx + 1, max=max_tokens - 1))[:, 0]
        tokens[rows, upd_idx] = upd_tok
        tok_idx = torch.where(eq_blank, tok_idx, torch.clamp(tok_idx + 1, max=max_tokens - 1))
        frame_idx = torch.where(eq_blank, frame_idx + 1, frame_idx)
        prev = torch.where(eq_blank, prev, cur)
        h = torch.where(eq_blank, h, h2)
        c = torch.where(eq_blank, c, c2)
    return tokens, margin
