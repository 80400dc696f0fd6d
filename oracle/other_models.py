"""CPU oracle: DeepSpeech2 (CTC) and ContextNet encoders — TEST INFRASTRUCTURE ONLY.

Restates (paths relative to tensorflow_asr/):
  * models/encoders/deepspeech2.py:56-60,105-117 (Conv2D 'same' + BN + relu), :250-256 (Bidirectional LSTM, masked), :367-373 (FC),
    models/ctc/deepspeech2.py:22-54 (vocab Dense), models/ctc/base_ctc.py:74-81,102-126 (call, greedy = tf.nn.ctc_greedy_decoder)
  * models/encoders/contextnet.py:74-86 (SeparableConv1D causal + BN + act), :152-164 (squeeze-excite with masked global average),
    :251-262 (block: convs -> SE -> + residual -> act), :306-311 (encoder)
Keras / TF semantics restated from their published behaviour: 'same' padding = ceil(n/s) outputs with the extra pad at the END;
Bidirectional(LSTM) with a mask processes each utterance's valid frames in both directions and emits zeros on padded frames;
GlobalAveragePooling1D with a mask averages valid frames only.  **Parity unpinned against the reference** (no reference test).
Per-utterance loops are used deliberately (different evaluation order from the product's packed / batched code).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .conformer import batch_norm, lstm


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def deepspeech2_forward(p, cfg, features, length, training=False):
    """features [B,T,F,1] -> (logits [B,T',V], len').  cfg: dict(conv_kernels, conv_strides, rnn_nlayers, rnn_units, fc_nlayers)."""
    y = features.permute(0, 3, 1, 2)
    length = length.clone()
    for i, (k, s) in enumerate(zip(cfg["conv_kernels"], cfg["conv_strides"])):
        pt, pf = _same_pad(y.shape[2], k[0], s[0]), _same_pad(y.shape[3], k[1], s[1])
        y = F.pad(y, (pf[0], pf[1], pt[0], pt[1]))
        y = F.conv2d(y, p[f"encoder.conv_module.block_{i}.conv2d.kernel"].permute(3, 2, 0, 1), p[f"encoder.conv_module.block_{i}.conv2d.bias"], stride=tuple(s))
        y = batch_norm(y.permute(0, 2, 3, 1), p, f"encoder.conv_module.block_{i}.bn", training)
        y = torch.relu(y).permute(0, 3, 1, 2)
        length = (length + s[0] - 1) // s[0]
    y = y.permute(0, 2, 3, 1)
    B, T, Fq, C = y.shape
    y = y.reshape(B, T, Fq * C)
    for i in range(cfg["rnn_nlayers"]):
        outs = []
        for b in range(B):
            n = int(length[b].clamp(min=1, max=T))
            xb = y[b:b + 1, :n]
            fw, _ = lstm(xb, None, p, f"encoder.rnn_module.block_{i}.blstm.forward_lstm")
            bw, _ = lstm(torch.flip(xb, dims=[1]), None, p, f"encoder.rnn_module.block_{i}.blstm.backward_lstm")
            ob = torch.cat([fw, torch.flip(bw, dims=[1])], dim=-1)
            outs.append(F.pad(ob, (0, 0, 0, T - n)))
        y = torch.cat(outs, dim=0)
    for i in range(cfg["fc_nlayers"]):
        y = torch.relu(y @ p[f"encoder.fc_module.block_{i}.fc.kernel"] + p[f"encoder.fc_module.block_{i}.fc.bias"])
    return y @ p["decoder.vocab.kernel"] + p["decoder.vocab.bias"], length


def ctc_greedy(logits, length, blank=0):
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) -> list of token lists."""
    out = []
    for b in range(logits.shape[0]):
        best = logits[b, : int(length[b])].argmax(-1).tolist()
        toks, prev = [], None
        for k in best:
            if k != prev and k != blank:
                toks.append(k)
            prev = k
        out.append(toks)
    return out


def _sep_conv_module(x, length, p, pre, k, stride, activation, training):
    C = x.shape[-1]
    y = F.pad(x.transpose(1, 2), (k - 1, 0))
    y = F.conv1d(y, p[pre + ".conv.depthwise_kernel"].t().unsqueeze(1), None, stride=stride, groups=C).transpose(1, 2)
    y = y @ p[pre + ".conv.pointwise_kernel"] + p[pre + ".conv.bias"]
    length = (length + stride - 1) // stride
    y = batch_norm(y, p, pre + ".bn", training)
    if activation in ("silu", "swish"):
        y = F.silu(y)
    elif activation == "relu":
        y = torch.relu(y)
    return y, length


def contextnet_forward(p, blocks, features, length, training=False, prefix="encoder."):
    B, T, Fq, C = features.shape
    y = features.reshape(B, T, Fq * C)
    length = length.clone()
    for bi, cfg in enumerate(blocks):
        k, act, stride = cfg["kernel_size"], cfg.get("activation", "silu"), cfg.get("strides", 1)
        inp, inp_len = y, length
        for i in range(cfg["nlayers"]):
            y, length = _sep_conv_module(y, length, p, f"{prefix}block_{bi}.conv_module_{i}", k, stride if i == cfg["nlayers"] - 1 else 1, act, training)
        y, length = _sep_conv_module(y, length, p, f"{prefix}block_{bi}.se.conv_module", k, 1, act, training)
        se = torch.stack([y[b, : max(int(length[b]), 1)].mean(0) if int(length[b]) > 0 else y[b, :1].mean(0) * 0 for b in range(B)])[:, None, :]
        se = F.silu(se @ p[f"{prefix}block_{bi}.se.fc1.kernel"] + p[f"{prefix}block_{bi}.se.fc1.bias"])
        se = torch.sigmoid(se @ p[f"{prefix}block_{bi}.se.fc2.kernel"] + p[f"{prefix}block_{bi}.se.fc2.bias"])
        y = y * se
        if cfg.get("residual", True):
            res, _ = _sep_conv_module(inp, inp_len, p, f"{prefix}block_{bi}.residual", k, stride, "linear", training)
            y = y + res
        y = F.silu(y)
    return y, length
