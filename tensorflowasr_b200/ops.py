"""Op-level dispatch for the Conformer / transducer hot path.

Every function here is one row (or a fused group of rows) of SURVEY.md §8a.  Ops whose sm_100a kernel exists route
through the C ABI (`_lib`); the remaining ones run on torch CUDA library kernels (cuBLAS / cuDNN) and are marked
LIBRARY — they are the replacement queue, not the product.  Nothing here ever runs on the CPU.

`set_precision("bf16" | "fp32")`: "bf16" = bf16 tensor-core operands with fp32 accumulation, fp32 norms / softmax /
loss (>= the reference's mixed_float16 GPU policy, utils/env_util.py:128-132); "fp32" = strict fp32 (TF32 off), used by
the parity tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import _lib

LN_EPS = 1e-3  # keras.layers.LayerNormalization default
BN_EPS = 1e-3  # keras.layers.BatchNormalization default
BN_MOMENTUM = 0.99

_PRECISION = "bf16"
TIMERS = _lib.TIMERS


def set_precision(p: str):
    global _PRECISION
    assert p in ("bf16", "fp32")
    _PRECISION = p
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def precision() -> str:
    return _PRECISION


_FUSED = True


def set_fused(flag: bool):
    """Route the Conformer modules through the hand-written fused kernels (bf16 mode on CUDA only)."""
    global _FUSED
    _FUSED = bool(flag)


def fused_active(x) -> bool:
    return _FUSED and _PRECISION == "bf16" and x.is_cuda


import os as _os

_FUSED_SEQ = _os.environ.get("TFASR_FUSED_SEQCONV", "1") == "1"


def set_fused_seq(flag: bool):
    """Route the ContextNet / DeepSpeech2 encoder bodies through fused_seq.py's kernels.  Default since round 2:
    tests/test_gpu_seqconv.py green on a B200 (kernel level vs torch, both encoders vs the fp32 library route)."""
    global _FUSED_SEQ
    _FUSED_SEQ = bool(flag)


def fused_seq_active(x) -> bool:
    return _FUSED_SEQ and fused_active(x)


_DWCONV_WGRAD_V2 = _os.environ.get("TFASR_DWCONV_WGRAD_V2", "0") == "1"


def set_dwconv_wgrad_v2(flag: bool):
    """ConvModule backward: depthwise weight gradient on the second-generation kernel (csrc/dwconv_wgrad_v2.cu).  Stays OFF: measured on
    a B200 in round 2 it is slower than the first generation at the Conformer-S shape (461 vs 383 us per step over 16 blocks)."""
    global _DWCONV_WGRAD_V2
    _DWCONV_WGRAD_V2 = bool(flag)


def dwconv_wgrad_v2_active() -> bool:
    return _DWCONV_WGRAD_V2


_JOINT_REDUCE_V2 = _os.environ.get("TFASR_JOINT_REDUCE_V2", "0") == "1"


def set_joint_reduce_v2(flag: bool):
    """Training head: de / dp reductions on csrc/joint_v2.cu's kernel.  Stays OFF: measured slower on a B200 (317 vs 162 us)."""
    global _JOINT_REDUCE_V2
    _JOINT_REDUCE_V2 = bool(flag)


def joint_reduce_v2_active() -> bool:
    return _JOINT_REDUCE_V2


_IM2COL_V2 = _os.environ.get("TFASR_IM2COL_V2", "1") == "1"


def set_im2col_v2(flag: bool):
    """Conv2dSubsampling's second conv: im2col / col2im on the second-generation kernels (csrc/subsample_v2.cu).  Default since round 2: 15.51 -> 14.86 ms per
    bench step on a B200 (TFASR_IM2COL_V2=0 selects the first generation)."""
    global _IM2COL_V2
    _IM2COL_V2 = bool(flag)


def im2col_v2_active() -> bool:
    return _IM2COL_V2


_JOINT_V2 = _os.environ.get("TFASR_JOINT_V2", "1") == "1"


def set_joint_v2(flag: bool):
    """Training head: tanh(e + p) on the second-generation kernel (csrc/joint_v2.cu; 77 vs 153 us on a B200).  The de / dp reduction stays
    on the first-generation wide kernel (set_joint_reduce_v2)."""
    global _JOINT_V2
    _JOINT_V2 = bool(flag)


def joint_v2_active() -> bool:
    return _JOINT_V2


_RNNT_GRAD_V2 = _os.environ.get("TFASR_RNNT_GRAD_V2", "1") == "1"


def set_rnnt_grad_v2(flag: bool):
    """Training head: RNN-T gradient on the second-generation kernel (csrc/rnnt_grad_v2.cu).  Default since round 2 (688 vs 739 us
    on a B200)."""
    global _RNNT_GRAD_V2
    _RNNT_GRAD_V2 = bool(flag)


def rnnt_grad_v2_active() -> bool:
    return _RNNT_GRAD_V2


_CONV1_V2 = _os.environ.get("TFASR_CONV1_V2", "1") == "1"


def set_conv1_v2(flag: bool):
    """Conv2dSubsampling's first conv on the second-generation kernels (csrc/conv1.cu).  Default since round 2: forward 195 vs 435 us,
    weight gradient 345 vs 417 us on a B200."""
    global _CONV1_V2
    _CONV1_V2 = bool(flag)


def conv1_v2_active() -> bool:
    return _CONV1_V2


_POS_BATCHED = _os.environ.get("TFASR_POS_BATCHED", "1") == "1"


def set_pos_batched(flag: bool):
    """Conformer encoder: project the relative position table for ALL blocks in one GEMM (fused.PosAllPack; the attention kernels read /
    write column slices through tfasr_rel_attention_*_ld) instead of one projection, one bias column-sum and one weight-gradient GEMM per
    block.  Default since round 2 (15.51 -> 15.23 ms per bench step on a B200)."""
    global _POS_BATCHED
    _POS_BATCHED = bool(flag)


def pos_batched_active() -> bool:
    return _POS_BATCHED


_PEER_BN = _os.environ.get("TFASR_PEER_BN", "1") == "1"


def set_peer_bn(flag: bool):
    """Multi-GPU: exchange the sync-BatchNorm statistics with the one-launch all-reduce over peer memory (csrc/peer_allreduce.cu,
    dist.PeerAllReduce) instead of one NCCL all-reduce per BatchNorm pass.  Default since round 2: on two B200s the
    exchange takes 5.2 us against NCCL's 10.6 (36 of them per step), sums bit-identical on every rank (tests/test_gpu_peer_multi.py);
    ranks on different nodes fall back to NCCL (dist.small_all_reduce)."""
    global _PEER_BN
    _PEER_BN = bool(flag)


def peer_bn_active() -> bool:
    return _PEER_BN


_FUSED_LSE = _os.environ.get("TFASR_FUSED_LSE", "1") == "1"


def set_fused_lse(flag: bool):
    """Training head: take the RNN-T log-softmax normaliser from softmax statistics computed in the joint GEMM's epilogue (gemm feature
    class 8 + csrc/rnnt_combine.cu) instead of a separate pass over the 2 GB of logits (first step).  Default since round 2: tests/test_gpu_lse.py green
    on a B200, RNN-T group 0.50 -> 0.71 of the HBM roofline on algorithmic bytes."""
    global _FUSED_LSE
    _FUSED_LSE = bool(flag)


def fused_lse_active() -> bool:
    return _FUSED_LSE


_BF16_LOGITS = _os.environ.get("TFASR_BF16_LOGITS", "1") == "1"


def set_bf16_logits(flag: bool):
    """Training head (needs fused_lse): the joint GEMM stores the [B,T,U+1,V] logits as bf16 — statistics taken over the rounded values, so
    the loss and the gradient are exact for the stored logits — and csrc/rnnt_grad_v2.cu (bf16_io) replaces them by their gradient in
    place: 4 B per lattice node and vocabulary entry instead of 8 (second step).  Rows are pitched at a multiple of 64
    elements.  Default since round 2: parity tests green with it on a B200 (config-3 size: loss within 2.1e-3 of the float64 oracle,
    gradient cosines >= 0.999), step 12.65 -> 12.43 ms."""
    global _BF16_LOGITS
    _BF16_LOGITS = bool(flag)


def bf16_logits_active() -> bool:
    return _BF16_LOGITS and _FUSED_LSE


_LATTICE_V2 = _os.environ.get("TFASR_LATTICE_V2", "1") == "1"


def set_lattice_v2(flag: bool):
    """Training head (needs fused_lse): alpha / beta wavefronts with the utterance's transition planes staged in shared memory
    (csrc/rnnt_lattice_v2.cu) when the lattice fits (8 T (U+1) bytes <= 227 KB), else the first-generation kernel.  Default since round 2:
    78 us against 150 us at T' = 250, U = 64, costs equal to 1e-5."""
    global _LATTICE_V2
    _LATTICE_V2 = bool(flag)


def lattice_v2_active() -> bool:
    return _LATTICE_V2 and _FUSED_LSE


_PERSISTENT_DECODE = _os.environ.get("TFASR_PERSISTENT_DECODE", "1") == "1"


def set_persistent_decode(flag: bool):
    """Greedy decoding through the one-launch persistent kernel (csrc/decode.cu) instead of the host-driven loop.  Default since
    round 2: tokens bit-exact vs the host loop and the oracle on a B200, 54.5 vs 750.7 us per decode iteration at batch 32."""
    global _PERSISTENT_DECODE
    _PERSISTENT_DECODE = bool(flag)


def persistent_decode_active() -> bool:
    return _PERSISTENT_DECODE


_FUSED_FFN = _os.environ.get("TFASR_FUSED_FFN", "1") == "1"


def set_fused_ffn(flag: bool):
    """Conformer FFModule forward as ONE kernel (csrc/ffn_fused.cu: LayerNorm prologue, both GEMMs, the hidden activation handed over in
    SMEM) instead of LayerNorm + two GEMM launches."""
    global _FUSED_FFN
    _FUSED_FFN = bool(flag)


def fused_ffn_active() -> bool:
    return _FUSED_FFN


_ATTENTION_TC = _os.environ.get("TFASR_ATTENTION_TC", "auto")  # "1" always, "0" never, "auto": where the mma.sync generation has no kernel


def set_attention_tc(flag):
    """Relative-position attention forward on tcgen05 (csrc/attention_tc.cu: TMA-fed operands, TMEM accumulators, any sequence length).
    True / False force the route; "auto" (default) takes it for T' > 256 — the Conformer-L streaming decode of BASELINE configs[4], which
    the mma.sync generation (csrc/attention.cu: whole-utterance SMEM staging, T' <= 256) cannot run — and keeps the mma.sync kernels at
    T' <= 256, where they measured faster on a B200 in round 2 (39.6 vs 59.1 us per layer at B = 32, T' = 250: the softmax warps of the
    tcgen05 kernel are its critical path)."""
    global _ATTENTION_TC
    _ATTENTION_TC = flag if flag == "auto" else ("1" if flag else "0")


def attention_tc_active(T: int = 0) -> bool:
    if _ATTENTION_TC == "auto":
        return T > 256
    return _ATTENTION_TC == "1"


_SWITCHES = ("_FUSED_FFN", "_FUSED", "_FUSED_SEQ", "_DWCONV_WGRAD_V2", "_JOINT_REDUCE_V2", "_IM2COL_V2", "_JOINT_V2", "_RNNT_GRAD_V2", "_CONV1_V2", "_POS_BATCHED",
             "_PEER_BN", "_FUSED_LSE", "_PERSISTENT_DECODE", "_BF16_LOGITS", "_LATTICE_V2")


def switch_state() -> dict:
    """Snapshot of every kernel-route switch (tests flip them and put them back with restore_switches)."""
    g = globals()
    st = {k: g[k] for k in _SWITCHES}
    st["_ATTENTION_TC"] = _ATTENTION_TC
    return st


def restore_switches(state: dict):
    global _ATTENTION_TC
    globals().update({k: bool(v) for k, v in state.items() if k in _SWITCHES})
    if "_ATTENTION_TC" in state:
        _ATTENTION_TC = state["_ATTENTION_TC"]


def compute_dtype():
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


# ------------------------------------------------------------------------------------------------ small ops
def layer_norm(x, gamma, beta):
    """keras LayerNormalization(eps=1e-3) over the last axis; statistics in fp32.  LIBRARY"""
    y = F.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), LN_EPS)
    return y.to(compute_dtype())


def linear(x, kernel, bias=None, activation=None):
    """keras Dense / 1x1 Conv1D / EinsumDense: x @ kernel[K, N] + bias, optional swish.  LIBRARY (cuBLAS)"""
    dt = compute_dtype()
    y = F.linear(x.to(dt), kernel.to(dt).t(), None if bias is None else bias.to(dt))
    if activation == "swish":
        y = F.silu(y)
    elif activation == "relu":
        y = F.relu(y)
    elif activation == "tanh":
        y = torch.tanh(y)
    elif activation is not None:
        raise ValueError(activation)
    return y


def dropout(x, rate, training):
    if training and rate and rate > 0.0:
        return F.dropout(x, rate, True)
    return x


def residual(x, y, factor):
    """models/layers/residual.py:58-62."""
    return x + factor * y


_relpe_cache = {}


def relative_position_encoding(T, length, dmodel, device, memory_length=0):
    """RelativeSinusoidalPositionalEncoding.call (non-causal, interleaved), positional_encoding.py:114-174 -> [B, 2T+M-1, D].

    The table over positions [T+M-1 .. -(T-1)] is cached; the per-utterance roll by -(T - len_b) and the zero mask beyond
    2 len_b + M - 1 are one gather + one where."""
    key = (T, dmodel, str(device), memory_length)
    pe = _relpe_cache.get(key)
    if pe is None:
        pos = torch.cat([torch.arange(T + memory_length - 1, 0, -1), torch.arange(0, -T, -1)]).to(torch.float32)
        d = torch.arange(dmodel)
        timescales = torch.pow(torch.tensor(1.0 / 10000.0), (2 * (d // 2)).float() / dmodel)
        ang = pos[:, None] * timescales[None, :]
        pe = torch.where((d % 2 == 0)[None, :], torch.sin(ang), torch.cos(ang)).to(device)
        _relpe_cache[key] = pe
    R = pe.shape[0]
    length = length.to(device=device, dtype=torch.int64)
    idx = (torch.arange(R, device=device)[None, :] + (T - length)[:, None]) % R
    out = pe[idx]  # [B, R, D]
    valid = torch.arange(R, device=device)[None, :] < (2 * length + memory_length - 1)[:, None]
    return (out * valid[..., None]).to(compute_dtype())


def streaming_mask(chunk_size, history_size, T, device):
    """compute_streaming_mask, multihead_attention.py:104-143 (pinned by the reference's tests/test_mask.py)."""
    q = torch.arange(T, device=device) // chunk_size
    hist = T if history_size < 0 else history_size
    start = torch.clamp(q * chunk_size - hist, min=0)
    end = torch.clamp(q * chunk_size + chunk_size, max=T)
    k = torch.arange(T, device=device)[None, :]
    return (k < end[:, None]) & (k >= start[:, None])


def rel_attention(q, k, v, pos, ubias, vbias, length, N, H, chunk_size=None, history_size=None):
    """MultiHeadRelativeAttention._compute_attention, multihead_attention.py:543-582.  q,k,v [B,T,N*H]; pos [B,2T-1,N*H].
    rel_left_shift + slice is the index map out[t, j] = posattn[t, T-1-t+j].  Query-side mask only.  LIBRARY (to be fused)"""
    B, T, _ = q.shape
    R = pos.shape[1]
    dt = compute_dtype()
    q4 = q.view(B, T, N, H)
    k4 = k.view(B, T, N, H)
    v4 = v.view(B, T, N, H)
    p4 = pos.view(B, R, N, H)
    scale = 1.0 / math.sqrt(H)
    qu = ((q4 + ubias.to(dt)) * scale).permute(0, 2, 1, 3)  # [B,N,T,H]
    qv = ((q4 + vbias.to(dt)) * scale).permute(0, 2, 1, 3)
    content = torch.matmul(qu, k4.permute(0, 2, 3, 1))  # [B,N,T,S]
    positional = torch.matmul(qv, p4.permute(0, 2, 3, 1))  # [B,N,T,R]
    idx = (T - 1 - torch.arange(T, device=q.device)[:, None] + torch.arange(T, device=q.device)[None, :])
    positional = torch.gather(positional, 3, idx.expand(B, N, T, T))
    scores = (content + positional).float()
    mask = (torch.arange(T, device=q.device)[None, :] < length.to(q.device)[:, None])[:, None, :, None]
    if chunk_size is not None and history_size is not None:
        mask = mask & streaming_mask(chunk_size, history_size, T, q.device)[None, None]
    scores = scores.masked_fill(~mask, -1e9)
    attn = torch.softmax(scores, dim=-1).to(dt)
    ctx = torch.matmul(attn, v4.permute(0, 2, 1, 3))  # [B,N,T,H]
    return ctx.permute(0, 2, 1, 3).reshape(B, T, N * H)


# ------------------------------------------------------------------------------------------------ batch norm (sync)
class _SyncBatchNormFn(torch.autograd.Function):
    """Training-mode BatchNormalization(synchronized=True): statistics over every row of every replica (C2 in SURVEY.md §2.2).
    One small all-reduce forward (sum, sum of squares), one backward (sum dy, sum dy*xhat)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        C = x.shape[-1]
        xf = x.reshape(-1, C).float()
        stats = torch.stack([xf.sum(0), (xf * xf).sum(0)])
        count = torch.full((1,), float(xf.shape[0]), device=x.device)  # fill kernel (graph-capturable), not an H2D copy
        world = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        if world:
            packed = torch.cat([stats.reshape(-1), count])
            torch.distributed.all_reduce(packed)
            stats, count = packed[:-1].reshape(2, C), packed[-1:]
        mean = stats[0] / count
        var = (stats[1] / count - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + eps)
        xhat = (xf - mean) * rstd
        y = xhat * gamma.float() + beta.float()
        ctx.save_for_backward(xhat, rstd, gamma)
        ctx.count = count
        ctx.world = world
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        ctx.mark_non_differentiable(mean, var)
        return y.reshape(x.shape).to(x.dtype), mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        xhat, rstd, gamma = ctx.saved_tensors
        C = xhat.shape[-1]
        dyf = dy.reshape(-1, C).float()
        sum_dy = dyf.sum(0)
        sum_dy_xhat = (dyf * xhat).sum(0)
        red = torch.stack([sum_dy, sum_dy_xhat])
        if ctx.world:
            red = red.clone()
            torch.distributed.all_reduce(red)
        n = ctx.count
        dx = (gamma.float() * rstd) * (dyf - red[0] / n - xhat * (red[1] / n))
        return dx.reshape(ctx.shape).to(ctx.dtype), sum_dy_xhat.to(gamma.dtype), sum_dy.to(gamma.dtype), None


def batch_norm(x, gamma, beta, moving_mean, moving_var, training):
    """keras BatchNormalization(eps=1e-3, momentum=0.99, synchronized=True) over the last axis.  LIBRARY (elementwise)"""
    if training:
        y, mean, var = _SyncBatchNormFn.apply(x, gamma, beta, BN_EPS)
        with torch.no_grad():
            moving_mean.mul_(BN_MOMENTUM).add_(mean, alpha=1.0 - BN_MOMENTUM)
            moving_var.mul_(BN_MOMENTUM).add_(var, alpha=1.0 - BN_MOMENTUM)
        return y
    scale = gamma.float() * torch.rsqrt(moving_var.float() + BN_EPS)
    shift = beta.float() - moving_mean.float() * scale
    return (x.float() * scale + shift).to(x.dtype)


def _activation(x, name):
    if name == "swish":
        return F.silu(x)
    if name == "relu":
        return F.relu(x)
    if name in (None, "linear"):
        return x
    raise ValueError(name)


def conv2d_block(x, kernel, bias, stride, padding, norm, activation, training, norm_params):
    """One Conv2dSubsampling sub-block (subsampling.py:218-230 + convolution.py:132-144): causal pad -> VALID conv -> BN|LN ->
    activation.  x [B,T,F,Cin] NHWC; kernel [kh,kw,Cin,Cout].  LIBRARY (cuDNN)"""
    dt = compute_dtype()
    kh, kw = kernel.shape[0], kernel.shape[1]
    y = x.to(dt).permute(0, 3, 1, 2)
    if padding == "causal":
        y = F.pad(y, (kw - 1, 0, kh - 1, 0))
    elif padding == "same":
        raise NotImplementedError("'same' padding is outside the hot path")
    w = kernel.to(dt).permute(3, 2, 0, 1)
    y = F.conv2d(y, w, bias.to(dt), stride=stride)
    y = y.permute(0, 2, 3, 1)  # NHWC
    if norm == "batch":
        y = batch_norm(y, *norm_params, training)
    elif norm == "layer":
        y = layer_norm(y, *norm_params)
    return _activation(y, activation)


def glu_dwconv_norm_swish(y, dw_kernel, dw_bias, norm, training):
    """GLU -> causal depthwise conv (K taps) -> BN|LN -> swish (conformer.py:369-372).  y [B,T,2D]; dw_kernel [K,D].  LIBRARY"""
    dt = compute_dtype()
    a, b = y.chunk(2, dim=-1)
    g = a * torch.sigmoid(b)
    K, D = dw_kernel.shape
    gp = F.pad(g.transpose(1, 2), (K - 1, 0))
    c = F.conv1d(gp, dw_kernel.to(dt).t().unsqueeze(1), dw_bias.to(dt), groups=D).transpose(1, 2)
    if norm[0] == "batch":
        c = batch_norm(c, norm[1], norm[2], norm[3], norm[4], training)
    else:
        c = layer_norm(c, norm[1], norm[2])
    return F.silu(c)


# ------------------------------------------------------------------------------------------------ transducer head
def joint_logits(enc_proj, pred_proj, vocab_kernel, vocab_bias):
    """TransducerJoint.call after the two pre-joint Dense layers (base_transducer.py:199-207,291-292):
    logits[b,t,u,:] = tanh(e[b,t] + p[b,u]) @ Wv + bv, returned in fp32 [B,T,U1,V].  LIBRARY (cuBLAS)"""
    dt = compute_dtype()
    h = torch.tanh(enc_proj.to(dt)[:, :, None, :] + pred_proj.to(dt)[:, None, :, :])
    logits = F.linear(h, vocab_kernel.to(dt).t(), vocab_bias.to(dt))
    return logits.float()


class _TransducerLossHeadFn(torch.autograd.Function):
    """Joint network + RNN-T loss for training, with the [B,T,U1,V] logits and their gradient living only inside this node:
        h = tanh(e + p) (bf16) -> logits = h Wv + bv (fp32) -> K1/K2 (costs) ; backward: K3 writes d logits IN PLACE over the
        logits buffer -> dh = dlogits Wv^T, dWv = h^T dlogits, dbv = sum dlogits, de / dp = reductions of dh * (1 - h^2).
    """

    @staticmethod
    def forward(ctx, enc_proj, pred_proj, vocab_kernel, vocab_bias, labels, logit_length, label_length):
        lib = _lib.load()
        dt = compute_dtype()
        B, T, J = enc_proj.shape
        U1 = pred_proj.shape[1]
        V = vocab_kernel.shape[1]
        h = torch.tanh(enc_proj.to(dt)[:, :, None, :] + pred_proj.to(dt)[:, None, :, :])  # [B,T,U1,J]
        wv = vocab_kernel.to(dt)
        if dt == torch.bfloat16:
            try:
                logits = torch.mm(h.view(-1, J), wv, out_dtype=torch.float32)
            except TypeError:
                logits = torch.mm(h.view(-1, J), wv).float()
            logits = logits.add_(vocab_bias.float()).view(B, T, U1, V)
        else:
            logits = torch.addmm(vocab_bias.float(), h.view(-1, J), wv).view(B, T, U1, V)
        ws_bytes = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
        costs = torch.empty(B, dtype=torch.float32, device=logits.device)
        with TIMERS.region("rnnt_fwd"):
            _lib.check(
                lib.tfasr_rnnt_loss_fwd_f32(_lib.ptr(logits), _lib.ptr(labels), _lib.ptr(logit_length), _lib.ptr(label_length), B, T, U1, V,
                                            0, _lib.ptr(costs), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(logits.device)),
                "tfasr_rnnt_loss_fwd_f32",
            )
        ctx.save_for_backward(h, wv, logits, labels, logit_length, label_length, costs, ws)
        ctx.dtypes = (enc_proj.dtype, pred_proj.dtype, vocab_kernel.dtype, vocab_bias.dtype)
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        lib = _lib.load()
        h, wv, logits, labels, logit_length, label_length, costs, ws = ctx.saved_tensors
        B, T, U1, V = logits.shape
        J = h.shape[-1]
        scale = grad_costs.float().contiguous()
        with TIMERS.region("rnnt_bwd"):
            _lib.check(
                lib.tfasr_rnnt_loss_bwd_f32(_lib.ptr(logits), _lib.ptr(labels), _lib.ptr(logit_length), _lib.ptr(label_length), B, T, U1, V,
                                            0, _lib.ptr(costs), _lib.ptr(scale), _lib.ptr(logits), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr(logits.device)),
                "tfasr_rnnt_loss_bwd_f32",
            )
        dl32 = logits.view(-1, V)  # now holds d costs / d logits (scaled)
        dbias = dl32.sum(0)
        dl = dl32.to(h.dtype)
        h2 = h.view(-1, J)
        dwv = torch.mm(h2.t(), dl)  # [J, V]
        dh = torch.mm(dl, wv.t())  # [M, J]
        dz = (dh.float() * (1.0 - h2.float() * h2.float())).view(B, T, U1, J)
        de = dz.sum(2)
        dp = dz.sum(1)
        d0, d1, d2, d3 = ctx.dtypes
        return de.to(d0), dp.to(d1), dwv.to(d2), dbias.to(d3), None, None, None


def transducer_loss_head(enc_proj, pred_proj, vocab_kernel, vocab_bias, labels, logit_length, label_length):
    """-> costs [B] f32 (per-utterance RNN-T negative log-likelihood); labels/lengths int32 CUDA tensors."""
    return _TransducerLossHeadFn.apply(enc_proj, pred_proj, vocab_kernel, vocab_bias, labels, logit_length, label_length)


# ------------------------------------------------------------------------------------------------ prediction net
def embedding(tokens, table):
    """models/layers/embedding.py:41-44.  LIBRARY (gather)"""
    return F.embedding(tokens.long(), table).to(compute_dtype())


def lstm(x, length, kernel, recurrent_kernel, bias, state=None):
    """keras LSTM (gate order i,f,c,o — the same as cuDNN's i,f,g,o; one bias), base_transducer.py:74-88.
    x [B,U,E] -> ([B,U,R], (h, c)).  Masked steps (t >= length) output zeros (zero_output_for_mask=True); the LSTM is causal, so
    earlier outputs do not depend on the mask.  The returned state is the state after the LAST step of the padded sequence when
    `length` is given (training never reads it; decoding steps one token at a time with length=None).  LIBRARY (cuDNN RNN, fp32)"""
    B, U, _ = x.shape
    R = recurrent_kernel.shape[0]
    w_ih = kernel.float().t().contiguous()
    w_hh = recurrent_kernel.float().t().contiguous()
    b_ih = bias.float()
    b_hh = torch.zeros_like(b_ih)
    if state is None:
        h0 = x.new_zeros(1, B, R, dtype=torch.float32)
        c0 = x.new_zeros(1, B, R, dtype=torch.float32)
    else:
        h0, c0 = state[0].float().reshape(1, B, R).contiguous(), state[1].float().reshape(1, B, R).contiguous()
    train = torch.is_grad_enabled() and (x.requires_grad or kernel.requires_grad)  # cuDNN keeps its reserve space only in train mode
    y, h, c = torch._VF.lstm(x.float().contiguous(), (h0, c0), [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, train, False, True)
    if length is not None:
        mask = torch.arange(U, device=x.device)[None, :] < length.to(x.device)[:, None]
        y = y * mask[..., None]
    return y.to(compute_dtype()), (h[0], c[0])
