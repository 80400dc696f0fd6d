"""Module-level fused forward/backward for the ContextNet and DeepSpeech2 encoder bodies on the
hand-written sm_100a kernels: csrc/seqconv.cu (depthwise conv, squeeze-excite, Conv2D lowering), csrc/gemm.cu (pointwise convs,
conv-as-GEMM, dense layers), csrc/conv_bn.cu (sync BatchNorm + activation), csrc/lstm.cu (recurrence, units <= 384).

Same conventions as fused.py: bf16 channels-last activations, fp32 statistics, parameter gradients accumulated by the kernels straight
into the parameters' fp32 `.grad` buffers; the Functions only propagate the activation gradient.

Status: the plain-SIMT kernels used here are verified on the CPU emulator (tests/test_emu_kernels.py, incl. an AddressSanitizer pass);
the composed modules have GPU parity tests (tests/test_gpu_zz_seqconv.py) that were written in a session without GPU time, so the
encoders only take this path when `ops.set_fused_seq(True)` (or TFASR_FUSED_SEQCONV=1) asks for it.
"""
from __future__ import annotations

import torch

from . import kernels
from .fused import PackPlan, WeightPack, _grad, _splits, bn_eval, bn_train_bwd, bn_train_fwd, lstm


# ------------------------------------------------------------------------------------------------ SeparableConv1D + BN + activation
class _SepConvBnFn(torch.autograd.Function):
    """ConvModule.call of ContextNet (contextnet.py:74-86): SeparableConv1D(causal, strides) -> BatchNormalization(sync) -> activation.
        depthwise (seq_dwconv) -> pointwise + bias (tcgen05 GEMM) -> BN statistics / apply (+ swish | relu)
    The pointwise bias sits in front of a batch norm: its gradient is identically zero and is not computed."""

    @staticmethod
    def forward(ctx, x, dw_w, pack: WeightPack, bn, stride, act, training):
        B, T, C = x.shape
        w32 = dw_w.detach().float().contiguous()  # [K, C]
        d = kernels.seq_dwconv_fwd(x, w32, stride)
        To = d.shape[1]
        z = kernels.gemm(d.view(B * To, C), pack.wt, bias=pack.bias_f32())
        gamma, beta, mm, mv = bn
        if training:
            y, saved = bn_train_fwd(z, gamma, beta, mm, mv, act)
        else:
            y, saved = bn_eval(z, gamma, beta, mm, mv, act), None
        ctx.save_for_backward(x, d, z, w32)
        ctx.misc = (dw_w, pack, bn, stride, act, saved, (B, T, To, C))
        return y.view(B, To, -1)

    @staticmethod
    def backward(ctx, dy):
        x, d, z, w32 = ctx.saved_tensors
        dw_w, pack, bn, stride, act, saved, (B, T, To, C) = ctx.misc
        if saved is None:
            raise RuntimeError("ContextNet ConvModule fused backward needs training-mode batch statistics")
        N = z.shape[1]
        dz = bn_train_bwd(dy.contiguous().view(B * To, N), z, saved, bn[0], bn[1], act)
        pack.accumulate_wgrad(d.view(B * To, C), dz)
        dd = kernels.gemm(dz, pack.w)  # [M, C]
        dx = kernels.seq_dwconv_bwd(dd.view(B, To, C), x, w32, _grad(dw_w).view(-1, C), stride, need_dx=ctx.needs_input_grad[0])
        return dx, None, None, None, None, None, None


def sepconv_bn(x, dw_w, pack, bn, stride, act, training):
    """x [B,T,C] bf16 -> [B, ceil(T/stride), N] bf16.  `dw_w` (the depthwise kernel parameter) is an explicit autograd input so that the
    node is recorded even for the first layer, whose input features carry no gradient."""
    return _SepConvBnFn.apply(x, dw_w, pack, bn, int(stride), int(act), bool(training))


# ------------------------------------------------------------------------------------------------ squeeze-excite + residual + activation
class _SeBlockFn(torch.autograd.Function):
    """Tail of ConvBlock.call (contextnet.py:152-164, 258-261) after the SE conv module:
        se = sigmoid(fc2(act(fc1(masked_mean_t(y)))));  out = act(y * se + res)"""

    @staticmethod
    def forward(ctx, y, res, length, W1, b1, W2, b2, act):
        m = kernels.seq_masked_mean(y, length)
        w1, c1, w2, c2 = (t.detach().float().contiguous() for t in (W1, b1, W2, b2))
        z1, s = kernels.se_fc_fwd(m, w1, c1, w2, c2, act)
        out = kernels.se_apply_fwd(y, s, res, act)
        ctx.save_for_backward(y, res, length, m, z1, s, w1, w2)
        ctx.misc = (W1, b1, W2, b2, act)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, res, length, m, z1, s, w1, w2 = ctx.saved_tensors
        W1, b1, W2, b2, act = ctx.misc
        dres, ds = kernels.se_apply_bwd(dout.contiguous(), y, s, res, act)
        dm = kernels.se_fc_bwd(ds, s, z1, m, w1, w2, length, y.shape[1], act, _grad(W1), _grad(b1).view(-1), _grad(W2), _grad(b2).view(-1))
        dy = kernels.se_dy(dres, s, dm, length)
        return dy, (dres if res is not None else None), None, None, None, None, None, None


def se_block(y, res, length, W1, b1, W2, b2, act):
    """y, res [B,T,C] bf16 (res may be None), length [B] int32 -> act(y * se(y) + res)."""
    return _SeBlockFn.apply(y, res, length, W1, b1, W2, b2, int(act))


# ------------------------------------------------------------------------------------------------ Conv2D + BN + activation (DeepSpeech2)
class ConvPack(WeightPack):
    """bf16 operands of a Conv2D kernel [kh, kw, Cin, Cout] lowered to a GEMM: `w` [Kp, Cout] and `wt` [Cout, Kp] with
    Kp = round_up(kh*kw*Cin, 8) and zero rows beyond kh*kw*Cin (TMA row pitch must be a multiple of 16 bytes)."""

    def __init__(self, param, bias, plan: PackPlan):
        kh, kw, cin, cout = param.shape
        kreal = kh * kw * cin
        kp = (kreal + 7) // 8 * 8
        self.param, self.bias, self.shape2d = param, bias, (kreal, cout)
        self.kreal, self.kp = kreal, kp
        self.w = torch.zeros((kp, cout), dtype=torch.bfloat16, device=param.device)
        self.wt = torch.zeros((cout, kp), dtype=torch.bfloat16, device=param.device)
        plan.add(param, kreal, cout, self.w, self.wt, ldw=cout, ldwt=kp)

    def accumulate_wgrad(self, col, dz):
        g = _grad(self.param).view(self.kreal, -1)
        ks = _splits(col.shape[0])
        if self.kp == self.kreal:
            kernels.gemm(col, dz, mn_major=True, out=g, accumulate=True, k_splits=ks)
        else:
            tmp = torch.zeros((self.kp, g.shape[1]), dtype=torch.float32, device=g.device)
            kernels.gemm(col, dz, mn_major=True, out=tmp, accumulate=True, k_splits=ks)
            g.add_(tmp[: self.kreal])


class _Conv2dBnFn(torch.autograd.Function):
    """ConvBlock.call of DeepSpeech2 (deepspeech2.py:105-117): Conv2D(kernel, strides, "same" | causal) -> BatchNormalization -> activation,
    as im2col -> tcgen05 GEMM (+ bias) -> BN kernels.  The column matrix is recomputed in the backward instead of being kept alive."""

    @staticmethod
    def forward(ctx, x, w_param, pack: ConvPack, bn, geom, act, training):
        col = kernels.conv_im2col(x, geom)
        z = kernels.gemm(col, pack.wt, bias=pack.bias_f32())
        del col
        gamma, beta, mm, mv = bn
        if training:
            y, saved = bn_train_fwd(z, gamma, beta, mm, mv, act)
        else:
            y, saved = bn_eval(z, gamma, beta, mm, mv, act), None
        ctx.save_for_backward(x, z)
        ctx.misc = (pack, bn, geom, act, saved)
        return y.view(geom.B, geom.To, geom.Fo, -1)

    @staticmethod
    def backward(ctx, dy):
        x, z = ctx.saved_tensors
        pack, bn, geom, act, saved = ctx.misc
        if saved is None:
            raise RuntimeError("DeepSpeech2 ConvBlock fused backward needs training-mode batch statistics")
        dz = bn_train_bwd(dy.contiguous().view(z.shape), z, saved, bn[0], bn[1], act)
        col = kernels.conv_im2col(x, geom)
        pack.accumulate_wgrad(col, dz)
        del col
        dx = None
        if ctx.needs_input_grad[0]:
            dcol = kernels.gemm(dz, pack.w)  # [M, Kp]
            dx = kernels.conv_col2im(dcol, geom)
        return dx, None, None, None, None, None, None


def conv2d_bn(x, pack, bn, kernel_hw, strides, padding, act, training):
    """x [B,T,F,Cin] (fp32 features or bf16 activations) -> [B,To,Fo,Cout] bf16."""
    B, T, F, Cin = x.shape
    geom = kernels.conv_geom(B, T, F, Cin, kernel_hw[0], kernel_hw[1], strides[0], strides[1], padding)
    return _Conv2dBnFn.apply(x.contiguous(), pack.param, pack, bn, geom, int(act), bool(training))


# ------------------------------------------------------------------------------------------------ Dense + relu + dropout
class _DenseActDropFn(torch.autograd.Function):
    """FcBlock of DeepSpeech2 (deepspeech2.py:262-300): y = Drop(act(x W + b)), act in {relu, None}, all in the GEMM epilogue.
    relu'(z) is read back from the output (y > 0 exactly where z > 0 and the element was kept)."""

    @staticmethod
    def forward(ctx, x, pack: WeightPack, activation, drop_p, salt, seed, anchor=None):
        y = kernels.gemm(x, pack.wt, bias=pack.bias_f32(), activation=activation, dropout_p=drop_p, dropout_seed=seed, dropout_salt=salt)
        ctx.save_for_backward(x, y)
        ctx.misc = (pack, activation, drop_p, salt, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        pack, activation, drop_p, salt, seed = ctx.misc
        dz = dy.contiguous()
        if drop_p > 0.0:
            dz = kernels.scale_mask_colsum(dz, scale=1.0, dropout_p=drop_p, seed=seed, salt=salt)
        if activation == "relu":
            dz = torch.where(y > 0, dz, torch.zeros_like(dz))
        if pack.bias_grad_target() is not None:
            kernels.colsum(dz, pack.bias_grad_target())
        pack.accumulate_wgrad(x, dz)
        dx = kernels.gemm(dz, pack.w) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, None, None


def dense_act_drop(x, pack, activation=None, drop_p=0.0, salt=0, seed=None):
    if activation not in (None, "relu"):
        raise ValueError(activation)
    anchor = pack.param if (not x.requires_grad and torch.is_grad_enabled()) else None
    return _DenseActDropFn.apply(x, pack, activation, float(drop_p), int(salt), seed, anchor)


# ------------------------------------------------------------------------------------------------ bidirectional LSTM
def _reverse_valid(x, length):
    """Per-utterance time reversal of the valid frames: out[b, t] = x[b, len_b - 1 - t] for t < len_b, frames >= len_b stay in place
    (Keras Bidirectional reverses each sequence within its mask; cuDNN packed-sequence semantics)."""
    B, T = x.shape[0], x.shape[1]
    t = torch.arange(T, device=x.device)[None, :]
    ln = length.to(torch.int64).clamp(min=0, max=T)[:, None]
    idx = torch.where(t < ln, ln - 1 - t, t)
    return torch.gather(x, 1, idx[..., None].expand(-1, -1, x.shape[2]))


def bilstm(x, length, fw, bw):
    """Keras Bidirectional(LSTM(return_sequences=True), merge_mode="concat") (deepspeech2.py:186-208) on the cluster LSTM kernels:
    x [B,T,D] bf16, length [B] int32; fw / bw = (kernel [D,4R], recurrent_kernel [R,4R], bias [4R]).  -> [B,T,2R] bf16, zeros at t >= len."""
    yf = lstm(x, length, *fw)
    yb = _reverse_valid(lstm(_reverse_valid(x, length), length, *bw), length)
    return torch.cat([yf, yb], dim=-1)
