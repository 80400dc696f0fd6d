"""RnntLoss / CtcLoss / rnnt_loss — host-side mirror of the reference's loss API over the sm_100a kernels.

Mirrors:
  * `BaseLoss`            tensorflow_asr/losses/base_loss.py:21-42  (blank must be 0; int32 casts; the
                          `logit_length = max(logit_length, label_length)` clamp at :36)
  * `RnntLoss`            tensorflow_asr/losses/rnnt_loss.py:30-66
  * `rnnt_loss(...)`      tensorflow_asr/losses/impl/rnnt.py:11-35 (functional form; same keyword names)
  * `CtcLoss`             tensorflow_asr/losses/ctc_loss.py:42-66
Loss tensors are per-utterance vectors [B]; `reduction="sum_over_batch_size"` (Keras default) means the mean over B.
The gradient w.r.t. the logits is produced by the library (like the reference's custom_gradient / warp-transducer op)
and scaled by the upstream gradient inside the kernel (impl/rnnt.py:321-324).  No CPU fallback exists.
"""
from __future__ import annotations

import os

import torch

from . import _lib, schemas

TFASR_USE_CPU_LOSS = os.getenv("TFASR_USE_CPU_LOSS", "False") in ("true", "True", "1")  # kept for API parity; ignored
TFASR_USE_TF_CTC = os.getenv("TFASR_USE_TF_CTC", "False") in ("true", "True", "1")  # kept for API parity; ignored


class _RnntLossFn(torch.autograd.Function):
    """forward: K1 (lse) + K2 (alpha/beta wavefront); backward: K3 (grad), written in place over the logits buffer when
    `inplace_grad` (the logits are not needed by any other backward: the joint's vocab GEMM saves its input, not its output)."""

    @staticmethod
    def forward(ctx, logits, labels, logit_length, label_length, blank, inplace_grad):
        lib = _lib.load()
        B, T, U1, V = logits.shape
        if labels.shape != (B, U1 - 1):
            raise ValueError(f"rnnt_loss: labels must be [B, U] = [{B}, {U1 - 1}], got {tuple(labels.shape)}")
        logits_c = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.to(torch.float32).contiguous()
        ws_bytes = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
        costs = torch.empty(B, dtype=torch.float32, device=logits.device)
        _lib.check(
            lib.tfasr_rnnt_loss_fwd_f32(_lib.ptr(logits_c), _lib.ptr(labels), _lib.ptr(logit_length), _lib.ptr(label_length),
                                        B, T, U1, V, int(blank), _lib.ptr(costs), _lib.ptr(ws), ws_bytes,
                                        _lib.stream_ptr(logits.device)),
            "tfasr_rnnt_loss_fwd_f32",
        )
        ctx.save_for_backward(logits_c, labels, logit_length, label_length, costs, ws)
        ctx.blank = int(blank)
        ctx.inplace = bool(inplace_grad) and logits_c.data_ptr() == logits.data_ptr()
        ctx.in_dtype = logits.dtype
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        lib = _lib.load()
        logits, labels, logit_length, label_length, costs, ws = ctx.saved_tensors
        B, T, U1, V = logits.shape
        grads = logits if ctx.inplace else torch.empty_like(logits)
        scale = grad_costs.to(torch.float32).contiguous()
        _lib.check(
            lib.tfasr_rnnt_loss_bwd_f32(_lib.ptr(logits), _lib.ptr(labels), _lib.ptr(logit_length), _lib.ptr(label_length),
                                        B, T, U1, V, ctx.blank, _lib.ptr(costs), _lib.ptr(scale), _lib.ptr(grads), _lib.ptr(ws),
                                        ws.numel(), _lib.stream_ptr(logits.device)),
            "tfasr_rnnt_loss_bwd_f32",
        )
        if ctx.in_dtype != torch.float32:
            grads = grads.to(ctx.in_dtype)
        return grads, None, None, None, None, None


def rnnt_loss(logits, logits_length, labels, labels_length, blank=0, name=None, use_cpu=False, output_shapes=None,
              inplace_grad=False):
    """Functional RNN-T loss, keyword-compatible with tensorflow_asr/losses/impl/rnnt.py:11-20.

    logits [B, T, U+1, V] raw activations; labels [B, U]; returns costs [B] (f32; cast back for f16/bf16 inputs).
    `use_cpu` / `output_shapes` are accepted for signature parity and ignored (there is no CPU path).
    """
    _lib.require_cuda(logits, "rnnt_loss")
    assert blank == 0, "Only support blank=0"
    dev = logits.device
    labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    logits_length = logits_length.to(device=dev, dtype=torch.int32).contiguous()
    labels_length = labels_length.to(device=dev, dtype=torch.int32).contiguous()
    costs = _RnntLossFn.apply(logits, labels, logits_length, labels_length, blank, inplace_grad)
    if logits.dtype in (torch.float16, torch.bfloat16):
        costs = costs.to(logits.dtype)  # impl/rnnt.py:56-57, 328-329
    return costs


class _CtcLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, logit_length, label_length, blank):
        lib = _lib.load()
        B, T, V = logits.shape
        L = labels.shape[1]
        logits_c = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.to(torch.float32).contiguous()
        ws_bytes = lib.tfasr_ctc_workspace_bytes(B, T, L, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
        costs = torch.empty(B, dtype=torch.float32, device=logits.device)
        need_grad = logits.requires_grad
        grads = torch.empty_like(logits_c) if need_grad else None
        _lib.check(
            lib.tfasr_ctc_loss_f32(_lib.ptr(logits_c), _lib.ptr(labels), _lib.ptr(logit_length), _lib.ptr(label_length), B, T, L, V,
                                   int(blank), _lib.ptr(costs), _lib.ptr(grads), None, _lib.ptr(ws), ws_bytes,
                                   _lib.stream_ptr(logits.device)),
            "tfasr_ctc_loss_f32",
        )
        ctx.save_for_backward(grads)
        ctx.in_dtype = logits.dtype
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        (grads,) = ctx.saved_tensors
        g = grads * grad_costs.to(torch.float32)[:, None, None]
        return g.to(ctx.in_dtype), None, None, None, None


def ctc_loss(logits, logits_length, labels, labels_length, blank=0, name=None):
    """tf.nn.ctc_loss-shaped functional form (batch-major logits [B, T, V], dense labels [B, L]) -> costs [B]."""
    _lib.require_cuda(logits, "ctc_loss")
    assert blank == 0, "Only support blank=0"
    dev = logits.device
    labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    logits_length = logits_length.to(device=dev, dtype=torch.int32).contiguous()
    labels_length = labels_length.to(device=dev, dtype=torch.int32).contiguous()
    return _CtcLossFn.apply(logits, labels, logits_length, labels_length, blank)


class BaseLoss(torch.nn.Module):
    """tensorflow_asr/losses/base_loss.py:21-42."""

    def __init__(self, blank=0, reduction="sum_over_batch_size", name=None):
        super().__init__()
        assert blank == 0, "Only support blank=0"
        assert reduction in ("sum_over_batch_size", "sum", "none", None)
        self.blank = blank
        self.reduction = reduction
        self.name = name

    def call(self, y_true: schemas.TrainLabel, y_pred: schemas.TrainOutput):
        logit_length = y_pred.logits_length.to(torch.int32)
        labels = y_true.labels.to(torch.int32)
        label_length = y_true.labels_length.to(torch.int32)
        logit_length = torch.where(logit_length < label_length, label_length, logit_length)  # base_loss.py:36
        return y_pred.logits, logit_length, labels, label_length

    def _reduce(self, losses):
        if self.reduction == "sum_over_batch_size":
            return losses.mean()
        if self.reduction == "sum":
            return losses.sum()
        return losses

    def forward(self, y_true, y_pred):
        """Keras `Loss.__call__`: per-utterance losses from `call`, then the reduction."""
        return self._reduce(self.call(y_true, y_pred))

    def get_config(self):
        return {"blank": self.blank, "reduction": self.reduction, "name": self.name}


class RnntLoss(BaseLoss):
    """tensorflow_asr/losses/rnnt_loss.py:30-66."""

    def __init__(self, blank, reduction="sum_over_batch_size", output_shapes=None, name=None, inplace_grad=False):
        super().__init__(blank=blank, reduction=reduction, name=name)
        self.use_cpu = False  # no CPU path in the B200 build
        self.output_shapes = output_shapes
        self.inplace_grad = inplace_grad

    def call(self, y_true, y_pred):
        logits, logit_length, labels, label_length = super().call(y_true, y_pred)
        return rnnt_loss(logits=logits, logits_length=logit_length, labels=labels, labels_length=label_length, blank=self.blank,
                         name=self.name, use_cpu=self.use_cpu, output_shapes=self.output_shapes, inplace_grad=self.inplace_grad)

    def get_config(self):
        conf = super().get_config()
        conf.update({"output_shapes": self.output_shapes})
        return conf


class CtcLoss(BaseLoss):
    """tensorflow_asr/losses/ctc_loss.py:42-66."""

    def __init__(self, blank=0, reduction="sum_over_batch_size", name=None):
        super().__init__(blank=blank, reduction=reduction, name=name)

    def call(self, y_true, y_pred):
        logits, logit_length, labels, label_length = super().call(y_true, y_pred)
        return ctc_loss(logits=logits, logits_length=logit_length, labels=labels, labels_length=label_length, blank=self.blank,
                        name=self.name)
