"""Optimizer of the training step on ONE kernel launch.

The reference applies Keras Adam variable by variable (`optimizer.apply(gradients, trainable_weights)`, base_model.py:196-215; config
small.yml.j2:73-87).  `FlatAdamW` keeps every parameter, gradient and moment in flat fp32 buffers and updates them with
`tfasr_adamw_flat_f32` (csrc/param_ops.cu): parameters become views of `flat_p`, gradients are the views `dist.GradientReducer`
hands out (so the NCCL all-reduce, the l2-regularizer gradient, the Adam update and the zeroing for the next step all run over the
same contiguous memory).  The learning rate and the step count live on the device: CUDA-graph replays see the schedule.
"""
from __future__ import annotations

import torch

from . import _lib


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        if not params or not (params[0].is_cuda or _lib.EMULATED):
            raise RuntimeError("FlatAdamW needs CUDA parameters (libtfasr_b200 has no CPU path)")
        dev = params[0].device
        if not torch.is_tensor(lr):
            lr = torch.tensor(float(lr), device=dev)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._params = params
        from .dist import flat_layout

        self.offsets, padded = flat_layout(params)
        self.total = sum(p.numel() for p in params)
        self.flat_p = torch.zeros(padded, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + n].view(p.shape)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.flat_g = None
        self.l2_mask = None
        self.l2 = 0.0

    # ---- wiring done by BaseModel.compile
    def attach_grads(self, flat_g: torch.Tensor, params):
        """`flat_g`: the GradientReducer's flat buffer, laid out over the same parameters in the same order."""
        if [id(p) for p in params] != [id(p) for p in self._params]:
            raise ValueError("FlatAdamW and the gradient buffer must be built over the same parameter list")
        if flat_g.numel() != self.flat_p.numel() or flat_g.dtype != torch.float32:
            raise ValueError("gradient buffer must be fp32 with the layout of dist.flat_layout")
        self.flat_g = flat_g

    def set_l2(self, mask: torch.Tensor, l2: float):
        """mask: uint8 [padded], 1 where the Keras l2 regularizer applies (gradient 2 * l2 * w)."""
        self.l2_mask = mask if l2 else None
        self.l2 = float(l2 or 0.0)

    def _own_grads(self):
        self.flat_g = torch.zeros_like(self.flat_p)
        for p, off in zip(self._params, self.offsets):
            view = self.flat_g[off:off + p.numel()].view(p.shape)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, zero_grad: bool = False):
        if self.flat_g is None:
            self._own_grads()
        g = self.param_groups[0]
        lr = g["lr"]
        if not torch.is_tensor(lr):
            lr = g["lr"] = torch.tensor(float(lr), device=self.flat_p.device)
        self.step_t += 1.0
        lib = _lib.load()
        b1, b2 = g["betas"]
        _lib.check(lib.tfasr_adamw_flat_f32(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                            None if self.l2_mask is None else self.l2_mask.data_ptr(), self.flat_p.numel(), lr.data_ptr(),
                                            self.step_t.data_ptr(), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), self.l2,
                                            float(grad_scale), int(bool(zero_grad)), _lib.stream_ptr(self.flat_p.device)), "tfasr_adamw_flat_f32")
        return None

    def zero_grad(self, set_to_none: bool = False):
        if self.flat_g is not None:
            self.flat_g.zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    def state_dict(self):
        return {"flat_p_numel": self.flat_p.numel(), "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_t,
                "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.copy_(sd["step"])
