"""Module-level fused forward/backward for the Conformer block on the hand-written sm_100a kernels.

Each autograd.Function below is one reference module (FFModule / MHSAModule / ConvModule, models/encoders/conformer.py)
executed as a short chain of tcgen05 GEMMs with fused epilogues and row-wise fused kernels, bf16 activations, fp32
accumulation and statistics.  Parameter gradients are accumulated by the kernels straight into the parameters' fp32 `.grad`
buffers (split-K atomics), so the Functions only propagate the activation gradient.
"""
from __future__ import annotations

import torch

from . import dist, kernels

LN_EPS = 1e-3


class StepState:
    """Per-step device state shared by all fused modules: the dropout / SpecAugment counter (bumped once per step, inside the CUDA graph).
    Its start value mixes a user seed with the replica's rank (`reseed`; every replica draws its own masks, a restarted run can draw new ones)
    and it is part of the training state (`state_dict` / `load_state_dict`: BaseModel saves it beside the optimizer)."""

    def __init__(self):
        self.seed = None
        self.initial = 0

    @staticmethod
    def _mix(seed: int, rank: int) -> int:
        x = (int(seed) * 0x9E3779B97F4A7C15 + (int(rank) + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
        return x & 0x7FFFFFFFFFFFFFFF if seed or rank else 0  # (0, rank 0) keeps the historical stream

    def reseed(self, seed: int, rank: int = None):
        if rank is None:
            d = torch.distributed
            rank = d.get_rank() if d.is_available() and d.is_initialized() else 0
        self.initial = self._mix(seed, rank)
        if self.seed is not None:
            self.seed.fill_(self.initial)

    def seed_tensor(self, device):
        if self.seed is None or self.seed.device != device:
            self.seed = torch.full((1,), self.initial, dtype=torch.int64, device=device)
        return self.seed

    def bump(self):
        if self.seed is not None:
            self.seed.add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)

    def state_dict(self) -> dict:
        return {"counter": int(self.seed.item()) if self.seed is not None else self.initial}

    def load_state_dict(self, state: dict):
        self.initial = int(state["counter"])
        if self.seed is not None:
            self.seed.fill_(self.initial)


STATE = StepState()


class SideStream:
    """Weight-gradient GEMMs (and other work nobody on the backward chain waits for) run on a second stream: inside the captured step graph
    that makes them parallel branches, so the activation-gradient chain — ~1000 dependent, latency-bound launches — no longer queues behind
    them.  `run(fn, *tensors)` orders the side stream after everything issued so far, calls fn on it and keeps `tensors` (fn's inputs) alive
    until `join()`, which the end of every backward pass performs (autograd engine callback) and which UnpackPlan.run / BaseModel call before
    they touch the gradients."""

    def __init__(self):
        import os

        self.enabled = os.environ.get("TFASR_SIDE_WGRAD", "1") == "1"
        self.stream = None
        self.keep = []
        self.dirty = False
        self._queued = False

    def active(self, t) -> bool:
        from . import _lib

        return self.enabled and t.is_cuda and not _lib.EMULATED and not _lib.TIMERS.on

    def branch_stream(self, t):
        """A second side stream for an independent branch of the step (the prediction network), or None when side streams are off."""
        import os

        if not self.active(t) or os.environ.get("TFASR_SIDE_PRED", "1") != "1":
            return None
        if getattr(self, "_branch", None) is None or self._branch.device != t.device:
            self._branch = torch.cuda.Stream(device=t.device)
        return self._branch

    def run(self, fn, *tensors):
        if not tensors or not self.active(tensors[0]):
            fn()
            return
        main = torch.cuda.current_stream()
        if self.stream is None or self.stream.device != tensors[0].device:
            self.stream = torch.cuda.Stream(device=tensors[0].device)
        self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(tensors)
        self.dirty = True
        if not self._queued:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)
                self._queued = True
            except RuntimeError:  # not inside a backward pass: the caller joins
                pass

    def _end_of_backward(self):
        self._queued = False
        self.join()

    def join(self):
        if self.dirty:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.keep.clear()
            self.dirty = False


SIDE = SideStream()


def _wgrad(x, dz, out, k_splits):
    """out (fp32, a parameter's gradient or a packed gradient buffer) += x^T dz on the side stream."""
    SIDE.run(lambda: kernels.gemm(x, dz, mn_major=True, out=out, accumulate=True, k_splits=k_splits), x, dz, out)


def _grad(p: torch.Tensor) -> torch.Tensor:
    """fp32 gradient buffer of a parameter (allocated zeroed on first use; kernels accumulate into it)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32)
    return p.grad


class PackPlan:
    """All bf16 operand copies of the GEMM weights refreshed by ONE launch (csrc/param_ops.cu: pack_weights_kernel).  Entries name the
    Parameter (not a view of it) so that the device table can be rebuilt when parameter storage moves (FlatAdamW, .to())."""

    _DT = None

    def __init__(self):
        self.entries = []  # (param, byte_off, K, N, src_ld, dst_w, ldw, dst_wt, ldwt, c0, cs, flags)
        self.table = None
        self._ptrs = None

    def add(self, param, K, N, dst_w=None, dst_wt=None, ldw=0, ldwt=0, c0=0, cs=1, src_ld=None, src_col0=0, f32=False):
        assert param.dtype == torch.float32 and param.is_contiguous()
        self.entries.append((param, 4 * src_col0, K, N, N if src_ld is None else src_ld, dst_w, ldw, dst_wt, ldwt, c0, cs, 1 if f32 else 0))

    def _build(self, ptrs):
        import numpy as np

        if PackPlan._DT is None:
            PackPlan._DT = np.dtype([("src", "<u8"), ("dst_w", "<u8"), ("dst_wt", "<u8"), ("K", "<i4"), ("N", "<i4"), ("src_ld", "<i4"),
                                    ("ldw", "<i4"), ("ldwt", "<i4"), ("c0", "<i4"), ("cs", "<i4"), ("flags", "<i4")])
        arr = np.zeros(len(self.entries), dtype=PackPlan._DT)
        for i, (prm, off, K, N, sld, dw, ldw, dwt, ldwt, c0, cs, fl) in enumerate(self.entries):
            arr[i] = (ptrs[i] + off, 0 if dw is None else dw.data_ptr(), 0 if dwt is None else dwt.data_ptr(), K, N, sld, ldw, ldwt, c0, cs, fl)
        dev = self.entries[0][0].device
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self._ptrs = ptrs

    def run(self):
        if not self.entries:
            return
        ptrs = [e[0].data_ptr() for e in self.entries]
        if ptrs != self._ptrs:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: parameter storage moved during CUDA-graph capture")
            self._build(ptrs)
        from . import _lib

        lib = _lib.load()
        _lib.check(lib.tfasr_pack_weights(self.table.data_ptr(), len(self.entries), _lib.stream_ptr(self.table.device)), "tfasr_pack_weights")


class UnpackPlan:
    """Gradient-side counterpart of PackPlan: fused projections (Q|K|V, GLU-interleaved) accumulate their weight / bias gradients in
    persistent packed fp32 buffers; ONE launch at the end of the backward pass (autograd engine callback) adds every slice into its
    parameter's gradient and clears the packed buffers (csrc/param_ops.cu: unpack_add_kernel)."""

    def __init__(self):
        self.entries = []  # (param, grad_byte_off, K, N, grad_ld, packed, ldw, c0, cs)
        self.table = None
        self._ptrs = None
        self._queued = False

    def add(self, param, K, N, packed, ldw, c0=0, cs=1, grad_ld=None, grad_col0=0):
        self.entries.append((param, 4 * grad_col0, K, N, N if grad_ld is None else grad_ld, packed, ldw, c0, cs))

    def request(self):
        """Called from a backward node: flush once when the whole backward pass has finished."""
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._flush)

    def _flush(self):
        self._queued = False
        self.run()

    def run(self):
        SIDE.join()  # the packed buffers are written by weight-gradient GEMMs on the side stream
        if not self.entries:
            return
        import numpy as np

        ptrs = [_grad(e[0]).data_ptr() for e in self.entries]
        if ptrs != self._ptrs:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("UnpackPlan: gradient storage moved during CUDA-graph capture")
            if PackPlan._DT is None:
                PackPlan._DT = np.dtype([("src", "<u8"), ("dst_w", "<u8"), ("dst_wt", "<u8"), ("K", "<i4"), ("N", "<i4"), ("src_ld", "<i4"),
                                        ("ldw", "<i4"), ("ldwt", "<i4"), ("c0", "<i4"), ("cs", "<i4"), ("flags", "<i4")])
            arr = np.zeros(len(self.entries), dtype=PackPlan._DT)
            for i, (prm, off, K, N, gld, packed, ldw, c0, cs) in enumerate(self.entries):
                arr[i] = (ptrs[i] + off, packed.data_ptr(), 0, K, N, gld, ldw, 0, c0, cs, 0)
            self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(self.entries[0][5].device)
            self._ptrs = ptrs
        from . import _lib

        lib = _lib.load()
        _lib.check(lib.tfasr_unpack_add(self.table.data_ptr(), len(self.entries), _lib.stream_ptr(self.table.device)), "tfasr_unpack_add")


class WeightPack:
    """bf16 operand copies of one Dense kernel W [K, N] (Keras layout): `wt` = W^T [N, K] for the forward GEMM (B operand,
    K-major), `w` = W [K, N] for the data-gradient GEMM (dX = dY . W^T: B operand [K, N] with contraction N).
    With a `plan` the copies are persistent buffers refreshed by the plan's single launch; without, they are made here."""

    def __init__(self, param, bias=None, interleave_glu=False, plan: PackPlan = None, shape2d=None):
        K, N = shape2d if shape2d is not None else param.shape
        self.param = param
        self.bias = bias
        self.shape2d = (K, N)
        if plan is not None:
            dev = param.device
            self.w = torch.empty((K, N), dtype=torch.bfloat16, device=dev)
            self.wt = torch.empty((N, K), dtype=torch.bfloat16, device=dev)
            if interleave_glu:
                D = N // 2
                plan.add(param, K, D, self.w, self.wt, ldw=N, ldwt=K, c0=0, cs=2, src_ld=N, src_col0=0)
                plan.add(param, K, D, self.w, self.wt, ldw=N, ldwt=K, c0=1, cs=2, src_ld=N, src_col0=D)
            else:
                plan.add(param, K, N, self.w, self.wt, ldw=N, ldwt=K)
            return
        w = param.detach().reshape(K, N)
        if interleave_glu:
            # GLU halves (a = cols [0, D), b = cols [D, 2D)) interleaved as (a_0, b_0, a_1, b_1, ...) so that the GEMM epilogue
            # sees each pair in adjacent columns (activations/glu.py:25-28)
            D = N // 2
            w = torch.stack([w[:, :D], w[:, D:]], dim=2).reshape(K, N)
        self.w = w.to(torch.bfloat16).contiguous()
        self.wt = w.t().to(torch.bfloat16).contiguous()

    # hooks used by the generic Dense backward (overridden by packs that fan out to several parameters)
    def bias_f32(self):
        return None if self.bias is None else self.bias.detach().float().reshape(-1)

    def bias_grad_target(self):
        return None if self.bias is None else _grad(self.bias).view(-1)

    def accumulate_wgrad(self, x, dz):
        _wgrad(x, dz, _grad(self.param).view(self.w.shape), _splits(x.shape[0]))


def _splits(rows: int) -> int:
    """split-K factor for weight-gradient GEMMs (contraction over `rows`): enough CTAs to fill the machine."""
    return max(1, min(32, rows // 512))


# ------------------------------------------------------------------------------------------------ FFModule
class _FFModuleFn(torch.autograd.Function):
    """x + factor * Drop(W2 . Drop(swish(W1 . LN(x) + b1)) + b2)      (conformer.py:101-109)"""

    @staticmethod
    def forward(ctx, x, ln_g, ln_b, p1: WeightPack, p2: WeightPack, factor, drop_p, salt, seed):
        M, D = x.shape
        H = p1.wt.shape[0]
        from . import ops

        if ops.fused_ffn_active() and kernels.ffn_supported(D, H) and p1.bias is not None and p2.bias is not None:
            # the whole module in one launch: LN in the A-operand prologue, the hidden activation handed from GEMM 1 to GEMM 2 through SMEM
            out, a, z, h, mean, rstd = kernels.ffn_fwd(x, ln_g.detach(), ln_b.detach(), p1.wt, p1.bias.detach(), p2.wt, p2.bias.detach(), factor, drop_p,
                                                       seed, salt, salt + 1, LN_EPS)
        else:
            a, mean, rstd = kernels.layernorm_fwd(x, ln_g, ln_b, LN_EPS)
            z = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
            h = kernels.gemm(a, p1.wt, bias=p1.bias, activation="swish", aux=z, dropout_p=drop_p, dropout_seed=seed, dropout_salt=salt)
            out = kernels.gemm(h, p2.wt, bias=p2.bias, residual=x, residual_scale=factor, dropout_p=drop_p, dropout_seed=seed,
                               dropout_salt=salt + 1)
        ctx.save_for_backward(x, mean, rstd, a, z, h)
        ctx.misc = (ln_g, ln_b, p1, p2, factor, drop_p, salt, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mean, rstd, a, z, h = ctx.saved_tensors
        ln_g, ln_b, p1, p2, factor, drop_p, salt, seed = ctx.misc
        dout = dout.contiguous()
        M = x.shape[0]
        ks = _splits(M)
        from . import ops

        if ops.fused_ffn_active() and kernels.ffn_supported(x.shape[1], z.shape[1]) and p1.bias is not None and p2.bias is not None:
            # the gradient chain dout -> dy -> dz -> da in one launch; everything nobody downstream waits for goes to the side stream
            dy, dz, da = kernels.ffn_bwd(dout, z, p2.w, p1.w, factor, drop_p, seed, salt, salt + 1)
            db2, db1 = _grad(p2.bias), _grad(p1.bias)
            SIDE.run(lambda: kernels.colsum(dy, db2), dy, db2)
            _wgrad(h, dy, _grad(p2.param), ks)
            SIDE.run(lambda: kernels.colsum(dz, db1), dz, db1)
            _wgrad(a, dz, _grad(p1.param), ks)
            dx = kernels.layernorm_bwd(da, x, mean, rstd, ln_g, _grad(ln_g), _grad(ln_b), dres=dout)
            return dx, None, None, None, None, None, None, None, None
        # d(out)/d(y): residual factor and the second dropout; bias-2 gradient = column sums
        dy = kernels.scale_mask_colsum(dout, scale=factor, dropout_p=drop_p, seed=seed, salt=salt + 1, colsum=_grad(p2.bias))
        _wgrad(h, dy, _grad(p2.param), ks)  # dW2 += h^T dy
        # dz = (dy W2^T) * dropout1 * swish'(z); bias-1 gradient = column sums (fused in the epilogue)
        dz = kernels.gemm(dy, p2.w, mul=z, mul_mode="swish_grad", dropout_p=drop_p, dropout_seed=seed, dropout_salt=salt,
                          colsum=_grad(p1.bias))
        _wgrad(a, dz, _grad(p1.param), ks)  # dW1 += a^T dz
        da = kernels.gemm(dz, p1.w)
        dx = kernels.layernorm_bwd(da, x, mean, rstd, ln_g, _grad(ln_g), _grad(ln_b), dres=dout)
        return dx, None, None, None, None, None, None, None, None


def ff_module(x, ln_g, ln_b, p1, p2, factor, drop_p, salt, seed):
    """x [M, D] bf16 -> [M, D] bf16."""
    return _FFModuleFn.apply(x, ln_g, ln_b, p1, p2, float(factor), float(drop_p), int(salt), seed)


# ------------------------------------------------------------------------------------------------ plain LayerNorm
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b):
        y, mean, rstd = kernels.layernorm_fwd(x, g, b, LN_EPS)
        ctx.save_for_backward(x, mean, rstd)
        ctx.misc = (g, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        g, b = ctx.misc
        return kernels.layernorm_bwd(dy.contiguous(), x, mean, rstd, g, _grad(g), _grad(b)), None, None


def layer_norm(x, g, b):
    """x [..., D] -> bf16, keras LayerNormalization(eps=1e-3)."""
    shp = x.shape
    return _LayerNormFn.apply(x.reshape(-1, shp[-1]), g, b).reshape(shp)


# ------------------------------------------------------------------------------------------------ Dense (+ optional residual)
class _LinearFn(torch.autograd.Function):
    """y = [residual + factor *] Drop(act(x W + b))"""

    @staticmethod
    def forward(ctx, x, pack: WeightPack, activation, residual, factor, drop_p, salt, seed, anchor=None):
        # `anchor`: a parameter passed only so that autograd records this node when x itself carries no gradient
        M = x.shape[0]
        N = pack.wt.shape[0]
        z = torch.empty((M, N), dtype=torch.bfloat16, device=x.device) if activation == "swish" else None
        y = kernels.gemm(x, pack.wt, bias=pack.bias_f32(), activation=activation, aux=z, residual=residual, residual_scale=factor,
                         dropout_p=drop_p, dropout_seed=seed, dropout_salt=salt)
        ctx.save_for_backward(x, z)
        ctx.misc = (pack, activation, residual is not None, factor, drop_p, salt, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, z = ctx.saved_tensors
        pack, activation, has_res, factor, drop_p, salt, seed = ctx.misc
        dy = dy.contiguous()
        scale = factor if has_res else 1.0
        if activation == "swish":
            # dz = scale * mask * dy * swish'(z)
            d0 = kernels.scale_mask_colsum(dy, scale=scale, dropout_p=drop_p, seed=seed, salt=salt)
            sg = torch.sigmoid(z.float())
            dz = (d0.float() * (sg * (1 + z.float() * (1 - sg)))).to(torch.bfloat16)
            if pack.bias_grad_target() is not None:
                kernels.colsum(dz, pack.bias_grad_target())
        elif scale == 1.0 and drop_p == 0.0:
            dz = dy
            if pack.bias_grad_target() is not None:
                kernels.colsum(dz, pack.bias_grad_target())
        else:
            dz = kernels.scale_mask_colsum(dy, scale=scale, dropout_p=drop_p, seed=seed, salt=salt, colsum=pack.bias_grad_target())
        pack.accumulate_wgrad(x, dz)
        if not ctx.needs_input_grad[0]:
            return None, None, None, (dy if has_res else None), None, None, None, None, None
        dx = kernels.gemm(dz, pack.w)
        return dx, None, None, (dy if has_res else None), None, None, None, None, None


def linear(x, pack, activation=None, residual=None, factor=1.0, drop_p=0.0, salt=0, seed=None):
    """x [M, K] bf16 -> [M, N] bf16 through the tcgen05 GEMM (bias / activation / dropout / residual fused in the epilogue)."""
    anchor = None
    if not x.requires_grad and torch.is_grad_enabled():
        anchor = pack.param if pack.param is not None else pack.parts[0][0]
    return _LinearFn.apply(x, pack, activation, residual, float(factor), float(drop_p), int(salt), seed, anchor)


# ------------------------------------------------------------------------------------------------ prediction-network LSTM
class _LstmFn(torch.autograd.Function):
    """keras LSTM over the label axis (base_transducer.py:74-88): x . W_ih + b on the tcgen05 GEMM, the time loop on the cluster kernel of
    csrc/lstm.cu, and the weight / input gradients as GEMMs over all steps at once."""

    @staticmethod
    def forward(ctx, x, length, kernel, recurrent_kernel, bias):
        B, U, E = x.shape
        R = recurrent_kernel.shape[0]
        x2 = x.reshape(B * U, E).to(torch.bfloat16).contiguous()
        pack = WeightPack(kernel)  # [E, 4R]
        gx = kernels.gemm(x2, pack.wt, bias=bias.detach().float(), out_dtype=torch.float32).view(B, U, 4 * R)
        whh = recurrent_kernel.detach().float().contiguous()
        y, saved, _ = kernels.lstm_rec_fwd(gx, whh, length)
        ctx.save_for_backward(x2, y, saved[0], saved[1], whh, length)
        ctx.misc = (kernel, recurrent_kernel, bias, pack, B, U, E, R, x.requires_grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, y, saved_g, saved_c, whh, length = ctx.saved_tensors
        kernel, recurrent_kernel, bias, pack, B, U, E, R, need_dx = ctx.misc
        dg = kernels.lstm_rec_bwd(dy.to(torch.bfloat16).contiguous(), (saved_g, saved_c), whh, length).view(B * U, 4 * R)
        kernels.colsum(dg, _grad(bias).view(-1))
        ks = _splits(B * U)
        _wgrad(x2, dg, _grad(kernel).view(E, 4 * R), ks)  # dW_ih += x^T dgates
        # h_{t-1}: the outputs shifted by one label step (masked steps carry zero gate gradients, so their zeros are harmless)
        hprev = torch.zeros_like(y)
        hprev[:, 1:] = y[:, :-1]
        _wgrad(hprev.view(B * U, R), dg, _grad(recurrent_kernel).view(R, 4 * R), ks)
        dx = kernels.gemm(dg, pack.w).view(B, U, E) if need_dx else None
        return dx, None, None, None, None


def lstm(x, length, kernel, recurrent_kernel, bias):
    """x [B,U,E] -> y [B,U,R] bf16 (zeros at t >= length).  Training / full-sequence path (no carried state)."""
    length = None if length is None else length.to(device=x.device, dtype=torch.int32)
    return _LstmFn.apply(x, length, kernel, recurrent_kernel, bias)


# ------------------------------------------------------------------------------------------------ joint + RNN-T loss head
class _TransducerHeadFn(torch.autograd.Function):
    """Joint network + RNN-T loss for training on the hand-written kernels (base_transducer.py:199-207,280-293 + losses/impl/rnnt.py):

      forward   h = tanh(e + p) (bf16)  ->  logits = h Wv + bv (tcgen05 GEMM, f32 out)  ->  K1 (lse) + K2 (alpha/beta wavefronts)
      backward  K3' writes d logits as bf16 IN PLACE over the fp32 logits (+ vocab-bias gradient)  ->
                dz = (dlogits Wv^T) * (1 - h^2) (GEMM epilogue)  ->  dWv += h^T dlogits (MN-major split-K GEMM)  ->
                de = sum_u dz, dp = sum_t dz
    The [B,T,U+1,V] tensor exists once (fp32 logits, then reused for its own gradient).
    """

    @staticmethod
    def forward(ctx, e, p, pack: WeightPack, labels, logit_length, label_length):
        from . import _lib

        lib = _lib.load()
        B, T, J = e.shape
        U1 = p.shape[1]
        V = pack.wt.shape[0]
        M = B * T * U1
        from . import ops

        h = kernels.joint_tanh_fwd(e.contiguous(), p.contiguous())
        fused_lse = ops.fused_lse_active()
        parts = 2 * ((V + 255) // 256)
        bf16_logits = ops.bf16_logits_active() and V % 4 == 0
        rowstats = torch.empty((M, parts, 2), dtype=torch.float32, device=e.device) if fused_lse else None
        with _lib.TIMERS.region("joint_gemm_fwd"):
            if bf16_logits:
                # row pitch rounded up to 64 elements: every row starts on a 128-byte line, so the TMA boxes of this GEMM's stores and of the
                # backward GEMMs' loads never straddle an extra 32-byte sector (a pitch of V = 1000 made the dgrad GEMM 16 % slower)
                ldv = (V + 63) // 64 * 64
                logits = torch.empty((M, ldv), dtype=torch.bfloat16, device=e.device)[:, :V]
                kernels.gemm(h.view(M, J), pack.wt, bias=pack.bias, out=logits, rowstats=rowstats)
            else:
                logits = kernels.gemm(h.view(M, J), pack.wt, bias=pack.bias, out_dtype=torch.float32, rowstats=rowstats)
        ldx = logits.stride(0)
        ws_bytes = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=e.device)
        costs = torch.empty(B, dtype=torch.float32, device=e.device)
        lattice_v2 = fused_lse and ops.lattice_v2_active() and bool(lib.tfasr_rnnt_lattice_smem_supported(T, U1))
        with _lib.TIMERS.region("rnnt_fwd"):
            if lattice_v2:
                # the two steps of tfasr_rnnt_loss_fwd_rowstats_* called one by one, the wavefronts on the shared-memory kernel; the planes
                # are the workspace's (rnnt_loss.cu rnnt_ws_carve: lse, blank, truth, alpha, beta, each rounded up to 256 bytes)
                plane = (4 * M + 255) // 256 * 256
                wp = ws.data_ptr()
                comb, comb_name = ((lib.tfasr_rnnt_lse_from_rowstats_bf16, "tfasr_rnnt_lse_from_rowstats_bf16") if bf16_logits
                                   else (lib.tfasr_rnnt_lse_from_rowstats_f32, "tfasr_rnnt_lse_from_rowstats_f32"))
                ld_arg = (ldx,) if bf16_logits else ()
                _lib.check(comb(logits.data_ptr(), *ld_arg, rowstats.data_ptr(), parts, labels.data_ptr(), logit_length.data_ptr(),
                                label_length.data_ptr(), B, T, U1, V, wp, wp + plane, wp + 2 * plane, _lib.stream_ptr(e.device)), comb_name)
                _lib.check(lib.tfasr_rnnt_lattice_smem_f32(wp + plane, wp + 2 * plane, logit_length.data_ptr(), label_length.data_ptr(), B, T, U1,
                                                           wp + 3 * plane, wp + 4 * plane, costs.data_ptr(), _lib.stream_ptr(e.device)),
                           "tfasr_rnnt_lattice_smem_f32")
            elif bf16_logits:
                _lib.check(lib.tfasr_rnnt_loss_fwd_rowstats_bf16(logits.data_ptr(), ldx, rowstats.data_ptr(), parts, labels.data_ptr(),
                                                                 logit_length.data_ptr(), label_length.data_ptr(), B, T, U1, V, 0, costs.data_ptr(),
                                                                 ws.data_ptr(), ws_bytes, _lib.stream_ptr(e.device)),
                           "tfasr_rnnt_loss_fwd_rowstats_bf16")
            elif fused_lse:
                # the log-softmax normaliser comes from the GEMM epilogue's row statistics: the logits are not read again before K3
                _lib.check(lib.tfasr_rnnt_loss_fwd_rowstats_f32(logits.data_ptr(), rowstats.data_ptr(), parts, labels.data_ptr(),
                                                                logit_length.data_ptr(), label_length.data_ptr(), B, T, U1, V, 0, costs.data_ptr(),
                                                                ws.data_ptr(), ws_bytes, _lib.stream_ptr(e.device)),
                           "tfasr_rnnt_loss_fwd_rowstats_f32")
            else:
                _lib.check(lib.tfasr_rnnt_loss_fwd_f32(logits.data_ptr(), labels.data_ptr(), logit_length.data_ptr(), label_length.data_ptr(),
                                                       B, T, U1, V, 0, costs.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(e.device)),
                           "tfasr_rnnt_loss_fwd_f32")
        ctx.save_for_backward(h, logits, labels, logit_length, label_length, costs, ws)
        ctx.misc = (pack, (B, T, U1, J, V), e.dtype, p.dtype)
        return costs

    @staticmethod
    def backward(ctx, dcosts):
        from . import _lib

        lib = _lib.load()
        h, logits, labels, logit_length, label_length, costs, ws = ctx.saved_tensors
        pack, (B, T, U1, J, V), de_dtype, dp_dtype = ctx.misc
        M = B * T * U1
        scale = dcosts.float().contiguous()
        from . import ops

        if logits.dtype == torch.bfloat16:  # bf16 logits: the gradient replaces them element for element (row pitch of the forward's buffer)
            ld = logits.stride(0)
            dl = torch.as_strided(logits, (M, ld), (ld, 1))
            with _lib.TIMERS.region("rnnt_bwd"):
                _lib.check(lib.tfasr_rnnt_loss_bwd_bf16_io(logits.data_ptr(), ld, labels.data_ptr(), logit_length.data_ptr(), label_length.data_ptr(),
                                                           B, T, U1, V, 0, costs.data_ptr(), scale.data_ptr(), dl.data_ptr(), ld,
                                                           _grad(pack.bias).data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(h.device)),
                           "tfasr_rnnt_loss_bwd_bf16_io")
        else:
            dl, ld = logits.view(M, V).view(torch.bfloat16), 2 * V  # [M, 2V] bf16 view of the same memory; gradient lands in [:, :V]
            bwd_fn, bwd_name = ((lib.tfasr_rnnt_loss_bwd_bf16_v2, "tfasr_rnnt_loss_bwd_bf16_v2") if ops.rnnt_grad_v2_active()
                                else (lib.tfasr_rnnt_loss_bwd_bf16, "tfasr_rnnt_loss_bwd_bf16"))
            with _lib.TIMERS.region("rnnt_bwd"):
                _lib.check(bwd_fn(logits.data_ptr(), labels.data_ptr(), logit_length.data_ptr(), label_length.data_ptr(), B, T, U1, V, 0,
                                  costs.data_ptr(), scale.data_ptr(), dl.data_ptr(), ld, _grad(pack.bias).data_ptr(), ws.data_ptr(), ws.numel(),
                                  _lib.stream_ptr(h.device)), bwd_name)
        dlv = dl[:, :V]
        h2 = h.view(M, J)
        with _lib.TIMERS.region("joint_gemm_dgrad"):
            dz = kernels.gemm(dlv, pack.w, mul=h2, mul_mode="tanh_grad")  # [M, J]
        with _lib.TIMERS.region("joint_gemm_wgrad"):
            _wgrad(h2, dlv, _grad(pack.param), 32)  # dWv [J, V]  (the timers switch the side stream off: the region then times the GEMM)
        de, dp = kernels.joint_reduce_bwd(dz.view(B, T, U1, J))
        return de.to(de_dtype), dp.to(dp_dtype), None, None, None, None


def transducer_head(e, p, pack, labels, logit_length, label_length):
    """e [B,T,J], p [B,U1,J] bf16 -> per-utterance RNN-T costs [B] f32."""
    return _TransducerHeadFn.apply(e, p, pack, labels, logit_length, label_length)


# ------------------------------------------------------------------------------------------------ BatchNorm (+ swish) helpers
BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def _world():
    d = torch.distributed
    return d.get_world_size() if d.is_available() and d.is_initialized() else 1


def bn_train_fwd(x2d, gamma, beta, moving_mean, moving_var, act=1):
    """Training-mode keras BatchNormalization(synchronized=True) (+ swish) on x2d [R, C] bf16.
    Returns (y, saved) where saved = (scale, shift, mean, rstd, count) for the backward.  Updates the moving statistics."""
    R, C = x2d.shape
    world = _world()
    # replicas may hold different numbers of rows (each pads to its own longest utterance): the local row count travels with the sums
    # and the kernels divide by the reduced count (keras counts the same way, and so does ops._SyncBatchNormFn)
    sums = kernels.bn_stats(x2d, with_count=world > 1)
    if world > 1:
        dist.small_all_reduce(sums)  # C2: sync-BN statistics (+ row count)
    count = kernels.BN_COUNT_ON_DEVICE if world > 1 else float(R)
    with torch.no_grad():
        out = kernels.bn_finalize(sums, gamma.detach(), beta.detach(), moving_mean, moving_var, count, BN_EPS, BN_MOMENTUM)
    scale, shift, mean, rstd = out[0], out[1], out[2], out[3]
    y = kernels.bn_apply(x2d, scale, shift, act)
    return y, (scale, shift, mean, rstd, count)


def bn_train_bwd(dy2d, x2d, saved, gamma, beta, act=1):
    """-> dx2d; accumulates dgamma / dbeta into the parameters' .grad."""
    scale, shift, mean, rstd, count = saved
    world = _world()
    sums = kernels.bn_bwd_stats(dy2d, x2d, scale, shift, mean, rstd, act, with_count=world > 1)
    kernels.bn_bwd_finalize(sums, _grad(gamma).view(-1), _grad(beta).view(-1))  # local contribution; the gradient all-reduce sums over replicas
    if world > 1:
        sums = dist.small_all_reduce(sums.clone())
    return kernels.bn_bwd_apply(dy2d, x2d, scale, shift, mean, rstd, sums, count, act)


def bn_eval(x2d, gamma, beta, moving_mean, moving_var, act=1):
    scale = gamma.detach().float() * torch.rsqrt(moving_var.float() + BN_EPS)
    shift = beta.detach().float() - moving_mean.float() * scale
    return kernels.bn_apply(x2d, scale, shift, act)


# ------------------------------------------------------------------------------------------------ Conv2dSubsampling
class _SubsamplingFn(torch.autograd.Function):
    """Conv2dSubsampling with filters [C, C], 3x3 kernels, stride 2, causal padding, batch norms, swish (the Conformer-S
    configuration, small.yml.j2:26-33; subsampling.py:218-230):
        conv1 (direct, Cin = 1) -> BN + swish -> im2col -> tcgen05 GEMM (conv2) -> BN + swish -> merge (F, C)
    Conv biases sit in front of a batch norm, so their gradients are identically zero and are not computed."""

    @staticmethod
    def forward(ctx, feats, w1, sub, training):
        # w1 (conv_0.kernel) is an explicit autograd input: the features carry no gradient, and a Function with no
        # differentiable input would never have its backward called.
        B, T, F = feats.shape[0], feats.shape[1], feats.shape[2]
        x = feats.reshape(B, T, F).float().contiguous()
        b1 = sub["conv_0.bias"]
        w2, b2 = sub["conv_1.kernel"], sub["conv_1.bias"]
        C = w1.shape[-1]
        y1 = kernels.conv1_fwd(x, w1.detach().float().contiguous(), b1.detach().float())
        T1, F1 = y1.shape[1], y1.shape[2]
        y1f = y1.view(-1, C)
        bn0 = (sub["bn_0.gamma"], sub["bn_0.beta"], sub["bn_0.moving_mean"], sub["bn_0.moving_variance"])
        bn1 = (sub["bn_1.gamma"], sub["bn_1.beta"], sub["bn_1.moving_mean"], sub["bn_1.moving_variance"])
        if training:
            a1, s0 = bn_train_fwd(y1f, *bn0)
        else:
            a1, s0 = bn_eval(y1f, *bn0), None
        a1 = a1.view(B, T1, F1, C)
        col = kernels.im2col(a1)
        pack = WeightPack(w2.reshape(9 * C, C), b2)
        pack.param = w2
        y2 = kernels.gemm(col, pack.wt, bias=b2.detach().float())
        del col
        if training:
            a2, s1 = bn_train_fwd(y2, *bn1)
        else:
            a2, s1 = bn_eval(y2, *bn1), None
        T2, F2 = (T1 + 1) // 2, (F1 + 1) // 2
        ctx.save_for_backward(x, y1, a1, y2)
        ctx.misc = (sub, pack, s0, s1, (B, T, F, T1, F1, T2, F2, C), training)
        return a2.view(B, T2, F2 * C)

    @staticmethod
    def backward(ctx, dout):
        x, y1, a1, y2 = ctx.saved_tensors
        sub, pack, s0, s1, (B, T, F, T1, F1, T2, F2, C), training = ctx.misc
        if not training:
            raise RuntimeError("Conv2dSubsampling fused backward needs training-mode batch statistics")
        dy2 = bn_train_bwd(dout.contiguous().view(-1, C), y2, s1, sub["bn_1.gamma"], sub["bn_1.beta"])
        col = kernels.im2col(a1)  # recomputed: cheaper than keeping 9x the activation alive
        _wgrad(col, dy2, _grad(pack.param).view(9 * C, C), 32)
        del col
        dcol = kernels.gemm(dy2, pack.w)  # [M2, 9C]
        da1 = kernels.col2im(dcol, (B, T1, F1, C))
        del dcol
        dy1 = bn_train_bwd(da1.view(-1, C), y1.view(-1, C), s0, sub["bn_0.gamma"], sub["bn_0.beta"])
        dw1 = torch.zeros_like(sub["conv_0.kernel"], dtype=torch.float32)
        kernels.conv1_wgrad(x, dy1, dw1)
        return None, dw1, None, None


def conv2d_subsampling(feats, sub, training):
    """feats [B,T,F,1] -> [B, ceil(ceil(T/2)/2), F'' * C] bf16."""
    return _SubsamplingFn.apply(feats, sub["conv_0.kernel"], sub, training)


# ------------------------------------------------------------------------------------------------ MHSAModule packs
class QkvPack(WeightPack):
    """Fused Q/K/V projection: one GEMM [D -> 3NH]; its weight / bias gradients fan out to the three Keras EinsumDense layers."""

    def __init__(self, enc, pre, plan: PackPlan = None, unplan: UnpackPlan = None):
        m = pre + ".mhsa"
        D, N, H = enc.dmodel, enc.num_heads, enc.head_size
        NH = N * H
        self.parts = [(enc[f"{m}.{k}.kernel"], enc[f"{m}.{k}.bias"]) for k in ("query", "key", "value")]
        self.param, self.D, self.NH = None, D, NH
        self._db = None
        self.unplan = unplan
        if unplan is not None:
            dev = self.parts[0][0].device
            self.dw_packed = torch.zeros((D, 3 * NH), dtype=torch.float32, device=dev)
            self.db_packed = torch.zeros(3 * NH, dtype=torch.float32, device=dev)
            for i, (wp, bp) in enumerate(self.parts):
                unplan.add(wp, D, NH, self.dw_packed, 3 * NH, c0=i * NH)
                unplan.add(bp, 1, NH, self.db_packed, 3 * NH, c0=i * NH)
        if plan is not None:
            dev = self.parts[0][0].device
            self.w = torch.empty((D, 3 * NH), dtype=torch.bfloat16, device=dev)
            self.wt = torch.empty((3 * NH, D), dtype=torch.bfloat16, device=dev)
            self._bias = torch.empty(3 * NH, dtype=torch.float32, device=dev)
            for i, (wp, bp) in enumerate(self.parts):
                plan.add(wp, D, NH, self.w, self.wt, ldw=3 * NH, ldwt=D, c0=i * NH)
                plan.add(bp, 1, NH, self._bias, None, ldw=3 * NH, c0=i * NH, f32=True)
        else:
            w = torch.cat([k.detach().reshape(D, NH) for k, _ in self.parts], dim=1)
            self.w = w.to(torch.bfloat16).contiguous()
            self.wt = w.t().to(torch.bfloat16).contiguous()
            self._bias = torch.cat([b.detach().reshape(NH) for _, b in self.parts]).float().contiguous()
        self.bias = self._bias

    def bias_f32(self):
        return self._bias

    def bias_grad_target(self):
        if self.unplan is not None:
            return self.db_packed
        self._db = torch.zeros(3 * self.NH, dtype=torch.float32, device=self.w.device)
        return self._db

    def accumulate_wgrad(self, x, dz):
        D, NH = self.D, self.NH
        if self.unplan is not None:  # packed accumulation; fanned out to the three layers by the plan's single launch
            _wgrad(x, dz, self.dw_packed, _splits(x.shape[0]))
            self.unplan.request()
            return
        dw = torch.zeros((D, 3 * NH), dtype=torch.float32, device=x.device)
        kernels.gemm(x, dz, mn_major=True, out=dw, accumulate=True, k_splits=_splits(x.shape[0]))
        for i, (wp, bp) in enumerate(self.parts):
            _grad(wp).view(D, NH).add_(dw[:, i * NH:(i + 1) * NH])
            if self._db is not None:
                _grad(bp).view(NH).add_(self._db[i * NH:(i + 1) * NH])
        self._db = None


class PosAllPack(WeightPack):
    """The position projections of ALL blocks as one GEMM [D -> blocks * NH]: every block projects the same relative position table
    (conformer.py:672-701 hands one relpe to all blocks; multihead_attention.py:304 `encoding` EinsumDense per block), so one launch replaces
    `blocks` forward GEMMs, and one column-sum + one weight-gradient GEMM replace `blocks` of each in the backward pass.  Weight / bias
    gradients fan out to the per-block Keras layers through the UnpackPlan's single launch (as QkvPack)."""

    def __init__(self, enc, plan: PackPlan, unplan: UnpackPlan):
        D, NH, nb = enc.dmodel, enc.num_heads * enc.head_size, enc.num_blocks
        self.parts = [(enc[f"block_{b}.mhsa_module.mhsa.encoding.kernel"], enc[f"block_{b}.mhsa_module.mhsa.encoding.bias"]) for b in range(nb)]
        self.param, self.D, self.NH, self.nb = None, D, NH, nb
        self.unplan = unplan
        dev = self.parts[0][0].device
        W = nb * NH
        self.shape2d = (D, W)
        self.dw_packed = torch.zeros((D, W), dtype=torch.float32, device=dev)
        self.db_packed = torch.zeros(W, dtype=torch.float32, device=dev)
        self.w = torch.empty((D, W), dtype=torch.bfloat16, device=dev)
        self.wt = torch.empty((W, D), dtype=torch.bfloat16, device=dev)
        self._bias = torch.empty(W, dtype=torch.float32, device=dev)
        for i, (wp, bp) in enumerate(self.parts):
            unplan.add(wp, D, NH, self.dw_packed, W, c0=i * NH)
            unplan.add(bp, 1, NH, self.db_packed, W, c0=i * NH)
            plan.add(wp, D, NH, self.w, self.wt, ldw=W, ldwt=D, c0=i * NH)
            plan.add(bp, 1, NH, self._bias, None, ldw=W, c0=i * NH, f32=True)
        self.bias = self._bias
        self.dpos = None  # the blocks' shared gradient buffer, alive during one backward pass

    def bias_f32(self):
        return self._bias

    def bias_grad_target(self):
        return self.db_packed

    def accumulate_wgrad(self, x, dz):
        _wgrad(x, dz, self.dw_packed, _splits(x.shape[0]))
        self.unplan.request()


class MhsaPack:
    def __init__(self, enc, pre, plan: PackPlan = None, unplan: UnpackPlan = None, with_pos: bool = True):
        m = pre + ".mhsa"
        D, N, H = enc.dmodel, enc.num_heads, enc.head_size
        self.qkv = QkvPack(enc, pre, plan, unplan)
        self.pos = WeightPack(enc[f"{m}.encoding.kernel"], enc[f"{m}.encoding.bias"], plan=plan, shape2d=(D, N * H)) if with_pos else None
        self.out = WeightPack(enc[f"{m}.attention_output.kernel"], enc[f"{m}.attention_output.bias"], plan=plan, shape2d=(N * H, D))


class _RelAttentionFn(torch.autograd.Function):
    """Relative-position attention core (multihead_attention.py:543-582) on the tensor-core kernels of csrc/attention.cu.
    qkv [B*T, 3NH] bf16 (output of the fused projection), pos [B*R, NH] bf16 -> ctx [B*T, NH] bf16."""

    @staticmethod
    def forward(ctx, qkv, pos, u, v, length, B, T, N, H, chunk, history):
        out, probs = kernels.rel_attention_fwd(qkv, pos, u.detach(), v.detach(), length, B, T, N, H, chunk, history)
        ctx.save_for_backward(qkv, pos, probs, length, out)
        ctx.misc = (u, v, B, T, N, H, chunk, history)
        return out

    @staticmethod
    def backward(ctx, dctx):
        qkv, pos, probs, length, out = ctx.saved_tensors
        u, v, B, T, N, H, chunk, history = ctx.misc
        dqkv, dpos = kernels.rel_attention_bwd(qkv, pos, u.detach(), v.detach(), length, B, T, N, H, probs, dctx.contiguous(), _grad(u), _grad(v),
                                               chunk, history, ctx=out)
        return dqkv, dpos, None, None, None, None, None, None, None, None, None


def rel_attention(qkv, pos, u, v, length, B, T, N, H, chunk=None, history=None):
    return _RelAttentionFn.apply(qkv, pos, u, v, length, B, T, N, H, chunk, history)


class _RelAttentionSliceFn(torch.autograd.Function):
    """_RelAttentionFn on block `blk`'s column slice of the all-blocks position projection `pos_all` [B*R, blocks*NH] (PosAllPack).
    Backward: every block writes its slice of ONE shared gradient buffer; block 0 — whose backward runs last, every later block depends on
    its output — hands the buffer to autograd as the gradient of pos_all, the others return None (no `blocks` full-size accumulations)."""

    @staticmethod
    def forward(ctx, qkv, pos_all, u, v, length, B, T, N, H, chunk, history, blk, pack):
        NH = N * H
        out, probs = kernels.rel_attention_fwd(qkv, pos_all[:, blk * NH:(blk + 1) * NH], u.detach(), v.detach(), length, B, T, N, H, chunk, history)
        ctx.save_for_backward(qkv, pos_all, probs, length, out)
        ctx.misc = (u, v, B, T, N, H, chunk, history, blk, pack)
        return out

    @staticmethod
    def backward(ctx, dctx):
        qkv, pos_all, probs, length, out = ctx.saved_tensors
        u, v, B, T, N, H, chunk, history, blk, pack = ctx.misc
        NH = N * H
        if pack.dpos is None or pack.dpos.shape != pos_all.shape:
            pack.dpos = torch.zeros_like(pos_all)  # zeros: a backward pass that skips some block must not leave garbage in its slice
        sl = slice(blk * NH, (blk + 1) * NH)
        dqkv, _ = kernels.rel_attention_bwd(qkv, pos_all[:, sl], u.detach(), v.detach(), length, B, T, N, H, probs, dctx.contiguous(), _grad(u),
                                            _grad(v), chunk, history, ctx=out, dpos_out=pack.dpos[:, sl])
        g = None
        if blk == 0:
            g, pack.dpos = pack.dpos, None
        return dqkv, g, None, None, None, None, None, None, None, None, None, None, None


def rel_attention_slice(qkv, pos_all, blk, pack, u, v, length, B, T, N, H, chunk=None, history=None):
    return _RelAttentionSliceFn.apply(qkv, pos_all, u, v, length, B, T, N, H, chunk, history, blk, pack)


# ------------------------------------------------------------------------------------------------ ConvModule
class GluPack(WeightPack):
    """pw_conv_1 [D, 2D] with its GLU halves interleaved column-wise (a_0, b_0, a_1, b_1, ...) so that the GEMM epilogue applies
    a * sigmoid(b) on adjacent columns; gradients are de-interleaved back into the Keras layout."""

    def __init__(self, param, bias, plan: PackPlan = None, unplan: UnpackPlan = None):
        super().__init__(param, bias, interleave_glu=True, plan=plan)
        K, D2 = self.shape2d
        self.D = D2 // 2
        self.unplan = unplan
        if unplan is not None:
            self.dw_packed = torch.zeros((K, D2), dtype=torch.float32, device=param.device)
            self.db_packed = torch.zeros(D2, dtype=torch.float32, device=param.device)
            for half in (0, 1):
                unplan.add(param, K, self.D, self.dw_packed, D2, c0=half, cs=2, grad_ld=D2, grad_col0=half * self.D)
                unplan.add(bias, 1, self.D, self.db_packed, D2, c0=half, cs=2, grad_ld=D2, grad_col0=half * self.D)
        if plan is not None:
            self._bias = torch.empty(D2, dtype=torch.float32, device=param.device)
            plan.add(bias, 1, self.D, self._bias, None, ldw=D2, c0=0, cs=2, src_ld=D2, src_col0=0, f32=True)
            plan.add(bias, 1, self.D, self._bias, None, ldw=D2, c0=1, cs=2, src_ld=D2, src_col0=self.D, f32=True)
        else:
            b = bias.detach().float()
            self._bias = torch.stack([b[: self.D], b[self.D:]], dim=1).reshape(D2).contiguous()
        self._db = None

    def bias_f32(self):
        return self._bias

    def deinterleave_add(self, dw_inter, db_inter):
        K = dw_inter.shape[0]
        g = _grad(self.param)
        d3 = dw_inter.view(K, self.D, 2)
        g[:, : self.D].add_(d3[:, :, 0])
        g[:, self.D:].add_(d3[:, :, 1])
        gb = _grad(self.bias)
        b2 = db_inter.view(self.D, 2)
        gb[: self.D].add_(b2[:, 0])
        gb[self.D:].add_(b2[:, 1])


class _ConvModuleFn(torch.autograd.Function):
    """x + factor * Drop(PW2 . swish(BN(DWconv_causal(GLU(PW1 . LN(x))))))     (conformer.py:366-377), batch-norm variant."""

    @staticmethod
    def forward(ctx, x, ln_g, ln_b, p1: GluPack, dw_w, dw_b, bn, p2: WeightPack, B, T, factor, drop_p, salt, seed, training):
        M, D = x.shape
        a, mean, rstd = kernels.layernorm_fwd(x, ln_g, ln_b, LN_EPS)
        aux = torch.empty((M, 2 * D), dtype=torch.bfloat16, device=x.device)
        g = kernels.gemm(a, p1.wt, bias=p1.bias_f32(), glu=True, aux=aux)  # [M, D]
        w32 = dw_w.detach().float().contiguous()
        c = kernels.dwconv_fwd(g.view(B, T, D), w32, dw_b.detach().float()).view(M, D)
        gamma, beta, mm, mv = bn
        if training:
            y, saved = bn_train_fwd(c, gamma, beta, mm, mv)
        else:
            y, saved = bn_eval(c, gamma, beta, mm, mv), None
        out = kernels.gemm(y, p2.wt, bias=p2.bias_f32(), residual=x, residual_scale=factor, dropout_p=drop_p, dropout_seed=seed, dropout_salt=salt)
        ctx.save_for_backward(x, mean, rstd, a, aux, g, c, y, w32)
        ctx.misc = (ln_g, ln_b, p1, dw_w, dw_b, bn, p2, B, T, factor, drop_p, salt, seed, saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mean, rstd, a, aux, g, c, y, w32 = ctx.saved_tensors
        ln_g, ln_b, p1, dw_w, dw_b, bn, p2, B, T, factor, drop_p, salt, seed, saved = ctx.misc
        if saved is None:
            raise RuntimeError("ConvModule fused backward needs training-mode batch statistics")
        M, D = x.shape
        dout = dout.contiguous()
        dy = kernels.scale_mask_colsum(dout, scale=factor, dropout_p=drop_p, seed=seed, salt=salt, colsum=p2.bias_grad_target())
        p2.accumulate_wgrad(y, dy)
        dyb = kernels.gemm(dy, p2.w)
        dc = bn_train_bwd(dyb, c, saved, bn[0], bn[1])
        # the depthwise bias sits in front of a batch norm: its gradient is identically zero
        dg = kernels.dwconv_bwd(dc.view(B, T, D), g.view(B, T, D), w32, _grad(dw_w).view(-1, D), None).view(M, D)
        if getattr(p1, "unplan", None) is not None:
            daux = kernels.glu_bwd(aux, dg, colsum=p1.db_packed)
            _wgrad(a, daux, p1.dw_packed, _splits(M))
            p1.unplan.request()
        else:
            db1 = torch.zeros(2 * D, dtype=torch.float32, device=x.device)
            daux = kernels.glu_bwd(aux, dg, colsum=db1)
            dw1 = torch.zeros((D, 2 * D), dtype=torch.float32, device=x.device)
            kernels.gemm(a, daux, mn_major=True, out=dw1, accumulate=True, k_splits=_splits(M))
            p1.deinterleave_add(dw1, db1)
        da = kernels.gemm(daux, p1.w)
        dx = kernels.layernorm_bwd(da, x, mean, rstd, ln_g, _grad(ln_g), _grad(ln_b), dres=dout)
        return (dx,) + (None,) * 14


def conv_module(x, ln_g, ln_b, p1, dw_w, dw_b, bn, p2, B, T, factor, drop_p, salt, seed, training):
    return _ConvModuleFn.apply(x, ln_g, ln_b, p1, dw_w, dw_b, bn, p2, B, T, float(factor), float(drop_p), int(salt), seed, bool(training))
