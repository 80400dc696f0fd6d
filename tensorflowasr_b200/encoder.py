"""ConformerEncoder — host-side mirror of tensorflow_asr/models/encoders/conformer.py on B200.

Same constructor arguments and call contract as the reference (`ConformerEncoder.call((features, length), training)
-> (outputs, length)`, conformer.py:539-718); parameters carry the reference's Keras names
(`subsampling.conv_0.kernel`, `block_3.mhsa_module.mhsa.query.kernel`, ...) so that reference checkpoints map 1:1.
The arithmetic is dispatched op by op through `tensorflowasr_b200.ops`, which routes to the hand-written sm_100a
kernels behind the C ABI.
"""
from __future__ import annotations

import math
import zlib

import torch
from torch import nn

from . import fused, ops


def conv_output_length(input_length, filter_size, padding, stride, dilation=1):
    """utils/math_util.py:282-305."""
    if input_length is None:
        return None
    assert padding in {"same", "valid", "full", "causal"}
    dilated = filter_size + (filter_size - 1) * (dilation - 1)
    if padding in ("same", "causal"):
        out = input_length
    elif padding == "valid":
        out = input_length - dilated + 1
    else:
        out = input_length + dilated - 1
    return (out + stride - 1) // stride


def _glorot_(t: torch.Tensor, fan_in: int, fan_out: int, gen=None):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        t.uniform_(-lim, lim, generator=gen)
    return t


class _Params(nn.Module):
    """Flat, Keras-named parameter store: `self.p['block_0.ln.gamma']`."""

    def __init__(self):
        super().__init__()
        self._names = {}

    def add(self, name: str, shape, kind="zeros", fan=None, trainable=True, gen=None):
        t = torch.zeros(*shape)
        if kind == "ones":
            t.fill_(1.0)
        elif kind == "glorot":
            _glorot_(t, fan[0], fan[1], gen)
        elif kind == "uniform":
            with torch.no_grad():
                t.uniform_(-0.05, 0.05, generator=gen)
        key = name.replace(".", "__")
        if trainable:
            self.register_parameter(key, nn.Parameter(t))
        else:
            self.register_buffer(key, t)
        self._names[name] = key

    def __getitem__(self, name: str) -> torch.Tensor:
        return getattr(self, self._names[name])

    def names(self):
        return list(self._names)

    def named(self):
        return {n: getattr(self, k) for n, k in self._names.items()}

    def load_named(self, tensors: dict, prefix: str = "", strict=True):
        """Copy tensors keyed by reference names (optionally under `prefix.`) into this store."""
        with torch.no_grad():
            for n, k in self._names.items():
                src = tensors.get(prefix + n)
                if src is None:
                    if strict:
                        raise KeyError(prefix + n)
                    continue
                getattr(self, k).copy_(torch.as_tensor(src).reshape(getattr(self, k).shape))


class Conv2dSubsampling(_Params):
    """models/layers/subsampling.py:163-256."""

    def __init__(self, filters, strides=((2, 1), (2, 1)), kernels=((3, 3), (3, 3)), paddings=("causal", "causal"),
                 norms=("none", "none"), activations=("relu", "relu"), input_channels=1, name="conv2d_subsampling", gen=None, **kwargs):
        super().__init__()
        n = len(filters)

        def pair(v):
            return (v, v) if isinstance(v, int) else tuple(v)

        self.filters = list(filters)
        self.strides = [pair(s) for s in strides]
        self.kernels = [pair(k) for k in kernels]
        self.paddings = list(paddings)
        self.norms = list(norms)
        self.activations = list(activations)
        assert len(self.strides) == len(self.kernels) == len(self.paddings) == len(self.norms) == len(self.activations) == n
        self.time_reduction_factor = 1
        cin = input_channels
        for i in range(n):
            kh, kw = self.kernels[i]
            self.add(f"conv_{i}.kernel", (kh, kw, cin, filters[i]), "glorot", (cin * kh * kw, filters[i] * kh * kw), gen=gen)
            self.add(f"conv_{i}.bias", (filters[i],))
            if self.norms[i] == "batch":
                self.add(f"bn_{i}.gamma", (filters[i],), "ones")
                self.add(f"bn_{i}.beta", (filters[i],))
                self.add(f"bn_{i}.moving_mean", (filters[i],), trainable=False)
                self.add(f"bn_{i}.moving_variance", (filters[i],), "ones", trainable=False)
            elif self.norms[i] == "layer":
                self.add(f"ln_{i}.gamma", (filters[i],), "ones")
                self.add(f"ln_{i}.beta", (filters[i],))
            cin = filters[i]
            self.time_reduction_factor *= self.strides[i][0]

    def output_feature_dim(self, nfeat):
        for i in range(len(self.filters)):
            nfeat = conv_output_length(nfeat, self.kernels[i][1], self.paddings[i], self.strides[i][1])
        return nfeat * self.filters[-1]

    def _fusable(self):
        return (len(self.filters) == 2 and self.filters[0] == self.filters[1] and self.filters[0] % 8 == 0 and self.norms == ["batch", "batch"]
                and self.activations == ["swish", "swish"] and self.paddings == ["causal", "causal"]
                and self.kernels == [(3, 3), (3, 3)] and self.strides == [(2, 2), (2, 2)])

    def forward(self, inputs, training=False):
        x, length = inputs  # x [B, T, F, C]
        if ops.fused_active(x) and x.shape[-1] == 1 and self._fusable():
            y = fused.conv2d_subsampling(x, self, training)
            for i in range(2):
                length = conv_output_length(length, 3, "causal", 2)
            return y, length
        for i in range(len(self.filters)):
            x = ops.conv2d_block(
                x, self[f"conv_{i}.kernel"], self[f"conv_{i}.bias"], stride=self.strides[i], padding=self.paddings[i],
                norm=self.norms[i], activation=self.activations[i], training=training,
                norm_params=self._norm_params(i),
            )
            length = conv_output_length(length, self.kernels[i][0], self.paddings[i], self.strides[i][0])
        B, T, F_, C = x.shape
        return x.reshape(B, T, F_ * C), length  # merge_two_last_dims

    def _norm_params(self, i):
        if self.norms[i] == "batch":
            return (self[f"bn_{i}.gamma"], self[f"bn_{i}.beta"], self[f"bn_{i}.moving_mean"], self[f"bn_{i}.moving_variance"])
        if self.norms[i] == "layer":
            return (self[f"ln_{i}.gamma"], self[f"ln_{i}.beta"])
        return None


class ConformerEncoder(_Params):
    """models/encoders/conformer.py:539-718 (relmha, pre-norm modules, post-norm block — the shipped configuration)."""

    def __init__(
        self,
        subsampling,
        dmodel=144,
        num_blocks=16,
        mha_type="relmha",
        head_size=36,
        num_heads=4,
        kernel_size=32,
        padding="causal",
        interleave_relpe=True,
        use_attention_causal_mask=False,
        use_attention_auto_mask=True,
        ffm_scale_factor=4,
        ffm_residual_factor=0.5,
        mhsam_residual_factor=1.0,
        mhsam_use_attention_bias=False,
        mhsam_causal=False,
        mhsam_flash_attention=None,
        convm_scale_factor=2,
        convm_residual_factor=1.0,
        convm_use_group_conv=False,
        convm_dw_norm_type="batch",
        dropout=0.1,
        module_norm_position="pre",
        block_norm_position="post",
        memory_length=None,
        history_size=None,
        chunk_size=None,
        kernel_regularizer=None,
        bias_regularizer=None,
        activity_regularizer=None,
        name="conformer_encoder",
        num_feature_bins=80,
        gen=None,
        **kwargs,
    ):
        super().__init__()
        if mha_type != "relmha" or mhsam_causal or mhsam_use_attention_bias or memory_length:
            raise NotImplementedError("B200 hot path covers relmha, non-causal, shared attention biases, no KV memory (SURVEY.md §8a)")
        if module_norm_position != "pre" or block_norm_position != "post" or padding != "causal" or not interleave_relpe:
            raise NotImplementedError("B200 hot path covers the shipped Conformer configuration (pre/post norms, causal conv, interleaved relpe)")
        if convm_scale_factor != 2 or convm_use_group_conv or use_attention_causal_mask:
            raise NotImplementedError("unsupported ConvModule / mask option on the B200 hot path")
        cls = subsampling["class_name"].split(">")[-1]
        if cls != "Conv2dSubsampling":
            raise NotImplementedError(f"subsampling {cls} is outside the B200 hot path (Conv2dSubsampling only)")
        self.name = name
        self.dmodel, self.num_blocks, self.head_size, self.num_heads = dmodel, num_blocks, head_size, num_heads
        self.kernel_size = kernel_size
        self.ffm_scale_factor = ffm_scale_factor
        self.ffm_residual_factor = ffm_residual_factor
        self.mhsam_residual_factor = mhsam_residual_factor
        self.convm_residual_factor = convm_residual_factor
        self.dw_norm_type = convm_dw_norm_type
        self.dropout = dropout
        self.history_size, self.chunk_size = history_size, chunk_size
        self.conv_subsampling = Conv2dSubsampling(**subsampling["config"], gen=gen)
        self.time_reduction_factor = self.conv_subsampling.time_reduction_factor
        D, N, H = dmodel, num_heads, head_size
        fin = self.conv_subsampling.output_feature_dim(num_feature_bins)
        self.add("linear.kernel", (fin, D), "glorot", (fin, D), gen=gen)
        self.add("linear.bias", (D,))
        self.add("content_attention_bias", (N, H))
        self.add("positional_attention_bias", (N, H))
        for b in range(num_blocks):
            pre = f"block_{b}"
            for ff in ("ff_module_1", "ff_module_2"):
                self.add(f"{pre}.{ff}.ln.gamma", (D,), "ones")
                self.add(f"{pre}.{ff}.ln.beta", (D,))
                self.add(f"{pre}.{ff}.dense_1.kernel", (D, ffm_scale_factor * D), "glorot", (D, ffm_scale_factor * D), gen=gen)
                self.add(f"{pre}.{ff}.dense_1.bias", (ffm_scale_factor * D,))
                self.add(f"{pre}.{ff}.dense_2.kernel", (ffm_scale_factor * D, D), "glorot", (ffm_scale_factor * D, D), gen=gen)
                self.add(f"{pre}.{ff}.dense_2.bias", (D,))
            self.add(f"{pre}.mhsa_module.ln.gamma", (D,), "ones")
            self.add(f"{pre}.mhsa_module.ln.beta", (D,))
            for nm in ("query", "key", "value", "encoding"):
                self.add(f"{pre}.mhsa_module.mhsa.{nm}.kernel", (D, N, H), "glorot", (D, N * H), gen=gen)
                self.add(f"{pre}.mhsa_module.mhsa.{nm}.bias", (N, H))
            self.add(f"{pre}.mhsa_module.mhsa.attention_output.kernel", (N, H, D), "glorot", (N * H, D), gen=gen)
            self.add(f"{pre}.mhsa_module.mhsa.attention_output.bias", (D,))
            self.add(f"{pre}.conv_module.ln.gamma", (D,), "ones")
            self.add(f"{pre}.conv_module.ln.beta", (D,))
            self.add(f"{pre}.conv_module.pw_conv_1.kernel", (D, 2 * D), "glorot", (D, 2 * D), gen=gen)
            self.add(f"{pre}.conv_module.pw_conv_1.bias", (2 * D,))
            self.add(f"{pre}.conv_module.dw_conv.kernel", (kernel_size, D), "glorot", (kernel_size * D, kernel_size), gen=gen)
            self.add(f"{pre}.conv_module.dw_conv.bias", (D,))
            if convm_dw_norm_type == "batch":
                self.add(f"{pre}.conv_module.dw_bn.gamma", (D,), "ones")
                self.add(f"{pre}.conv_module.dw_bn.beta", (D,))
                self.add(f"{pre}.conv_module.dw_bn.moving_mean", (D,), trainable=False)
                self.add(f"{pre}.conv_module.dw_bn.moving_variance", (D,), "ones", trainable=False)
            else:
                self.add(f"{pre}.conv_module.dw_ln.gamma", (D,), "ones")
                self.add(f"{pre}.conv_module.dw_ln.beta", (D,))
            self.add(f"{pre}.conv_module.pw_conv_2.kernel", (D, D), "glorot", (D, D), gen=gen)
            self.add(f"{pre}.conv_module.pw_conv_2.bias", (D,))
            self.add(f"{pre}.ln.gamma", (D,), "ones")
            self.add(f"{pre}.ln.beta", (D,))

    # ------------------------------------------------------------------------------------------
    def _dw_norm(self, pre):
        if self.dw_norm_type == "batch":
            return ("batch", self[f"{pre}.conv_module.dw_bn.gamma"], self[f"{pre}.conv_module.dw_bn.beta"],
                    self[f"{pre}.conv_module.dw_bn.moving_mean"], self[f"{pre}.conv_module.dw_bn.moving_variance"])
        return ("layer", self[f"{pre}.conv_module.dw_ln.gamma"], self[f"{pre}.conv_module.dw_ln.beta"])

    # ------------------------------------------------------------------------------------------ fused (sm_100a kernel) path
    def _prepare_packs(self):
        """bf16 operand copies of the GEMM weights for this forward/backward (weights change every optimizer step): persistent
        buffers, refreshed by one batched launch (fused.PackPlan)."""
        batched = ops.pos_batched_active()
        if getattr(self, "_packs", None) is None or self._pack_plan is None or getattr(self, "_packs_pos_batched", None) != batched:
            self._packs_pos_batched = batched
            plan = fused.PackPlan()
            unplan = fused.UnpackPlan()
            packs = {}
            for b in range(self.num_blocks):
                for ff in ("ff_module_1", "ff_module_2"):
                    for d in ("dense_1", "dense_2"):
                        n = f"block_{b}.{ff}.{d}"
                        packs[n] = fused.WeightPack(self[n + ".kernel"], self[n + ".bias"], plan=plan)
                packs[f"block_{b}.mhsa_module"] = fused.MhsaPack(self, f"block_{b}.mhsa_module", plan, unplan, with_pos=not batched)
                cm = f"block_{b}.conv_module"
                packs[cm + ".pw_conv_1"] = fused.GluPack(self[cm + ".pw_conv_1.kernel"], self[cm + ".pw_conv_1.bias"], plan, unplan)
                packs[cm + ".pw_conv_2"] = fused.WeightPack(self[cm + ".pw_conv_2.kernel"], self[cm + ".pw_conv_2.bias"], plan=plan)
            if batched:
                packs["pos_all"] = fused.PosAllPack(self, plan, unplan)
            self._packs, self._pack_plan, self._unpack_plan = packs, plan, unplan
        self._pack_plan.run()

    def _salt(self, pre, k):
        return (zlib.crc32(pre.encode()) % 1000003) * 64 + k * 2 + 1

    def _ff(self, x, pre, training):
        # FFModule.call, conformer.py:101-109
        if ops.fused_active(x):
            B, T, D = x.shape
            p = self.dropout if training else 0.0
            blk = pre[len(self.name) + 1:] if pre.startswith(self.name + ".") else pre
            y = fused.ff_module(x.reshape(B * T, D).to(torch.bfloat16), self[pre + ".ln.gamma"], self[pre + ".ln.beta"],
                                self._packs[blk + ".dense_1"], self._packs[blk + ".dense_2"], self.ffm_residual_factor, p,
                                self._salt(pre, 0), fused.STATE.seed_tensor(x.device))
            return y.view(B, T, D)
        y = ops.layer_norm(x, self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
        y = ops.linear(y, self[pre + ".dense_1.kernel"], self[pre + ".dense_1.bias"], activation="swish")
        y = ops.dropout(y, self.dropout, training)
        y = ops.linear(y, self[pre + ".dense_2.kernel"], self[pre + ".dense_2.bias"])
        y = ops.dropout(y, self.dropout, training)
        return ops.residual(x, y, self.ffm_residual_factor)

    def _mhsa(self, x, relpe, length, pre, training, pos_all=None, blk=0):
        # MHSAModule.call, conformer.py:209-239 -> MultiHeadRelativeAttention.call, multihead_attention.py:584-667
        D, N, H = self.dmodel, self.num_heads, self.head_size
        m = pre + ".mhsa"
        if ops.fused_active(x):
            # LN, the fused Q|K|V projection, the position projection and the output projection (+ dropout + residual) run on the
            # tcgen05 GEMM / fused row kernels; the attention core on the mma.sync kernels of csrc/attention.cu.
            B, T, _ = x.shape
            NH = N * H
            p = self.dropout if training else 0.0
            pk = self._packs[pre]
            seed = fused.STATE.seed_tensor(x.device)
            x2 = x.reshape(B * T, D).to(torch.bfloat16)
            a = fused.layer_norm(x2, self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
            qkv = fused.linear(a, pk.qkv)
            if pos_all is not None:
                ctx = fused.rel_attention_slice(qkv, pos_all, blk, self._packs["pos_all"], self["content_attention_bias"],
                                                self["positional_attention_bias"], length.to(torch.int32), B, T, N, H, self.chunk_size,
                                                self.history_size)
            else:
                pos = fused.linear(relpe if relpe.dtype == torch.bfloat16 else relpe.reshape(-1, D).to(torch.bfloat16).contiguous(), pk.pos)
                ctx = fused.rel_attention(qkv, pos, self["content_attention_bias"], self["positional_attention_bias"], length.to(torch.int32),
                                          B, T, N, H, self.chunk_size, self.history_size)
            y = fused.linear(ctx, pk.out, residual=x2, factor=self.mhsam_residual_factor, drop_p=p,
                             salt=self._salt(pre, 0), seed=seed)
            return y.view(B, T, D)
        y = ops.layer_norm(x, self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
        q = ops.linear(y, self[m + ".query.kernel"].reshape(D, N * H), self[m + ".query.bias"].reshape(N * H))
        k = ops.linear(y, self[m + ".key.kernel"].reshape(D, N * H), self[m + ".key.bias"].reshape(N * H))
        v = ops.linear(y, self[m + ".value.kernel"].reshape(D, N * H), self[m + ".value.bias"].reshape(N * H))
        pos = ops.linear(relpe, self[m + ".encoding.kernel"].reshape(D, N * H), self[m + ".encoding.bias"].reshape(N * H))
        ctx = ops.rel_attention(q, k, v, pos, self["content_attention_bias"], self["positional_attention_bias"], length, N, H,
                                chunk_size=self.chunk_size, history_size=self.history_size)
        y = ops.linear(ctx, self[m + ".attention_output.kernel"].reshape(N * H, D), self[m + ".attention_output.bias"])
        y = ops.dropout(y, self.dropout, training)
        return ops.residual(x, y, self.mhsam_residual_factor)

    def _conv(self, x, pre, training):
        # ConvModule.call, conformer.py:366-377
        if ops.fused_active(x) and self.dw_norm_type == "batch" and (training or not torch.is_grad_enabled()):
            B, T, D = x.shape
            p = self.dropout if training else 0.0
            bn = (self[pre + ".dw_bn.gamma"], self[pre + ".dw_bn.beta"], self[pre + ".dw_bn.moving_mean"], self[pre + ".dw_bn.moving_variance"])
            y = fused.conv_module(x.reshape(B * T, D).to(torch.bfloat16), self[pre + ".ln.gamma"], self[pre + ".ln.beta"],
                                  self._packs[pre + ".pw_conv_1"], self[pre + ".dw_conv.kernel"], self[pre + ".dw_conv.bias"], bn,
                                  self._packs[pre + ".pw_conv_2"], B, T, self.convm_residual_factor, p, self._salt(pre, 1),
                                  fused.STATE.seed_tensor(x.device), training)
            return y.view(B, T, D)
        y = ops.layer_norm(x, self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
        y = ops.linear(y, self[pre + ".pw_conv_1.kernel"], self[pre + ".pw_conv_1.bias"])
        y = ops.glu_dwconv_norm_swish(y, self[pre + ".dw_conv.kernel"], self[pre + ".dw_conv.bias"], self._dw_norm(pre.rsplit(".", 1)[0]),
                                      training)
        y = ops.linear(y, self[pre + ".pw_conv_2.kernel"], self[pre + ".pw_conv_2.bias"])
        y = ops.dropout(y, self.dropout, training)
        return ops.residual(x, y, self.convm_residual_factor)

    def forward(self, inputs, initial_state=None, training=False, return_states=False):
        return self.call(inputs, initial_state=initial_state, training=training, return_states=return_states)

    def call(self, inputs, initial_state=None, training=False, return_states=False):
        x, length = inputs  # features [B, T, F, 1]
        if ops.fused_active(x):
            self._prepare_packs()
        x, length = self.conv_subsampling((x, length), training=training)
        x = ops.linear(x, self["linear.kernel"], self["linear.bias"])
        x = ops.dropout(x, self.dropout, training)
        relpe = ops.relative_position_encoding(x.shape[1], length, self.dmodel, x.device)  # not added to x (positional_encoding.py:174)
        relpe = ops.dropout(relpe, 0.0, training)
        if ops.fused_active(x):
            relpe = relpe.reshape(-1, self.dmodel).to(torch.bfloat16).contiguous()  # GEMM operand of every block's position projection: cast once
        pos_all = None
        if ops.fused_active(x) and self._packs_pos_batched:
            pos_all = fused.linear(relpe, self._packs["pos_all"])  # [B*R, blocks*NH]: every block's position projection in one GEMM
        for b in range(self.num_blocks):
            pre = f"block_{b}"
            x = self._ff(x, pre + ".ff_module_1", training)
            x = self._mhsa(x, relpe, length, pre + ".mhsa_module", training, pos_all=pos_all, blk=b)
            x = self._conv(x, pre + ".conv_module", training)
            x = self._ff(x, pre + ".ff_module_2", training)
            if ops.fused_active(x):
                x = fused.layer_norm(x.contiguous(), self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
            else:
                x = ops.layer_norm(x, self[pre + ".ln.gamma"], self[pre + ".ln.beta"])
        if return_states:
            return x, length, None
        return x, length

    def call_next(self, features, features_length, previous_encoder_states, *args, **kwargs):
        return self.call((features, features_length), initial_state=previous_encoder_states, training=False, return_states=True)

    def get_initial_state(self, batch_size: int):
        return []
