"""CtcModel / DeepSpeech2 and the ContextNet transducer — host-side mirrors of
tensorflow_asr/models/ctc/{base_ctc,deepspeech2}.py, models/encoders/{deepspeech2,contextnet}.py and
models/transducer/contextnet.py.

The frontend (log-mel / spectrogram) and the CTC / RNN-T losses run on the hand-written kernels behind the C ABI.  The encoder bodies
(§8a rows A22 / A23) have two routes: the LIBRARY route below (cuDNN / cuBLAS through torch; the fp32 parity route and the default
in this round) and the fused route of fused_seq.py (own depthwise-conv / squeeze-excite / Conv2D-lowering kernels + tcgen05 GEMM +
BatchNorm kernels + cluster LSTM), taken in bf16 mode when `ops.set_fused_seq(True)` is set.  Parameter names follow the Keras layers.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import fused, fused_seq, kernels, ops, schemas
from .base_model import BaseModel
from .encoder import _Params, conv_output_length
from .losses import CtcLoss
from .transducer import Transducer


def _same_pad(n, k, s):
    """Keras/TF 'same' padding: total = max((ceil(n/s) - 1) * s + k - n, 0); before = total // 2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


# ------------------------------------------------------------------------------------------------ DeepSpeech2
class DeepSpeech2Encoder(_Params):
    """models/encoders/deepspeech2.py:416-533: ConvModule (Conv2D + BN + act) -> RnnModule (BiLSTM stack, dropout) -> FcModule."""

    def __init__(self, conv_type="conv2d", conv_kernels=((11, 41), (11, 21), (11, 21)), conv_strides=((2, 2), (1, 2), (1, 2)),
                 conv_filters=(32, 32, 96), conv_padding="same", conv_activation="relu", conv_initializer=None, rnn_nlayers=5,
                 rnn_type="lstm", rnn_units=1024, rnn_bidirectional=True, rnn_unroll=False, rnn_rowconv=0, rnn_rowconv_activation="relu",
                 rnn_dropout=0.1, fc_nlayers=0, fc_units=1024, fc_activation="relu", fc_dropout=0.1, fc_initializer=None,
                 num_feature_bins=160, name="deepspeech2_encoder", gen=None, **kwargs):
        super().__init__()
        if conv_type != "conv2d" or rnn_type != "lstm" or not rnn_bidirectional or rnn_rowconv:
            raise NotImplementedError("B200 path covers the shipped DeepSpeech2: conv2d + bidirectional LSTM (deepspeech2/base.yml.j2)")
        self.conv_kernels = [tuple(k) for k in conv_kernels]
        self.conv_strides = [tuple(s) for s in conv_strides]
        self.conv_filters = list(conv_filters)
        self.conv_padding, self.conv_activation = conv_padding, conv_activation
        self.rnn_nlayers, self.rnn_units, self.rnn_dropout = rnn_nlayers, rnn_units, rnn_dropout
        self.fc_nlayers, self.fc_units, self.fc_activation, self.fc_dropout = fc_nlayers, fc_units, fc_activation, fc_dropout
        cin, f = 1, num_feature_bins
        self.time_reduction_factor = 1
        for i, (k, s, co) in enumerate(zip(self.conv_kernels, self.conv_strides, self.conv_filters)):
            pre = f"conv_module.block_{i}"
            self.add(f"{pre}.conv2d.kernel", (k[0], k[1], cin, co), "glorot", (cin * k[0] * k[1], co * k[0] * k[1]), gen=gen)
            self.add(f"{pre}.conv2d.bias", (co,))
            self.add(f"{pre}.bn.gamma", (co,), "ones")
            self.add(f"{pre}.bn.beta", (co,))
            self.add(f"{pre}.bn.moving_mean", (co,), trainable=False)
            self.add(f"{pre}.bn.moving_variance", (co,), "ones", trainable=False)
            cin = co
            f = conv_output_length(f, k[1], conv_padding, s[1])
            self.time_reduction_factor *= s[0]
        din = f * cin
        for i in range(rnn_nlayers):
            for d in ("forward", "backward"):
                pre = f"rnn_module.block_{i}.blstm.{d}_lstm"
                self.add(f"{pre}.kernel", (din, 4 * rnn_units), "glorot", (din, 4 * rnn_units), gen=gen)
                self.add(f"{pre}.recurrent_kernel", (rnn_units, 4 * rnn_units), "glorot", (rnn_units, 4 * rnn_units), gen=gen)
                self.add(f"{pre}.bias", (4 * rnn_units,))
                with torch.no_grad():
                    self[f"{pre}.bias"][rnn_units: 2 * rnn_units] = 1.0
            din = 2 * rnn_units
        for i in range(fc_nlayers):
            self.add(f"fc_module.block_{i}.fc.kernel", (din, fc_units), "glorot", (din, fc_units), gen=gen)
            self.add(f"fc_module.block_{i}.fc.bias", (fc_units,))
            din = fc_units
        self.dmodel = din

    def _bilstm(self, x, length, i):
        """Keras Bidirectional(LSTM(zero_output_for_mask=True)): the backward direction runs over each utterance's valid frames in
        reverse — cuDNN packed-sequence semantics.  x [B,T,D] -> [B,T,2R].  LIBRARY (cuDNN)"""
        B, T, _ = x.shape
        R = self.rnn_units
        ws = []
        for d in ("forward", "backward"):
            pre = f"rnn_module.block_{i}.blstm.{d}_lstm"
            ws += [self[f"{pre}.kernel"].float().t().contiguous(), self[f"{pre}.recurrent_kernel"].float().t().contiguous(),
                   self[f"{pre}.bias"].float(), torch.zeros_like(self[f"{pre}.bias"]).float()]
        lens = torch.as_tensor(length).to("cpu", torch.int64).clamp(min=1, max=T)
        packed = torch.nn.utils.rnn.pack_padded_sequence(x.float(), lens, batch_first=True, enforce_sorted=False)
        h0 = x.new_zeros(2, B, R, dtype=torch.float32)
        train = torch.is_grad_enabled()
        out, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes, (h0, h0.clone()), ws, True, 1, 0.0, train, True)
        out = torch.nn.utils.rnn.PackedSequence(out, packed.batch_sizes, packed.sorted_indices, packed.unsorted_indices)
        y, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=T)
        return y.to(ops.compute_dtype())

    def forward(self, inputs, training=False):
        return self.call(inputs, training=training)

    # ---- fused route (fused_seq.py)
    def _fusable(self, training):
        return (self.conv_padding in ("same", "causal") and self.conv_activation in ("relu", "swish", "silu", "linear", None)
                and all(c % 8 == 0 for c in self.conv_filters) and self.fc_activation in ("relu", None)
                and (training or not torch.is_grad_enabled()))

    def _prepare_packs(self):
        if getattr(self, "_packs", None) is None:
            plan, packs = fused.PackPlan(), {}
            for i in range(len(self.conv_kernels)):
                pre = f"conv_module.block_{i}"
                packs[pre] = fused_seq.ConvPack(self[f"{pre}.conv2d.kernel"], self[f"{pre}.conv2d.bias"], plan)
            for i in range(self.fc_nlayers):
                pre = f"fc_module.block_{i}.fc"
                packs[pre] = fused.WeightPack(self[f"{pre}.kernel"], self[f"{pre}.bias"], plan=plan)
            self._packs, self._pack_plan = packs, plan
        self._pack_plan.run()

    def _call_fused(self, x, length, training):
        """conv module: im2col -> tcgen05 GEMM -> BN + activation kernels; BiLSTM: input projections on the GEMM, recurrence on the cluster
        LSTM kernels when rnn_units <= 384 (else cuDNN, LIBRARY); fc module: GEMM with bias + relu + dropout in the epilogue."""
        self._prepare_packs()
        act = kernels.act_code(self.conv_activation)
        y = x.float().contiguous()  # [B,T,F,1]
        for i, (k, s) in enumerate(zip(self.conv_kernels, self.conv_strides)):
            pre = f"conv_module.block_{i}"
            bn = (self[f"{pre}.bn.gamma"], self[f"{pre}.bn.beta"], self[f"{pre}.bn.moving_mean"], self[f"{pre}.bn.moving_variance"])
            y = fused_seq.conv2d_bn(y, self._packs[pre], bn, k, s, self.conv_padding, act, training)
            length = conv_output_length(length, k[0], self.conv_padding, s[0])
        B, T, Fq, C = y.shape
        y = y.reshape(B, T, Fq * C)
        len32 = torch.as_tensor(length, device=y.device).to(torch.int32).contiguous()
        seed = fused.STATE.seed_tensor(y.device)
        own_lstm = kernels.lstm_supported(self.rnn_units)
        for i in range(self.rnn_nlayers):
            if own_lstm:
                ws = [tuple(self[f"rnn_module.block_{i}.blstm.{d}_lstm.{n}"] for n in ("kernel", "recurrent_kernel", "bias")) for d in ("forward", "backward")]
                y = fused_seq.bilstm(y, len32, ws[0], ws[1])
            else:
                y = self._bilstm(y, length, i).to(torch.bfloat16)
            y = ops.dropout(y, self.rnn_dropout, training)
        y2 = y.reshape(B * T, -1).contiguous()
        for i in range(self.fc_nlayers):
            p = self.fc_dropout if training else 0.0
            y2 = fused_seq.dense_act_drop(y2, self._packs[f"fc_module.block_{i}.fc"], self.fc_activation, p, salt=0xD52 + i, seed=seed)
        return y2.view(B, T, -1), length

    def call(self, inputs, training=False):
        x, length = inputs  # [B,T,F,1]
        if ops.fused_seq_active(x) and self._fusable(training):
            return self._call_fused(x, length, training)
        dt = ops.compute_dtype()
        y = x.to(dt).permute(0, 3, 1, 2)
        for i, (k, s) in enumerate(zip(self.conv_kernels, self.conv_strides)):
            pre = f"conv_module.block_{i}"
            if self.conv_padding == "same":
                pt, pf = _same_pad(y.shape[2], k[0], s[0]), _same_pad(y.shape[3], k[1], s[1])
                y = F.pad(y, (pf[0], pf[1], pt[0], pt[1]))
            elif self.conv_padding == "causal":
                y = F.pad(y, (k[1] - 1, 0, k[0] - 1, 0))
            y = F.conv2d(y, self[f"{pre}.conv2d.kernel"].to(dt).permute(3, 2, 0, 1), self[f"{pre}.conv2d.bias"].to(dt), stride=s)
            y = ops.batch_norm(y.permute(0, 2, 3, 1), self[f"{pre}.bn.gamma"], self[f"{pre}.bn.beta"], self[f"{pre}.bn.moving_mean"],
                               self[f"{pre}.bn.moving_variance"], training)
            y = ops._activation(y, self.conv_activation).permute(0, 3, 1, 2)
            length = conv_output_length(length, k[0], self.conv_padding, s[0])
        y = y.permute(0, 2, 3, 1)
        B, T, Fq, C = y.shape
        y = y.reshape(B, T, Fq * C)
        for i in range(self.rnn_nlayers):
            y = self._bilstm(y, length, i)
            y = ops.dropout(y, self.rnn_dropout, training)
        for i in range(self.fc_nlayers):
            y = ops.linear(y, self[f"fc_module.block_{i}.fc.kernel"], self[f"fc_module.block_{i}.fc.bias"], activation=self.fc_activation)
            y = ops.dropout(y, self.fc_dropout, training)
        return y, length


class CtcDecoder(_Params):
    """The `decoder` of CtcModel for DeepSpeech2: one Dense to the vocabulary (models/ctc/deepspeech2.py:22-54)."""

    def __init__(self, vocab_size, input_dim, gen=None):
        super().__init__()
        self.add("vocab.kernel", (input_dim, vocab_size), "glorot", (input_dim, vocab_size), gen=gen)
        self.add("vocab.bias", (vocab_size,))

    def forward(self, inputs, training=False):
        x, length = inputs
        return ops.linear(x, self["vocab.kernel"], self["vocab.bias"]).float(), length


class CtcModel(BaseModel):
    """models/ctc/base_ctc.py:22-160."""

    def __init__(self, blank, speech_config, encoder, decoder, name="ctc_model", **kwargs):
        super().__init__(speech_config=speech_config, name=name)
        self.blank = blank
        self.encoder = encoder
        self.decoder = decoder
        self.time_reduction_factor = 1

    def compile(self, optimizer, output_shapes=None, **kwargs):
        return super().compile(loss=CtcLoss(blank=self.blank, name="ctc_loss"), optimizer=optimizer, **kwargs)

    def call(self, inputs: schemas.TrainInput, training=False):
        features, features_length = self.feature_extraction((inputs.inputs, inputs.inputs_length), training=training)
        logits, logits_length = self.encoder((features, features_length), training=training)
        logits, logits_length = self.decoder((logits, logits_length), training=training)
        return schemas.TrainOutput(logits=logits, logits_length=torch.as_tensor(logits_length, device=logits.device).to(torch.int32))

    @torch.no_grad()
    def recognize(self, inputs: schemas.PredictInput, **kwargs):
        """tf.nn.ctc_greedy_decoder(merge_repeated=True): arg-max per frame, collapse repeats, drop blanks (base_ctc.py:102-126).
        tokens [B, max_decoded] int32, zero padded (tf.sparse.to_dense)."""
        out = self.call(schemas.TrainInput(inputs.inputs, inputs.inputs_length, None, None), training=False)
        best = out.logits.argmax(-1)  # [B,T]
        B, T = best.shape
        valid = torch.arange(T, device=best.device)[None, :] < out.logits_length[:, None]
        prev = torch.cat([torch.full((B, 1), -1, device=best.device, dtype=best.dtype), best[:, :-1]], 1)
        keep = valid & (best != self.blank) & (best != prev)
        n = keep.sum(1)
        maxn = int(n.max().item()) if B else 0
        tokens = torch.zeros(B, maxn, dtype=torch.int32, device=best.device)
        pos = keep.cumsum(1) - 1
        rows = torch.arange(B, device=best.device)[:, None].expand(B, T)
        tokens[rows[keep], pos[keep]] = best[keep].to(torch.int32)
        return schemas.PredictOutput(tokens=tokens, next_tokens=None, next_encoder_states=None, next_decoder_states=None)

    @torch.no_grad()
    def recognize_beam(self, inputs: schemas.PredictInput, beam_width: int = 10, blank=None, **kwargs):
        """tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width) (base_ctc.py:128-160): the encoder / decoder run on the GPU,
        the prefix beam search on the host like TensorFlow's own CPU op (ctc_decode.py).  `blank=None` keeps the reference call's quirk
        (TensorFlow's default blank = the last class); pass `self.blank` for the index the model trains with.  tokens [B, max_decoded]
        int32, zero padded (tf.sparse.to_dense)."""
        from .ctc_decode import ctc_beam_search

        out = self.call(schemas.TrainInput(inputs.inputs, inputs.inputs_length, None, None), training=False)
        logits = out.logits.float().cpu().numpy()
        lens = out.logits_length.cpu().tolist()
        paths = [ctc_beam_search(logits[b], lens[b], beam_width=beam_width, blank=blank) for b in range(logits.shape[0])]
        maxn = max((len(p) for p in paths), default=0)
        tokens = torch.zeros(len(paths), maxn, dtype=torch.int32, device=out.logits.device)
        for b, p in enumerate(paths):
            if p:
                tokens[b, : len(p)] = torch.tensor(p, dtype=torch.int32)
        return schemas.PredictOutput(tokens=tokens, next_tokens=None, next_encoder_states=None, next_decoder_states=None)

    def _weight_groups(self):
        return {"encoder": self.encoder, "decoder": self.decoder}


class DeepSpeech2(CtcModel):
    """models/ctc/deepspeech2.py:57-140."""

    def __init__(self, blank, vocab_size, speech_config, conv_type="conv2d", conv_kernels=((11, 41), (11, 21), (11, 21)),
                 conv_strides=((2, 2), (1, 2), (1, 2)), conv_filters=(32, 32, 96), conv_padding="same", conv_activation="relu",
                 conv_initializer=None, rnn_nlayers=5, rnn_type="lstm", rnn_units=1024, rnn_bidirectional=True, rnn_unroll=False,
                 rnn_rowconv=0, rnn_rowconv_activation="relu", rnn_dropout=0.1, fc_nlayers=0, fc_units=1024, fc_activation="relu",
                 fc_dropout=0.1, fc_initializer=None, kernel_regularizer=None, bias_regularizer=None, name="deepspeech2", seed=None, **kwargs):
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        enc = DeepSpeech2Encoder(conv_type, conv_kernels, conv_strides, conv_filters, conv_padding, conv_activation, conv_initializer,
                                 rnn_nlayers, rnn_type, rnn_units, rnn_bidirectional, rnn_unroll, rnn_rowconv, rnn_rowconv_activation,
                                 rnn_dropout, fc_nlayers, fc_units, fc_activation, fc_dropout, fc_initializer,
                                 num_feature_bins=speech_config.get("num_feature_bins", 160), gen=gen)
        super().__init__(blank=blank, speech_config=speech_config, encoder=enc, decoder=CtcDecoder(vocab_size, enc.dmodel, gen=gen), name=name)
        self.vocab_size = vocab_size
        self.time_reduction_factor = enc.time_reduction_factor


def deepspeech2_base_config(vocab_size=29):
    """examples/models/ctc/deepspeech2/base.yml.j2:5-44."""
    return dict(blank=0, vocab_size=vocab_size,
                speech_config=dict(sample_rate=16000, frame_ms=25, stride_ms=10, nfft=512, num_feature_bins=160, feature_type="spectrogram"),
                conv_type="conv2d", conv_kernels=[[11, 41], [11, 21]], conv_strides=[[2, 2], [1, 2]], conv_filters=[32, 32],
                conv_activation="relu", conv_padding="same", rnn_nlayers=5, rnn_type="lstm", rnn_units=512, rnn_bidirectional=True,
                rnn_rowconv=0, rnn_dropout=0.5, fc_nlayers=1, fc_units=1024, fc_activation="relu", fc_dropout=0.5)


# ------------------------------------------------------------------------------------------------ ContextNet
class ContextNetEncoder(_Params):
    """models/encoders/contextnet.py:277-341: blocks of [SeparableConv1D(causal) + BN + act] x n -> squeeze-excite -> + residual conv."""

    def __init__(self, blocks=(), alpha=1.0, num_feature_bins=80, name="contextnet_encoder", gen=None, **kwargs):
        super().__init__()
        self.blocks_cfg = [dict(b) for b in blocks]
        self.alpha = alpha
        cin = num_feature_bins
        self.time_reduction_factor = 1
        for bi, cfg in enumerate(self.blocks_cfg):
            filters = int(cfg["filters"] * alpha)
            k = cfg["kernel_size"]
            names = [f"conv_module_{i}" for i in range(cfg["nlayers"])] + ["se.conv_module"] + (["residual"] if cfg.get("residual", True) else [])
            c_in = cin
            for nm in names:
                src = cin if nm == "residual" else c_in
                pre = f"block_{bi}.{nm}"
                self.add(f"{pre}.conv.depthwise_kernel", (k, src), "glorot", (k * src, k), gen=gen)
                self.add(f"{pre}.conv.pointwise_kernel", (src, filters), "glorot", (src, filters), gen=gen)
                self.add(f"{pre}.conv.bias", (filters,))
                self.add(f"{pre}.bn.gamma", (filters,), "ones")
                self.add(f"{pre}.bn.beta", (filters,))
                self.add(f"{pre}.bn.moving_mean", (filters,), trainable=False)
                self.add(f"{pre}.bn.moving_variance", (filters,), "ones", trainable=False)
                if nm != "residual":
                    c_in = filters
            self.add(f"block_{bi}.se.fc1.kernel", (filters, filters // 8), "glorot", (filters, filters // 8), gen=gen)
            self.add(f"block_{bi}.se.fc1.bias", (filters // 8,))
            self.add(f"block_{bi}.se.fc2.kernel", (filters // 8, filters), "glorot", (filters // 8, filters), gen=gen)
            self.add(f"block_{bi}.se.fc2.bias", (filters,))
            cin = filters
            self.time_reduction_factor *= cfg.get("strides", 1)
        self.dmodel = cin

    def _conv_module(self, x, length, pre, k, stride, activation, training):
        """SeparableConv1D(causal) -> BN -> activation (contextnet.py:74-86).  LIBRARY (cuDNN / cuBLAS)"""
        dt = ops.compute_dtype()
        dwk, pwk = self[f"{pre}.conv.depthwise_kernel"], self[f"{pre}.conv.pointwise_kernel"]
        C = dwk.shape[1]
        y = F.pad(x.to(dt).transpose(1, 2), (k - 1, 0))
        y = F.conv1d(y, dwk.to(dt).t().unsqueeze(1), None, stride=stride, groups=C).transpose(1, 2)
        y = ops.linear(y, pwk, self[f"{pre}.conv.bias"])
        length = conv_output_length(length, k, "causal", stride)
        y = ops.batch_norm(y, self[f"{pre}.bn.gamma"], self[f"{pre}.bn.beta"], self[f"{pre}.bn.moving_mean"], self[f"{pre}.bn.moving_variance"],
                           training)
        if activation in ("silu", "swish"):
            y = F.silu(y)
        elif activation == "relu":
            y = F.relu(y)
        return y, length

    def forward(self, inputs, training=False):
        return self.call(inputs, training=training)

    # ---- fused route (fused_seq.py)
    def _fusable(self, nfeat, training):
        if not (training or not torch.is_grad_enabled()) or nfeat % 8:
            return False
        for cfg in self.blocks_cfg:
            f = int(cfg["filters"] * self.alpha)
            if f % 8 or f < 8 or cfg["kernel_size"] > 32 or cfg.get("activation", "silu") not in ("silu", "swish", "relu", "linear"):
                return False
        return True

    def _prepare_packs(self):
        if getattr(self, "_packs", None) is None:
            plan, packs = fused.PackPlan(), {}
            for n in self.names():
                if n.endswith(".conv.pointwise_kernel"):
                    pre = n[: -len(".conv.pointwise_kernel")]
                    packs[pre] = fused.WeightPack(self[n], self[f"{pre}.conv.bias"], plan=plan)
            self._packs, self._pack_plan = packs, plan
        self._pack_plan.run()

    def _conv_module_fused(self, x, length, pre, k, stride, act, training):
        bn = (self[f"{pre}.bn.gamma"], self[f"{pre}.bn.beta"], self[f"{pre}.bn.moving_mean"], self[f"{pre}.bn.moving_variance"])
        y = fused_seq.sepconv_bn(x, self[f"{pre}.conv.depthwise_kernel"], self._packs[pre], bn, stride, act, training)
        return y, conv_output_length(length, k, "causal", stride)

    def _call_fused(self, x, length, training):
        """Every ConvModule = depthwise kernel -> tcgen05 GEMM (pointwise + bias) -> BN + activation kernels; every block tail = masked-mean,
        squeeze-excite FC and scale + residual + activation kernels (fused_seq.py)."""
        self._prepare_packs()
        B, T, Fq, C = x.shape
        y = x.reshape(B, T, Fq * C).to(torch.bfloat16).contiguous()
        length = torch.as_tensor(length, device=x.device)
        for bi, cfg in enumerate(self.blocks_cfg):
            k, actn, stride = cfg["kernel_size"], cfg.get("activation", "silu"), cfg.get("strides", 1)
            act = kernels.act_code(actn)
            inp, inp_len = y, length
            for i in range(cfg["nlayers"]):
                last = i == cfg["nlayers"] - 1
                y, length = self._conv_module_fused(y, length, f"block_{bi}.conv_module_{i}", k, stride if last else 1, act, training)
            y, length = self._conv_module_fused(y, length, f"block_{bi}.se.conv_module", k, 1, act, training)
            res = None
            if cfg.get("residual", True):
                res, _ = self._conv_module_fused(inp, inp_len, f"block_{bi}.residual", k, stride, 0, training)
            y = fused_seq.se_block(y, res, length.to(torch.int32).contiguous(), self[f"block_{bi}.se.fc1.kernel"], self[f"block_{bi}.se.fc1.bias"],
                                   self[f"block_{bi}.se.fc2.kernel"], self[f"block_{bi}.se.fc2.bias"], act)
        return y, length

    def call(self, inputs, training=False):
        x, length = inputs  # [B,T,F,1]
        B, T, Fq, C = x.shape
        if ops.fused_seq_active(x) and self._fusable(Fq * C, training):
            return self._call_fused(x, length, training)
        y = x.reshape(B, T, Fq * C)  # Reshape (merge_two_last_dims)
        for bi, cfg in enumerate(self.blocks_cfg):
            k, act, stride = cfg["kernel_size"], cfg.get("activation", "silu"), cfg.get("strides", 1)
            inp, inp_len = y, length
            for i in range(cfg["nlayers"]):
                last = i == cfg["nlayers"] - 1
                y, length = self._conv_module(y, length, f"block_{bi}.conv_module_{i}", k, stride if last else 1, act, training)
            # SEModule (contextnet.py:152-164): conv module -> masked global average -> fc1 -> act -> fc2 -> sigmoid -> scale
            y, length = self._conv_module(y, length, f"block_{bi}.se.conv_module", k, 1, act, training)
            Tn = y.shape[1]
            mask = (torch.arange(Tn, device=y.device)[None, :] < torch.as_tensor(length, device=y.device)[:, None]).to(y.dtype)[..., None]
            se = (y * mask).sum(1, keepdim=True) / mask.sum(1, keepdim=True).clamp(min=1.0)
            se = ops.linear(se, self[f"block_{bi}.se.fc1.kernel"], self[f"block_{bi}.se.fc1.bias"], activation="swish" if act in ("silu", "swish") else act)
            se = torch.sigmoid(ops.linear(se, self[f"block_{bi}.se.fc2.kernel"], self[f"block_{bi}.se.fc2.bias"]).float()).to(y.dtype)
            y = y * se
            if cfg.get("residual", True):
                res, _ = self._conv_module(inp, inp_len, f"block_{bi}.residual", k, stride, "linear", training)
                y = y + res
            y = F.silu(y) if act in ("silu", "swish") else F.relu(y)
        return y, length

    def call_next(self, features, features_length, *args, **kwargs):
        y, length = self.call((features, features_length), training=False)
        return y, length, None


class ContextNet(Transducer):
    """models/transducer/contextnet.py:23-110."""

    def __init__(self, blank, vocab_size, speech_config, encoder_blocks, encoder_alpha=1.0, encoder_trainable=True,
                 prediction_label_encode_mode="embedding", prediction_embed_dim=512, prediction_num_rnns=1, prediction_rnn_units=320,
                 prediction_rnn_type="lstm", prediction_rnn_implementation=2, prediction_rnn_unroll=False, prediction_layer_norm=True,
                 prediction_projection_units=0, prediction_trainable=True, joint_dim=1024, joint_activation="tanh",
                 prejoint_encoder_linear=True, prejoint_prediction_linear=True, postjoint_linear=False, joint_mode="add",
                 joint_trainable=True, kernel_regularizer=None, bias_regularizer=None, name="contextnet", seed=None, **kwargs):
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        enc = ContextNetEncoder(blocks=encoder_blocks, alpha=encoder_alpha, num_feature_bins=speech_config.get("num_feature_bins", 80), gen=gen)
        super().__init__(blank=blank, vocab_size=vocab_size, speech_config=speech_config, encoder=enc,
                         prediction_label_encoder_mode=prediction_label_encode_mode, prediction_embed_dim=prediction_embed_dim,
                         prediction_num_rnns=prediction_num_rnns, prediction_rnn_units=prediction_rnn_units, prediction_rnn_type=prediction_rnn_type,
                         prediction_layer_norm=prediction_layer_norm, prediction_projection_units=prediction_projection_units, joint_dim=joint_dim,
                         joint_activation=joint_activation, joint_mode=joint_mode, prejoint_encoder_linear=prejoint_encoder_linear,
                         prejoint_prediction_linear=prejoint_prediction_linear, postjoint_linear=postjoint_linear, name=name, gen=gen)
        self.dmodel = enc.dmodel
        self.time_reduction_factor = enc.time_reduction_factor


def contextnet_config(vocab_size=1000, alpha=0.5):
    """examples/models/transducer/contextnet/small.yml.j2 (alpha 0.5 = small; alpha 1.0 = the paper's M): 23 blocks C0..C22."""
    def blk(n, f, s=1, res=True):
        return dict(nlayers=n, kernel_size=5, filters=f, strides=s, residual=res, activation="silu", padding="causal")
    blocks = [blk(1, 256, res=False)] + [blk(5, 256)] * 2 + [blk(5, 256, 2)] + [blk(5, 256)] * 3 + [blk(5, 256, 2)] + [blk(5, 256)] * 3 \
        + [blk(5, 512)] * 3 + [blk(5, 512, 2)] + [blk(5, 512)] * 7 + [blk(1, 640, res=False)]
    return dict(blank=0, vocab_size=vocab_size,
                speech_config=dict(sample_rate=16000, frame_ms=25, stride_ms=10, num_feature_bins=80, feature_type="log_mel_spectrogram"),
                encoder_alpha=alpha, encoder_blocks=blocks, prediction_embed_dim=640, prediction_num_rnns=1, prediction_rnn_units=640,
                prediction_layer_norm=True, joint_dim=640, joint_activation="tanh", joint_mode="add")
