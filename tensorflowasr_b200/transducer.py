"""Transducer / Conformer — host-side mirror of tensorflow_asr/models/transducer/{base_transducer,conformer}.py on B200.

Same class names, constructor kwargs, `call(TrainInput) -> TrainOutput`, `call_next`, `recognize` contracts as the reference
(base_transducer.py:305-712, conformer.py:22-141); parameters keep the Keras names (`prediction.lstm_0.kernel`,
`joint.vocab.kernel`, ...).  Training uses the fused joint + RNN-T loss head (`ops.transducer_loss_head`), which never hands
the [B,T,U+1,V] logits to Python; `call` still returns them for drop-in use.
"""
from __future__ import annotations

import torch

from . import fused, kernels, ops, schemas
from .base_model import BaseModel
from .encoder import ConformerEncoder, _Params
from .losses import RnntLoss


class TransducerPrediction(_Params):
    """base_transducer.py:38-173 (embedding label encoder, LSTM stack, optional LayerNorm / projection)."""

    def __init__(self, blank, vocab_size, label_encoder_mode="embedding", embed_dim=0, num_rnns=1, rnn_units=512, rnn_type="lstm",
                 rnn_implementation=2, rnn_unroll=False, layer_norm=True, projection_units=0, name="transducer_prediction",
                 gen=None, **kwargs):
        super().__init__()
        if label_encoder_mode != "embedding" or rnn_type != "lstm":
            raise NotImplementedError("B200 hot path covers the embedding + LSTM prediction network (SURVEY.md §8a A14)")
        self.blank, self.vocab_size = blank, vocab_size
        self.num_rnns, self.rnn_units, self.use_ln, self.projection_units = num_rnns, rnn_units, layer_norm, projection_units
        self.add("embedding.embeddings", (vocab_size, embed_dim), "uniform", gen=gen)
        din = embed_dim
        for i in range(num_rnns):
            self.add(f"lstm_{i}.kernel", (din, 4 * rnn_units), "glorot", (din, 4 * rnn_units), gen=gen)
            self.add(f"lstm_{i}.recurrent_kernel", (rnn_units, 4 * rnn_units), "glorot", (rnn_units, 4 * rnn_units), gen=gen)
            self.add(f"lstm_{i}.bias", (4 * rnn_units,))
            with torch.no_grad():
                self[f"lstm_{i}.bias"][rnn_units : 2 * rnn_units] = 1.0  # unit_forget_bias
            if layer_norm:
                self.add(f"ln_{i}.gamma", (rnn_units,), "ones")
                self.add(f"ln_{i}.beta", (rnn_units,))
            din = rnn_units
            if projection_units > 0:
                self.add(f"projection_{i}.kernel", (rnn_units, projection_units), "glorot", (rnn_units, projection_units), gen=gen)
                self.add(f"projection_{i}.bias", (projection_units,))
                din = projection_units
        self.output_dim = din

    def get_initial_state(self, batch_size: int):
        """-> zeros [B, num_rnns, nstates=2, rnn_units] (base_transducer.py:108-121)."""
        dev = self["embedding.embeddings"].device
        return torch.zeros(batch_size, self.num_rnns, 2, self.rnn_units, device=dev)

    def forward(self, inputs, training=False):
        tokens, length = inputs
        y = ops.embedding(tokens, self["embedding.embeddings"])
        for i in range(self.num_rnns):
            if ops.fused_active(y) and kernels.lstm_supported(self.rnn_units):
                # input projection on the tcgen05 GEMM, time loop on the cluster kernel (csrc/lstm.cu)
                y = fused.lstm(y, length, self[f"lstm_{i}.kernel"], self[f"lstm_{i}.recurrent_kernel"], self[f"lstm_{i}.bias"])
            else:
                y, _ = ops.lstm(y, length, self[f"lstm_{i}.kernel"], self[f"lstm_{i}.recurrent_kernel"], self[f"lstm_{i}.bias"])
            if self.use_ln:
                y = ops.layer_norm(y, self[f"ln_{i}.gamma"], self[f"ln_{i}.beta"])
            if self.projection_units > 0:
                y = ops.linear(y, self[f"projection_{i}.kernel"], self[f"projection_{i}.bias"])
        return y, length

    def call_next(self, inputs, previous_decoder_states):
        """inputs [B,1] tokens; states [B, num_rnns, 2, units] -> ([B,1,units], new states) (base_transducer.py:134-161)."""
        y = ops.embedding(inputs, self["embedding.embeddings"])
        new_states = []
        for i in range(self.num_rnns):
            st = (previous_decoder_states[:, i, 0], previous_decoder_states[:, i, 1])
            y, (h, c) = ops.lstm(y, None, self[f"lstm_{i}.kernel"], self[f"lstm_{i}.recurrent_kernel"], self[f"lstm_{i}.bias"], state=st)
            new_states.append(torch.stack([h, c], dim=1))
            if self.use_ln:
                y = ops.layer_norm(y, self[f"ln_{i}.gamma"], self[f"ln_{i}.beta"])
            if self.projection_units > 0:
                y = ops.linear(y, self[f"projection_{i}.kernel"], self[f"projection_{i}.bias"])
        return y, torch.stack(new_states, dim=1)


class TransducerJoint(_Params):
    """base_transducer.py:215-302 (pre-joint Dense on both branches, add, tanh, vocab Dense)."""

    def __init__(self, vocab_size, joint_dim=1024, activation="tanh", prejoint_encoder_linear=True, prejoint_prediction_linear=True,
                 postjoint_linear=False, joint_mode="add", encoder_dim=144, prediction_dim=320, name="tranducer_joint", gen=None, **kwargs):
        super().__init__()
        if activation.lower() != "tanh" or joint_mode != "add" or postjoint_linear or not prejoint_encoder_linear or not prejoint_prediction_linear:
            raise NotImplementedError("B200 hot path covers the shipped joint: Dense+Dense -> add -> tanh -> Dense (SURVEY.md §8a A15)")
        self.vocab_size, self.joint_dim = vocab_size, joint_dim
        self.add("enc.kernel", (encoder_dim, joint_dim), "glorot", (encoder_dim, joint_dim), gen=gen)
        self.add("enc.bias", (joint_dim,))
        self.add("pred.kernel", (prediction_dim, joint_dim), "glorot", (prediction_dim, joint_dim), gen=gen)
        self.add("pred.bias", (joint_dim,))
        self.add("vocab.kernel", (joint_dim, vocab_size), "glorot", (joint_dim, vocab_size), gen=gen)
        self.add("vocab.bias", (vocab_size,))

    def project(self, enc_out, pred_out):
        e = ops.linear(enc_out, self["enc.kernel"], self["enc.bias"])
        p = ops.linear(pred_out, self["pred.kernel"], self["pred.bias"])
        return e, p

    def forward(self, inputs, training=False):
        e, p = self.project(*inputs)
        return ops.joint_logits(e, p, self["vocab.kernel"], self["vocab.bias"])


class Transducer(BaseModel):
    """base_transducer.py:305-712."""

    def __init__(self, blank, vocab_size, speech_config, encoder, prediction_label_encoder_mode="embedding", prediction_embed_dim=512,
                 prediction_num_rnns=1, prediction_rnn_units=320, prediction_rnn_type="lstm", prediction_rnn_implementation=2,
                 prediction_rnn_unroll=False, prediction_layer_norm=True, prediction_projection_units=0, prediction_trainable=True,
                 joint_dim=1024, joint_activation="tanh", joint_mode="add", joint_trainable=True, prejoint_encoder_linear=True,
                 prejoint_prediction_linear=True, postjoint_linear=False, kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, recurrent_regularizer=None, name="transducer", gen=None, **kwargs):
        super().__init__(speech_config=speech_config, name=name)
        self.blank = blank
        self.vocab_size = vocab_size
        self.encoder = encoder
        self.predict_net = TransducerPrediction(
            blank=blank, vocab_size=vocab_size, label_encoder_mode=prediction_label_encoder_mode, embed_dim=prediction_embed_dim,
            num_rnns=prediction_num_rnns, rnn_units=prediction_rnn_units, rnn_type=prediction_rnn_type, layer_norm=prediction_layer_norm,
            projection_units=prediction_projection_units, gen=gen)
        self.joint_net = TransducerJoint(
            vocab_size=vocab_size, joint_dim=joint_dim, activation=joint_activation, prejoint_encoder_linear=prejoint_encoder_linear,
            prejoint_prediction_linear=prejoint_prediction_linear, postjoint_linear=postjoint_linear, joint_mode=joint_mode,
            encoder_dim=encoder.dmodel, prediction_dim=self.predict_net.output_dim, gen=gen)
        self.time_reduction_factor = 1

    def compile(self, optimizer, output_shapes=None, **kwargs):
        loss = RnntLoss(blank=self.blank, output_shapes=output_shapes, name="rnnt_loss")
        return super().compile(loss=loss, optimizer=optimizer, **kwargs)

    # ------------------------------------------------------------------ reference `call`
    def _branches(self, inputs: schemas.TrainInput, training):
        # The prediction network (embedding -> LSTM -> LN: ~65 dependent label steps) does not depend on the encoder until the joint: on the
        # GPU it runs on a second stream — a parallel branch of the captured step graph — and autograd runs its backward there as well.
        side = fused.SIDE.branch_stream(inputs.inputs)
        if side is not None:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                pred, *_ = self.predict_net((inputs.predictions, inputs.predictions_length), training=training)
        features, features_length = self.feature_extraction((inputs.inputs, inputs.inputs_length), training=training)
        enc, logits_length, *_ = self.encoder((features, features_length), training=training)
        if side is not None:
            main.wait_stream(side)
            pred.record_stream(main)  # allocated on the side stream, consumed by the joint on the main stream
        else:
            pred, *_ = self.predict_net((inputs.predictions, inputs.predictions_length), training=training)
        return enc, logits_length, pred

    def call(self, inputs: schemas.TrainInput, training=False):
        enc, logits_length, pred = self._branches(inputs, training)
        logits = self.joint_net((enc, pred), training=training)
        return schemas.TrainOutput(logits=logits, logits_length=torch.as_tensor(logits_length, device=logits.device).to(torch.int32))

    # ------------------------------------------------------------------ fused training head
    def compute_losses(self, data: schemas.TrainData, training=True):
        """Per-utterance RNN-T losses [B] through the fused joint + loss head: same value as
        `RnntLoss.call(y, self.call(x))` (base_model.py:137-147 with rnnt_loss.py:50-61) without materialising logits for Python."""
        x, y = data
        enc, logits_length, pred = self._branches(x, training)
        e, p = self.joint_net.project(enc, pred)
        dev = e.device
        logit_length = torch.as_tensor(logits_length, device=dev).to(torch.int32)
        label_length = y.labels_length.to(device=dev, dtype=torch.int32)
        labels = y.labels.to(device=dev, dtype=torch.int32).contiguous()
        logit_length = torch.where(logit_length < label_length, label_length, logit_length).contiguous()  # base_loss.py:36
        if ops.fused_active(e):
            pack = fused.WeightPack(self.joint_net["vocab.kernel"], self.joint_net["vocab.bias"])
            return fused.transducer_head(e.to(torch.bfloat16), p.to(torch.bfloat16), pack, labels, logit_length, label_length.contiguous())
        return ops.transducer_loss_head(e, p, self.joint_net["vocab.kernel"], self.joint_net["vocab.bias"], labels, logit_length,
                                        label_length.contiguous())

    # ------------------------------------------------------------------ decoding (greedy)
    def call_next(self, current_frames, previous_tokens, previous_decoder_states):
        """-> (log-probs [B,1,1,V], new decoder states) — base_transducer.py:437-464."""
        y, new_states = self.predict_net.call_next(previous_tokens, previous_decoder_states)
        ytu = self.joint_net((current_frames, y), training=False)
        return torch.log_softmax(ytu.float(), dim=-1), new_states

    def get_initial_tokens(self, batch_size=1):
        return torch.full((batch_size, 1), self.blank, dtype=torch.int32, device=self.joint_net["vocab.bias"].device)

    def get_initial_encoder_states(self, batch_size=1):
        return []

    def get_initial_decoder_states(self, batch_size=1):
        return self.predict_net.get_initial_state(batch_size)

    @torch.no_grad()
    def recognize(self, inputs: schemas.PredictInput, max_tokens_per_frame: int = 3, **kwargs):
        """Greedy decode dispatcher (base_transducer.py:474-494): batch 1 -> recognize_single, otherwise recognize_batch."""
        if inputs.inputs.shape[0] == 1:
            return self.recognize_single(inputs, max_tokens_per_frame=max_tokens_per_frame, **kwargs)
        return self.recognize_batch(inputs, **kwargs)

    def _encode_for_decode(self, inputs):
        features, features_length = self.feature_extraction((inputs.inputs, inputs.inputs_length), training=False)
        enc, enc_len, next_enc_states = self.encoder.call_next(features, features_length, inputs.previous_encoder_states)
        B = enc.shape[0]
        dev = enc.device
        enc_len = torch.as_tensor(enc_len, device=dev).to(torch.int64)
        prev_tok = (inputs.previous_tokens if inputs.previous_tokens is not None else self.get_initial_tokens(B)).to(dev).to(torch.int32)
        states = (inputs.previous_decoder_states if inputs.previous_decoder_states is not None else self.get_initial_decoder_states(B)).to(dev)
        return enc, enc_len, next_enc_states, prev_tok, states

    # ------------------------------------------------------------------ persistent decode kernel (csrc/decode.cu)
    def _persistent_ok(self):
        pn = self.predict_net
        return ops.persistent_decode_active() and pn.num_rnns == 1 and pn.projection_units == 0

    def _decode_weights(self):
        pn, jn = self.predict_net, self.joint_net

        def f(t):
            return t.detach().float().contiguous()

        return dict(emb=f(pn["embedding.embeddings"]), w_ih=f(pn["lstm_0.kernel"]), w_hh=f(pn["lstm_0.recurrent_kernel"]), b=f(pn["lstm_0.bias"]),
                    ln_g=f(pn["ln_0.gamma"]) if pn.use_ln else None, ln_b=f(pn["ln_0.beta"]) if pn.use_ln else None,
                    wp=f(jn["pred.kernel"]), bp=f(jn["pred.bias"]), wv=f(jn["vocab.kernel"]), bv=f(jn["vocab.bias"]))

    def _encoder_projection(self, enc):
        """joint.enc Dense for all frames at once (the only encoder-side term of call_next): [B,T,J] fp32."""
        return ops.linear(enc, self.joint_net["enc.kernel"], self.joint_net["enc.bias"]).float().contiguous()

    def _recognize_batch_persistent(self, enc, enc_len, next_enc_states, prev_tok, states):
        """recognize_batch as two launches of the persistent kernel.  The reference loop runs until EVERY utterance has reached its last
        frame, and utterances that got there earlier keep evolving meanwhile: pass 1 (mode 0) runs each utterance to its own frame
        condition and reports its iteration count, the loop length is L = max_b iters[b], pass 2 (mode 1) runs the remaining
        L - iters[b] iterations.  Returns None when the token-budget stop of the reference loop could have fired first (all utterances
        at 2T' tokens): the caller then takes the host loop, which implements that case verbatim."""
        B, T, _ = enc.shape
        dev = enc.device
        max_tokens = 2 * T + 1
        w = self._decode_weights()
        e_proj = self._encoder_projection(enc)
        nframes = enc_len.to(torch.int32).contiguous()
        pt = prev_tok[:, 0].to(torch.int32).contiguous().clone()
        st = states[:, 0].float().contiguous().clone()
        fi = torch.zeros(B, dtype=torch.int32, device=dev)
        ti = torch.ones(B, dtype=torch.int32, device=dev)
        toks = torch.full((B, max_tokens), self.blank, dtype=torch.int32, device=dev)
        iters = torch.zeros(B, dtype=torch.int32, device=dev)
        cap = torch.zeros(B, dtype=torch.int32, device=dev)
        kernels.greedy_decode(w, e_proj, nframes, pt, st, fi, ti, toks, iters, cap, self.blank, mode=0, max_iters=4 * max_tokens + 8)
        host = torch.stack([iters, cap]).cpu()  # the one host round trip of the decode
        i_b, j_b = host[0], host[1]
        if bool((i_b < 0).any()):
            return None
        L = int(i_b.max())
        if bool((j_b >= 0).all()) and int(j_b.max()) < L:
            return None
        budget = (L - iters).contiguous()
        kernels.greedy_decode(w, e_proj, nframes, pt, st, fi, ti, toks, budget, cap, self.blank, mode=1)
        if bool((ti >= max_tokens - 1).all()):
            return None
        return schemas.PredictOutput(tokens=toks, next_tokens=pt[:, None], next_encoder_states=next_enc_states, next_decoder_states=st[:, None])

    def _recognize_single_persistent(self, enc, nframes, next_enc_states, prev_tok, states, max_tokens_per_frame):
        dev = enc.device
        n = max(nframes, 0) * max_tokens_per_frame
        w = self._decode_weights()
        e_proj = self._encoder_projection(enc)
        pt = prev_tok[:, 0].to(torch.int32).contiguous().clone()
        st = states[:, 0].float().contiguous().clone()
        toks = torch.full((1, max(n, 1)), self.blank, dtype=torch.int32, device=dev)
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        kernels.greedy_decode(w, e_proj, torch.full((1,), nframes, dtype=torch.int32, device=dev), pt, st, z.clone(), torch.full_like(z, -1), toks, z.clone(),
                              z.clone(), self.blank, mode=2, per_frame_cap=max_tokens_per_frame)
        return schemas.PredictOutput(tokens=toks[:, :n], next_tokens=pt[:, None], next_encoder_states=next_enc_states, next_decoder_states=st[:, None])

    @torch.no_grad()
    def recognize_single(self, inputs: schemas.PredictInput, max_tokens_per_frame: int = 3, **kwargs):
        """base_transducer.py:577-712 (batch 1): walk the frames; a frame is left on blank or after `max_tokens_per_frame`
        non-blank symbols; tokens [1, nframes * max_tokens_per_frame], blank-padded."""
        enc, enc_len, next_enc_states, prev_tok, states = self._encode_for_decode(inputs)
        nframes = int(enc_len[0])
        if self._persistent_ok():
            return self._recognize_single_persistent(enc, nframes, next_enc_states, prev_tok, states, max_tokens_per_frame)
        toks = torch.full((max(nframes, 0) * max_tokens_per_frame,), self.blank, dtype=torch.int32, device=enc.device)
        frame, token_index, per_frame = 0, -1, 0
        while frame < nframes:
            logp, new_states = self.call_next(enc[:, frame : frame + 1], prev_tok, states)
            cur = int(logp[0, 0, 0].argmax())
            if cur == self.blank:
                frame, per_frame = frame + 1, 0
            else:
                per_frame += 1
                token_index += 1
                toks[token_index] = cur
                prev_tok = torch.full_like(prev_tok, cur)
                states = new_states
                if per_frame >= max_tokens_per_frame:
                    frame, per_frame = frame + 1, 0
        return schemas.PredictOutput(tokens=toks[None, :], next_tokens=prev_tok, next_encoder_states=next_enc_states, next_decoder_states=states)

    @torch.no_grad()
    def recognize_batch(self, inputs: schemas.PredictInput, **kwargs):
        """base_transducer.py:496-575, loop semantics kept verbatim (including its quirks): the loop runs until EVERY utterance has
        reached its last frame index (nframes - 1) or the token budget; finished utterances keep reading their clamped last frame until
        frame_index > nframes; emitted symbols are stored from column 2 on (tokens_indices starts at 1), blanks overwrite column 0;
        tokens [B, 2 * T' + 1]."""
        enc, enc_len, next_enc_states, prev_tok, states = self._encode_for_decode(inputs)
        if self._persistent_ok():
            out = self._recognize_batch_persistent(enc, enc_len, next_enc_states, prev_tok, states)
            if out is not None:
                return out
        B, T, _ = enc.shape
        dev = enc.device
        nframes = enc_len[:, None]  # [B,1]
        frame_idx = torch.zeros(B, 1, dtype=torch.int64, device=dev)
        max_tokens = T * 2 + 1
        tokens = torch.full((B, max_tokens), self.blank, dtype=torch.int32, device=dev)
        tok_idx = torch.ones(B, 1, dtype=torch.int64, device=dev)
        rows = torch.arange(B, device=dev)
        while True:
            stop = bool((frame_idx >= nframes - 1).all()) or bool((tok_idx >= max_tokens - 1).all())
            if stop:
                break
            cur_frames = enc[rows, torch.minimum(frame_idx, nframes - 1)[:, 0]].unsqueeze(1)
            logp, new_states = self.call_next(cur_frames, prev_tok, states)
            cur = logp[:, 0, 0].argmax(-1).to(torch.int32)[:, None]  # [B,1]
            eq_blank = (cur == self.blank) | (tok_idx >= max_tokens) | (frame_idx > nframes)
            upd_tok = torch.where(eq_blank, torch.full_like(cur, self.blank), cur)[:, 0]
            upd_idx = torch.where(eq_blank, torch.zeros_like(tok_idx), torch.clamp(tok_idx + 1, max=max_tokens - 1))[:, 0]
            tokens[rows, upd_idx] = upd_tok
            tok_idx = torch.where(eq_blank, tok_idx, torch.clamp(tok_idx + 1, max=max_tokens - 1))
            frame_idx = torch.where(eq_blank, frame_idx + 1, frame_idx)
            prev_tok = torch.where(eq_blank, prev_tok, cur)
            states = torch.where(eq_blank[:, :, None, None], states, new_states)
        return schemas.PredictOutput(tokens=tokens, next_tokens=prev_tok, next_encoder_states=next_enc_states, next_decoder_states=states)

    def recognize_beam(self, inputs: schemas.PredictInput, beam_width: int = 10, **kwargs):
        """The reference's transducer beam search is a stub that returns greedy (base_transducer.py:841-842); mirrored."""
        return self.recognize(inputs=inputs, **kwargs)


class Conformer(Transducer):
    """models/transducer/conformer.py:22-141."""

    def __init__(self, blank, vocab_size, speech_config, encoder_subsampling, encoder_dmodel=144, encoder_num_blocks=16,
                 encoder_head_size=36, encoder_num_heads=4, encoder_mha_type="relmha", encoder_interleave_relpe=True,
                 encoder_use_attention_causal_mask=False, encoder_use_attention_auto_mask=True, encoder_kernel_size=31,
                 encoder_padding="causal", encoder_ffm_scale_factor=4, encoder_ffm_residual_factor=0.5,
                 encoder_mhsam_residual_factor=1.0, encoder_mhsam_use_attention_bias=False, encoder_mhsam_causal=False,
                 encoder_mhsam_flash_attention=False, encoder_convm_scale_factor=2, encoder_convm_residual_factor=1.0,
                 encoder_convm_use_group_conv=False, encoder_convm_dw_norm_type="batch", encoder_dropout=0.1,
                 encoder_module_norm_position="pre", encoder_block_norm_position="post", encoder_memory_length=None,
                 encoder_history_size=None, encoder_chunk_size=None, encoder_trainable=True, prediction_label_encode_mode="embedding",
                 prediction_embed_dim=512, prediction_num_rnns=1, prediction_rnn_units=320, prediction_rnn_type="lstm",
                 prediction_rnn_implementation=2, prediction_rnn_unroll=False, prediction_layer_norm=True,
                 prediction_projection_units=0, prediction_trainable=True, joint_dim=1024, joint_activation="tanh",
                 prejoint_encoder_linear=True, prejoint_prediction_linear=True, postjoint_linear=False, joint_mode="add",
                 joint_trainable=True, kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 recurrent_regularizer=None, name="conformer", seed=None, **kwargs):
        gen = None
        if seed is not None:
            gen = torch.Generator().manual_seed(int(seed))
        encoder = ConformerEncoder(
            subsampling=encoder_subsampling, dmodel=encoder_dmodel, num_blocks=encoder_num_blocks, head_size=encoder_head_size,
            num_heads=encoder_num_heads, mha_type=encoder_mha_type, interleave_relpe=encoder_interleave_relpe,
            use_attention_causal_mask=encoder_use_attention_causal_mask, use_attention_auto_mask=encoder_use_attention_auto_mask,
            kernel_size=encoder_kernel_size, padding=encoder_padding, ffm_scale_factor=encoder_ffm_scale_factor,
            ffm_residual_factor=encoder_ffm_residual_factor, mhsam_residual_factor=encoder_mhsam_residual_factor,
            mhsam_use_attention_bias=encoder_mhsam_use_attention_bias, mhsam_causal=encoder_mhsam_causal,
            mhsam_flash_attention=encoder_mhsam_flash_attention, convm_scale_factor=encoder_convm_scale_factor,
            convm_residual_factor=encoder_convm_residual_factor, convm_use_group_conv=encoder_convm_use_group_conv,
            convm_dw_norm_type=encoder_convm_dw_norm_type, dropout=encoder_dropout, module_norm_position=encoder_module_norm_position,
            block_norm_position=encoder_block_norm_position, memory_length=encoder_memory_length, history_size=encoder_history_size,
            chunk_size=encoder_chunk_size, name="encoder", num_feature_bins=speech_config.get("num_feature_bins", 80), gen=gen)
        super().__init__(
            blank=blank, vocab_size=vocab_size, speech_config=speech_config, encoder=encoder,
            prediction_label_encoder_mode=prediction_label_encode_mode, prediction_embed_dim=prediction_embed_dim,
            prediction_num_rnns=prediction_num_rnns, prediction_rnn_units=prediction_rnn_units, prediction_rnn_type=prediction_rnn_type,
            prediction_layer_norm=prediction_layer_norm, prediction_projection_units=prediction_projection_units, joint_dim=joint_dim,
            joint_activation=joint_activation, joint_mode=joint_mode, prejoint_encoder_linear=prejoint_encoder_linear,
            prejoint_prediction_linear=prejoint_prediction_linear, postjoint_linear=postjoint_linear, name=name, gen=gen)
        self.dmodel = encoder_dmodel
        self.time_reduction_factor = self.encoder.time_reduction_factor


def conformer_small_config(vocab_size=1000):
    """examples/models/transducer/conformer/small.yml.j2:5-65, SpecAugment included."""
    return dict(
        blank=0, vocab_size=vocab_size,
        speech_config=dict(sample_rate=16000, frame_ms=25, stride_ms=10, nfft=512, num_feature_bins=80, feature_type="log_mel_spectrogram",
                           augmentation_config=dict(feature_augment=dict(  # small.yml.j2:12-24 (active in training mode only)
                               time_masking=dict(prob=1.0, num_masks=10, mask_factor=-1, p_upperbound=0.05, mask_value=0),
                               freq_masking=dict(prob=1.0, num_masks=1, mask_factor=27, mask_value=0)))),
        encoder_subsampling=dict(class_name="tensorflow_asr.models.layers.subsampling>Conv2dSubsampling",
                                 config=dict(filters=[144, 144], kernels=[3, 3], strides=[2, 2], paddings=["causal", "causal"],
                                             norms=["batch", "batch"], activations=["swish", "swish"])),
        encoder_ffm_residual_factor=0.5, encoder_mhsam_residual_factor=1.0, encoder_convm_residual_factor=1.0, encoder_dmodel=144,
        encoder_num_blocks=16, encoder_head_size=36, encoder_num_heads=4, encoder_mha_type="relmha", encoder_interleave_relpe=True,
        encoder_use_attention_causal_mask=False, encoder_use_attention_auto_mask=True, encoder_mhsam_use_attention_bias=False,
        encoder_kernel_size=31, encoder_dropout=0.1, encoder_padding="causal", encoder_memory_length=None,
        prediction_label_encode_mode="embedding", prediction_embed_dim=320, prediction_num_rnns=1, prediction_rnn_units=320,
        prediction_rnn_type="lstm", prediction_rnn_implementation=2, prediction_rnn_unroll=False, prediction_layer_norm=True,
        prediction_projection_units=0, joint_dim=320, prejoint_encoder_linear=True, prejoint_prediction_linear=True,
        postjoint_linear=False, joint_activation="tanh", joint_mode="add",
    )
