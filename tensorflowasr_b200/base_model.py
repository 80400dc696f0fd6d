"""BaseModel — host-side mirror of tensorflow_asr/models/base_model.py (the trainer / step functions) on B200.

Keeps the reference's surface (`make`, `compile`, `train_step`, `_train_step`, `_apply_gradients`, `train_step_ga`,
`test_step`, `predict_step`, `recognize`, `tokenizer`, base_model.py:33-331).  The distribution strategy
(`tf.distribute.MirroredStrategy`, utils/env_util.py:57-70) becomes one process per GPU with a bucketed NCCL all-reduce of
the gradients (`tensorflowasr_b200.dist`): the loss is scaled by 1/num_replicas and the gradients are SUMMED, exactly the
arrangement base_model.py:166-173 describes.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib as _lib_mod
from . import dist as tdist
from . import schemas
from .features import FeatureExtraction

L2_DEFAULT = 1e-6  # models/encoders/conformer.py:27 (kernel_regularizer = l2(1e-6))


class TransformerSchedule:
    """optimizers/schedules.py:19-45: lr = scale * dmodel^-0.5 * min(step^-0.5, step * warmup^-1.5), clipped to [min_lr, max_lr]."""

    def __init__(self, dmodel, scale=1.0, warmup_steps=4000, max_lr=None, min_lr=None):
        self.dmodel, self.scale, self.warmup_steps = float(dmodel), float(scale), float(warmup_steps)
        self.max_lr = eval(max_lr) if isinstance(max_lr, str) else max_lr
        self.min_lr = eval(min_lr) if isinstance(min_lr, str) else min_lr

    def __call__(self, current_step):
        step = max(float(current_step), 1e-9)
        lr = self.scale * (self.dmodel ** -0.5) * min(step ** -0.5, step * (self.warmup_steps ** -1.5))
        if self.max_lr is not None:
            lr = min(self.max_lr, lr)
        if self.min_lr is not None:
            lr = max(self.min_lr, lr)
        return lr


def make_optimizer(model: nn.Module, learning_rate=None, beta_1=0.9, beta_2=0.98, epsilon=1e-9, weight_decay=1e-6, flat=True):
    """Keras Adam as configured in small.yml.j2:73-87 (decoupled weight decay): on CUDA the single-launch FlatAdamW (optim.py), on CPU
    (tests of the host logic) torch's AdamW."""
    lr0 = learning_rate(1) if callable(learning_rate) else (learning_rate or 1e-3)
    params = [p for p in model.parameters() if p.requires_grad]
    on_gpu = params[0].is_cuda or _lib_mod.EMULATED
    # lr lives in a device tensor so that a captured CUDA graph picks up the schedule's new value at every replay
    lr = torch.tensor(float(lr0), device=params[0].device) if on_gpu else lr0
    if on_gpu and flat:
        from .optim import FlatAdamW

        opt = FlatAdamW(params, lr=lr, betas=(beta_1, beta_2), eps=epsilon, weight_decay=weight_decay or 0.0)  # one launch per step
    else:
        opt = torch.optim.AdamW(params, lr=lr, betas=(beta_1, beta_2), eps=epsilon, weight_decay=weight_decay or 0.0, fused=on_gpu,
                                capturable=on_gpu)
    opt._tfasr_schedule = learning_rate if callable(learning_rate) else None
    return opt


class BaseModel(nn.Module):
    def __init__(self, speech_config: dict, name="base_model", **kwargs):
        super().__init__()
        self.name = name
        self.feature_extraction = FeatureExtraction(**speech_config)
        self._tokenizer = None
        self.optimizer = None
        self.tfasr_loss = None
        self.use_ga = False
        self.ga_steps = 1
        self.gwn_config = None
        self.gradn_config = None
        self.l2 = L2_DEFAULT
        self._iterations = 0
        self._loss_sum = 0.0
        self._loss_count = 0
        self._reducer = None
        self._flat_opt = False

    # ------------------------------------------------------------------ misc surface
    @property
    def tokenizer(self):
        return self._tokenizer

    @tokenizer.setter
    def tokenizer(self, tokenizer):
        self._tokenizer = tokenizer

    @property
    def trainable_weights(self):
        return [p for p in self.parameters() if p.requires_grad]

    def make(self, input_shape=[None], prediction_shape=[None], batch_size=None, **kwargs) -> schemas.TrainOutput:
        """Builds nothing (parameters exist from construction) but returns the per-replica output shapes like the reference
        (base_model.py:68-100).  Dynamic dims stay None."""
        assert batch_size is not None and batch_size > 0
        self._batch_size = batch_size
        self._per_replica_batch_size = int(batch_size / tdist.world_size())
        n = input_shape[0]
        u = prediction_shape[0]
        t = None
        if n is not None:
            t = int(self.feature_extraction.get_nframes(n + self.feature_extraction.padding))
            t = self._reduced_length(t)
        return self._output_shapes(self._per_replica_batch_size, t, u)

    def _reduced_length(self, t):
        return t

    def _output_shapes(self, b, t, u):
        return schemas.TrainOutput(logits=(b, t, u, getattr(self, "vocab_size", None)), logits_length=(b,))

    def compile(self, loss, optimizer=None, run_eagerly=False, ga_steps=None, gwn_config=None, gradn_config=None, **kwargs):
        if optimizer is None or isinstance(optimizer, dict):
            cfg = dict(optimizer or {})
            optimizer = make_optimizer(self, **cfg.get("config", cfg))
        self.optimizer = optimizer
        if isinstance(ga_steps, int) and ga_steps > 1:
            self.use_ga, self.ga_steps = True, ga_steps
        else:
            self.use_ga, self.ga_steps = False, 1
        self.gwn_config = gwn_config
        self.gradn_config = gradn_config
        self.tfasr_loss = loss
        # every parameter's .grad becomes a view into ONE flat fp32 buffer: the fused kernels accumulate weight gradients straight
        # into it, zeroing is one memset, and the data-parallel all-reduce needs no packing
        self._reducer = tdist.GradientReducer(self.trainable_weights)
        self._flat_opt = False
        from .optim import FlatAdamW

        if isinstance(self.optimizer, FlatAdamW):
            self.optimizer.attach_grads(self._reducer.flat, self._reducer.params)
            mask = torch.zeros(self._reducer.flat.numel(), dtype=torch.uint8, device=self._reducer.flat.device)
            by_id = {id(p): n for n, p in self.named_parameters()}
            for p, off in zip(self._reducer.params, self._reducer.offsets):
                if _is_regularized(by_id[id(p)]):
                    mask[off:off + p.numel()] = 1
            self.optimizer.set_l2(mask, self.l2)
            self._flat_opt = True
        self._graph = None
        self._use_graph = False
        return self

    def call(self, inputs: schemas.TrainInput, training=False):
        raise NotImplementedError()

    def forward(self, inputs: schemas.TrainInput, training=False):
        return self.call(inputs, training=training)

    # ------------------------------------------------------------------ step functions
    def compute_losses(self, data: schemas.TrainData, training=True):
        """Per-utterance losses [B].  Default = the reference's route: y_pred = self(x); tfasr_loss.call(y, y_pred)."""
        x, y = data
        y_pred = self(x, training=training)
        return self.tfasr_loss.call(y, y_pred)

    def regularization_loss(self):
        """Sum of the Keras l2(1e-6) regularizers attached to kernels and norm scales/offsets (conformer.py:27 and the
        `*_regularizer=kernel_regularizer` arguments).  Applied through its gradient 2*l2*w in `_apply_gradients`."""
        return sum(self.l2 * (p.float() ** 2).sum() for n, p in self.named_parameters() if _is_regularized(n) and p.requires_grad)

    def tfasr_compute_loss(self, losses):
        """base_model.py:137-147 + Keras `sum_over_batch_size` reduction, scaled by 1/num_replicas for the gradient SUM."""
        return losses.mean() / tdist.world_size()

    def enable_cuda_graph(self, flag: bool = True):
        """Capture the whole training step (forward, backward, all-reduce, optimizer) in ONE CUDA graph and replay it per step:
        the B200 replacement for `tf.function(one_step_on_data)` (base_model.py:252-269).  Shapes must stay fixed."""
        self._use_graph = bool(flag)
        self._graph = None
        return self

    def _train_step(self, data: schemas.TrainData):
        from . import fused

        fused.STATE.bump()  # new dropout streams for this step
        losses = self.compute_losses(data, training=True)
        loss = self.tfasr_compute_loss(losses)
        loss.backward()
        fused.SIDE.join()  # weight-gradient GEMMs on the side stream (normally already joined by the end-of-backward callback)
        self._last_losses = losses.detach()
        return [p.grad for p in self.trainable_weights]

    def _apply_gradients(self, gradients=None, lr_from_host=True):
        params = self.trainable_weights
        if self._reducer is not None:
            self._reducer.all_reduce()  # C1: one bucketed NCCL all-reduce (SUM) over NVLink (no-op on one GPU)
        if self._flat_opt:
            # l2-regularizer gradient, AdamW and the zeroing of the flat gradient buffer: one kernel
            if lr_from_host:
                self._advance_schedule()
            self.optimizer.step(zero_grad=True)
            return
        if self.l2:
            reg = [p for n, p in self.named_parameters() if p.requires_grad and p.grad is not None and _is_regularized(n)]
            if reg:
                torch._foreach_add_([p.grad for p in reg], [p.detach() for p in reg], alpha=2.0 * self.l2)
        if lr_from_host:
            self._advance_schedule()
        self.optimizer.step()
        if self._reducer is not None:
            self._reducer.zero()
        else:
            self.optimizer.zero_grad(set_to_none=False)
        del params

    def _advance_schedule(self):
        self._iterations += 1
        sched = getattr(self.optimizer, "_tfasr_schedule", None)
        if sched is not None:
            lr = sched(self._iterations)
            for g in self.optimizer.param_groups:
                if torch.is_tensor(g["lr"]):
                    g["lr"].fill_(lr)
                else:
                    g["lr"] = lr

    def train_step(self, data: schemas.TrainData):
        if self._use_graph and data.inputs.inputs.is_cuda:
            return self._graph_train_step(data)
        self._train_step(data)
        self._apply_gradients()
        return {"loss": self._last_losses}

    # ------------------------------------------------------------------ CUDA-graph step
    def _graph_train_step(self, data: schemas.TrainData):
        flat = [data.inputs.inputs, data.inputs.inputs_length, data.inputs.predictions, data.inputs.predictions_length,
                data.labels.labels, data.labels.labels_length]
        key = tuple(None if t is None else (tuple(t.shape), t.dtype) for t in flat)  # CTC models carry no predictions
        if self._graph is None or self._graph["key"] != key:
            static = [None if t is None else torch.empty_like(t) for t in flat]
            for s, t in zip(static, flat):
                if t is not None:
                    s.copy_(t)
            sdata = schemas.TrainData(schemas.TrainInput(*static[:4]), schemas.TrainLabel(*static[4:]))
            # warm-up on a side stream (allocator, cuDNN / cuBLAS handles, lazily built tables), then capture.  The warm-up steps are
            # real steps on the first batch: everything they change (parameters, Adam moments and step count, BatchNorm moving
            # statistics and every other buffer, the schedule position, the dropout counter) is put back afterwards, so that a
            # graph-mode run is step-equivalent to the eager one
            snap = self._snapshot_train_state()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._advance_schedule()
                    self._train_step(sdata)
                    self._apply_gradients(lr_from_host=False)
            torch.cuda.current_stream().wait_stream(side)
            self._restore_train_state(snap)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._train_step(sdata)
                self._apply_gradients(lr_from_host=False)
                out = self._last_losses.clone()
            self._graph = {"key": key, "graph": graph, "static": static, "out": out}
        g = self._graph
        for s, t in zip(g["static"], flat):
            if t is not None and s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self._advance_schedule()
        g["graph"].replay()
        return {"loss": g["out"]}

    def _snapshot_train_state(self):
        from . import fused

        snap = {"iterations": self._iterations, "tensors": []}
        seen = set()

        def keep(t):
            if torch.is_tensor(t) and t.data_ptr() not in seen:
                seen.add(t.data_ptr())
                snap["tensors"].append((t, t.detach().clone()))

        opt = self.optimizer
        if self._flat_opt:
            for t in (opt.flat_p, opt.exp_avg, opt.exp_avg_sq, opt.step_t, self._reducer.flat):
                keep(t)
        else:
            for p in self.trainable_weights:
                keep(p.data)
                if p.grad is not None:
                    keep(p.grad)
            for st in opt.state.values():
                for v in st.values():
                    keep(v)
        for g in opt.param_groups:
            keep(g["lr"])
        for b in self.buffers():
            keep(b)
        if fused.STATE.seed is not None:
            keep(fused.STATE.seed)
        return snap

    def _restore_train_state(self, snap):
        with torch.no_grad():
            for t, saved in snap["tensors"]:
                t.copy_(saved)
        self._iterations = snap["iterations"]

    def train_step_ga(self, data: schemas.TrainData, do_apply=None):
        """Gradient accumulation (base_model.py:200-209): gradients accumulate in .grad across calls; apply on the last one."""
        if self._reducer is not None:
            self._reducer.defer = True  # no collective may start while later micro-batches still accumulate into the flat buffer
        try:
            self._train_step(data)
        finally:
            if self._reducer is not None:
                self._reducer.defer = False
        if do_apply:
            if self.ga_steps > 1:
                torch._foreach_div_([p.grad for p in self.trainable_weights if p.grad is not None], float(self.ga_steps))
            self._apply_gradients()  # all_reduce() launches every bucket now (none was launched by the hooks) and waits
        return {"loss": self._last_losses}

    # ------------------------------------------------------------------ training state beyond the weights
    def set_training_seed(self, seed: int):
        """Seed of the dropout / SpecAugment streams, mixed with this replica's rank (the reference seeds python / numpy / TF globally,
        env_util.py:147-165, and lets every replica draw its own masks)."""
        from . import fused

        fused.STATE.reseed(seed)
        return self

    def training_state(self) -> dict:
        """What a restart needs besides `named_weights()`: optimizer state, schedule position, the dropout / SpecAugment counter."""
        from . import fused

        return {"iterations": self._iterations, "dropout": fused.STATE.state_dict(),
                "optimizer": self.optimizer.state_dict() if self.optimizer is not None else None}

    def load_training_state(self, state: dict):
        from . import fused

        self._iterations = int(state["iterations"])
        fused.STATE.load_state_dict(state["dropout"])
        if state.get("optimizer") is not None and self.optimizer is not None:
            self.optimizer.load_state_dict(state["optimizer"])
        return self

    def fit(self, x=None, epochs=1, steps_per_epoch=None, callbacks=None, initial_epoch=0, verbose=0, validation_data=None,
            validation_steps=None, **kwargs):
        """keras.Model.fit for this path (tensorflow_asr/scripts/train.py:113-121): drives `train_step` over an iterable of TrainData batches with the next
        batch's host->device copy on a copy stream and the per-step loss read back one step late (training.py).  Returns a History."""
        from . import training

        return training.fit(self, x, epochs=epochs, steps_per_epoch=steps_per_epoch, callbacks=callbacks, initial_epoch=initial_epoch,
                            verbose=verbose, validation_data=validation_data, validation_steps=validation_steps)

    @torch.no_grad()
    def test_step(self, data: schemas.TrainData):
        losses = self.compute_losses(data, training=False)
        return {"loss": losses}

    @torch.no_grad()
    def predict_step(self, data: schemas.TrainData):
        x, y_true = data
        batch_size = x.inputs.shape[0]
        inputs = schemas.PredictInput(
            inputs=x.inputs, inputs_length=x.inputs_length, previous_tokens=self.get_initial_tokens(batch_size=batch_size),
            previous_encoder_states=self.get_initial_encoder_states(batch_size=batch_size),
            previous_decoder_states=self.get_initial_decoder_states(batch_size=batch_size))
        _tokens = self.recognize(inputs=inputs).tokens
        _beam_tokens = self.recognize_beam(inputs=inputs).tokens
        return {"tokens": _tokens, "beam_tokens": _beam_tokens, "labels": y_true.labels}

    # ------------------------------------------------------------------ inference surface
    def get_initial_tokens(self, batch_size=1):
        return torch.zeros(batch_size, 1, dtype=torch.int32, device=next(self.parameters()).device)

    def get_initial_encoder_states(self, batch_size=1):
        return None

    def get_initial_decoder_states(self, batch_size=1):
        return None

    def recognize(self, inputs: schemas.PredictInput, **kwargs) -> schemas.PredictOutput:
        raise NotImplementedError()

    def recognize_beam(self, inputs: schemas.PredictInput, beam_width: int = 10, **kwargs) -> schemas.PredictOutput:
        raise NotImplementedError()

    # ------------------------------------------------------------------ weights by reference name
    def named_weights(self) -> dict:
        """{reference-style name: tensor}, e.g. 'encoder.block_0.ln.gamma', 'prediction.lstm_0.kernel', 'joint.vocab.kernel'."""
        out = {}
        for prefix, mod in self._weight_groups().items():
            for n, t in mod.named().items():
                out[f"{prefix}.{n}"] = t
        return out

    def load_named_weights(self, tensors: dict, strict=True):
        for prefix, mod in self._weight_groups().items():
            mod.load_named(tensors, prefix=prefix + ".", strict=strict)

    def _weight_groups(self) -> dict:
        groups = {}
        if hasattr(self, "encoder"):
            groups["encoder"] = self.encoder
            if hasattr(self.encoder, "conv_subsampling"):
                groups["encoder.subsampling"] = self.encoder.conv_subsampling
        if hasattr(self, "predict_net"):
            groups["prediction"] = self.predict_net
        if hasattr(self, "joint_net"):
            groups["joint"] = self.joint_net
        if hasattr(self, "decoder"):
            groups["decoder"] = self.decoder
        return groups


def _is_regularized(param_name: str) -> bool:
    """Keras: kernel_regularizer on kernels / embeddings, and on LN/BN gamma & beta; bias_regularizer is None."""
    n = param_name.replace("__", ".")
    if n.endswith("bias") or "attention_bias" in n:
        return False
    return n.endswith(("kernel", "gamma", "beta", "embeddings"))
