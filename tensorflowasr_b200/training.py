"""The training loop around `train_step`: what `keras.Model.fit` does for the reference (`tensorflow_asr/scripts/train.py:113-121` calls
`model.fit(train_dataset, epochs, steps_per_epoch, callbacks, initial_epoch)`; the step function it drives is `base_model.py:194-269`).

B200-first shape of that loop — nothing here computes, it only keeps the device busy:
  * host batches (pinned memory) go to the device on a COPY stream, one batch ahead of the step that is running (the role of tf.data's
    prefetch + keras' data adapter); the step's stream waits on the copy's event only,
  * the step's loss is read back asynchronously into pinned memory and consumed ONE step later, so the host never drains the queue between
    steps (keras reads its logs lazily in the same way); every step's loss is still read, and `fit` returns only after the last one arrived,
  * with `enable_cuda_graph` the step itself is one graph replay, so a step costs the host a few microseconds.
"""
from __future__ import annotations

import torch

from . import schemas


class History:
    """keras.callbacks.History: `.history[name]` = one value per epoch, `.epoch` = the epoch indices; `.steps["loss"]` adds the per-step values."""

    def __init__(self):
        self.history: dict = {}
        self.epoch: list = []
        self.steps: dict = {"loss": []}


def _map(fn, data: schemas.TrainData) -> schemas.TrainData:
    i, l = data.inputs, data.labels
    f = lambda t: None if t is None else fn(t)  # noqa: E731
    return schemas.TrainData(schemas.TrainInput(f(i.inputs), f(i.inputs_length), f(i.predictions), f(i.predictions_length)),
                             schemas.TrainLabel(f(l.labels), f(l.labels_length)))


def _tensors(data: schemas.TrainData):
    i, l = data.inputs, data.labels
    return [t for t in (i.inputs, i.inputs_length, i.predictions, i.predictions_length, l.labels, l.labels_length) if t is not None]


class DevicePrefetcher:
    """Iterates TrainData batches, host -> device one batch ahead on its own stream.  Batches already on the device pass through.

    The device side is a ring of `depth` preallocated staging sets (no allocation per step): set k is rewritten by the copy stream only after
    the step that consumed it has been enqueued and finished on the consumer's stream (an event recorded at the following __next__)."""

    def __init__(self, batches, device: torch.device, depth: int = 2):
        self.it = iter(batches)
        self.dev = device
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.bytes_moved = 0
        self.ring = [None] * depth      # staging tensors per slot
        self.consumed = [None] * depth  # event: the consumer of the slot's previous content is done
        self.slot = 0
        self._handed = None  # slot handed out by the last __next__ (its consumer is enqueued by the time __next__ runs again)
        self._next = None
        self._load()

    def _staging(self, k, ts):
        ring = self.ring[k]
        if ring is None or len(ring) != len(ts) or any(r.shape != t.shape or r.dtype != t.dtype for r, t in zip(ring, ts)):
            ring = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in ts]
            self.ring[k] = ring
            self.consumed[k] = None
        return ring

    def _load(self):
        try:
            data = next(self.it)
        except StopIteration:
            self._next = None
            return
        ts = _tensors(data)
        if self.stream is None or all(t.device == self.dev for t in ts):
            self._next = (data if self.stream is not None else _map(lambda t: t.to(self.dev), data), None, None)
            return
        k = self.slot
        self.slot = (k + 1) % len(self.ring)
        ring = self._staging(k, ts)
        if self.consumed[k] is not None:
            self.stream.wait_event(self.consumed[k])
        with torch.cuda.stream(self.stream):
            for r, t in zip(ring, ts):
                r.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.bytes_moved += sum(t.numel() * t.element_size() for t in ts)
        it = iter(ring)
        moved = _map(lambda t: next(it), data)
        self._next = (moved, ev, (data, k))  # the host batch stays referenced until its copy has been consumed

    def __iter__(self):
        return self

    def __next__(self) -> schemas.TrainData:
        if self._handed is not None:  # everything enqueued since the last hand-out is that slot's consumer
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            self.consumed[self._handed] = ev
            self._handed = None
        cur = self._next
        if cur is None:
            raise StopIteration
        data, ev, host = cur
        if ev is not None:
            torch.cuda.current_stream(self.dev).wait_event(ev)
            self._handed = host[1]
        self._load()  # the next batch's copy overlaps the step that is about to be enqueued
        return data


class _LossReader:
    """Per-step scalar loss -> pinned host memory, asynchronously, `depth` steps deep."""

    def __init__(self, device: torch.device, depth: int = 2):
        self.cuda = device.type == "cuda"
        self.slots = [torch.zeros(1, dtype=torch.float32).pin_memory() if self.cuda else torch.zeros(1) for _ in range(depth)]
        self.events = [torch.cuda.Event() if self.cuda else None for _ in range(depth)]
        self.pending: list = []
        self.count = 0
        self.d2h_bytes = 0

    def push(self, losses: torch.Tensor):
        assert not self.full()
        k = self.count % len(self.slots)
        self.count += 1
        self.slots[k].copy_(losses.detach().float().mean().reshape(1), non_blocking=True)
        if self.cuda:
            self.events[k].record()
        self.d2h_bytes += 4
        self.pending.append(k)

    def pop(self) -> float:
        k = self.pending.pop(0)
        if self.cuda:
            self.events[k].synchronize()
        return float(self.slots[k][0])

    def full(self) -> bool:
        return len(self.pending) >= len(self.slots)


def _call(callbacks, name, *args):
    for cb in callbacks or ():
        fn = getattr(cb, name, None)
        if fn is not None:
            fn(*args)


def _validate(model, validation_data, validation_steps, dev) -> float | None:
    """Mean of `test_step` losses over the validation batches (keras: evaluate at the end of every epoch) — same prefetcher, same late reads."""
    source = DevicePrefetcher(validation_data, dev)
    reader = _LossReader(dev)
    vals, step = [], 0
    while validation_steps is None or step < validation_steps:
        try:
            data = next(source)
        except StopIteration:
            break
        reader.push(model.test_step(data)["loss"])
        step += 1
        if reader.full():
            vals.append(reader.pop())
    while reader.pending:
        vals.append(reader.pop())
    return sum(vals) / len(vals) if vals else None


def fit(model, x, epochs=1, steps_per_epoch=None, callbacks=None, initial_epoch=0, verbose=0, validation_data=None,
        validation_steps=None) -> History:
    """`x`: an iterable of schemas.TrainData batches (host — ideally pinned — or device tensors), re-iterated every epoch, or an iterator when
    `steps_per_epoch` bounds each epoch (an exhausted iterator ends training, like keras).  `validation_data` (re-iterable, same batch type) is
    evaluated with `test_step` at the end of every epoch -> `val_loss` (tensorflow_asr/scripts/train.py:113-121 passes both).  Returns the
    History; callbacks may implement on_train_begin / on_epoch_begin(epoch) / on_train_batch_end(step, logs) / on_epoch_end(epoch, logs) /
    on_train_end."""
    dev = next(model.parameters()).device
    hist = History()
    model.train()
    _call(callbacks, "on_train_begin")
    source = None
    moved = read = 0
    for epoch in range(initial_epoch, epochs):
        _call(callbacks, "on_epoch_begin", epoch)
        if source is None or (steps_per_epoch is None):
            source = DevicePrefetcher(x, dev)
        reader = _LossReader(dev)
        losses, step = [], 0

        def drain_one():
            v = reader.pop()
            losses.append(v)
            _call(callbacks, "on_train_batch_end", len(losses) - 1, {"loss": v})

        exhausted = False
        while steps_per_epoch is None or step < steps_per_epoch:
            try:
                data = next(source)
            except StopIteration:
                exhausted = True
                break
            out = model.train_step(data)
            reader.push(out["loss"])
            step += 1
            if reader.full():
                drain_one()  # the loss of the step before the one just enqueued
        while reader.pending:
            drain_one()
        moved += source.bytes_moved
        source.bytes_moved = 0
        read += reader.d2h_bytes
        if not losses:
            break
        logs = {"loss": sum(losses) / len(losses)}
        if validation_data is not None:
            val = _validate(model, validation_data, validation_steps, dev)
            if val is not None:
                logs["val_loss"] = val
                hist.history.setdefault("val_loss", []).append(val)
            model.train()
        hist.epoch.append(epoch)
        hist.history.setdefault("loss", []).append(logs["loss"])
        hist.steps["loss"].extend(losses)
        if verbose:
            print(f"epoch {epoch + 1}/{epochs}: {len(losses)} steps, " + ", ".join(f"{k} {v:.4f}" for k, v in logs.items()))
        _call(callbacks, "on_epoch_end", epoch, logs)
        if exhausted and steps_per_epoch is not None:
            break
    _call(callbacks, "on_train_end")
    hist.h2d_bytes, hist.d2h_bytes = moved, read
    return hist
