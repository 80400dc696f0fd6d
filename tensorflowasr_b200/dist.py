"""Data parallelism: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Replaces `tf.distribute.MirroredStrategy()` (tensorflow_asr/utils/env_util.py:57-70) and the implicit cross-replica gradient SUM
inside `optimizer.apply` (models/base_model.py:192).  Utterances are independent, so the batch shards over ranks with no
data-path collective; the only exchanges are (C1) the gradient all-reduce below, (C2) the sync-BatchNorm statistics
(`ops._SyncBatchNormFn`) and (C3) scalar metric means.

GradientReducer keeps every parameter's .grad as a view into ONE flat fp32 buffer, so the all-reduce needs no packing copies;
buckets (~25 MB) are reduced in reverse parameter order as soon as backward has produced all of their gradients (a
post-accumulate hook per parameter counts arrivals), overlapping NCCL with the rest of backward.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_from_env(backend: str | None = None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT). Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        return 0, 1, dev
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        if not dist.is_initialized():
            dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size(), dev


def shard_batch(n_global: int, rank_: int | None = None, world: int | None = None):
    """Even split of a global batch (reference: global = per-replica x replicas, drop_remainder; datasets.py:108,161)."""
    world = world_size() if world is None else world
    rank_ = rank() if rank_ is None else rank_
    per = n_global // world
    return rank_ * per, (rank_ + 1) * per


FLAT_ALIGN = 64  # elements: every parameter's slice of a flat buffer starts on a 256-byte boundary (TMA / float4 kernels address them)


def flat_layout(params, align: int = FLAT_ALIGN):
    """-> (offsets, total): element offsets of `params` in a flat buffer, each aligned to `align` elements; total padded likewise."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class GradientReducer:
    def __init__(self, params, bucket_bytes: int = 25 * 1024 * 1024, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, total = flat_layout(self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []  # (start, end, [param indices])
        cur_start, cur_idx = 0, []
        for i, p in enumerate(self.params):
            off, n = self.offsets[i], p.numel()
            p.grad = self.flat[off : off + n].view_as(p)
            cur_idx.append(i)
            end = self.offsets[i + 1] if i + 1 < len(self.params) else total
            if (end - cur_start) * 4 >= bucket_bytes:
                self.buckets.append((cur_start, end, cur_idx))
                cur_start, cur_idx = end, []
        if cur_idx:
            self.buckets.append((cur_start, total, cur_idx))
        self._bucket_of = {}
        for b, (_, _, idx) in enumerate(self.buckets):
            for i in idx:
                self._bucket_of[i] = b
        self._pending = [len(idx) for _, _, idx in self.buckets]
        self._works = []
        self._launched = [False] * len(self.buckets)
        self.overlap = overlap and world_size() > 1
        if self.overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(param):
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)

        return hook

    def _launch(self, b):
        s, e, _ = self.buckets[b]
        self._works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))
        self._launched[b] = True

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        """Finish the step's gradient SUM: launch whatever has not been launched by the hooks, then wait."""
        if world_size() > 1:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch(b)
            for w in self._works:
                w.wait()
        self._works = []
        self._launched = [False] * len(self.buckets)
        self._pending = [len(idx) for _, _, idx in self.buckets]


def all_reduce_mean_scalar(t: torch.Tensor) -> torch.Tensor:
    """C3: loss/metric mean over replicas (base_model.py:257-261)."""
    if world_size() > 1:
        t = t.clone()
        dist.all_reduce(t)
        t /= world_size()
    return t
