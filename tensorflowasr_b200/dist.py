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
        # gradient accumulation: micro-steps must not start collectives on buffers later micro-batches still accumulate into;
        # `defer` switches the overlap hooks off, the one all_reduce() of the apply step then reduces every bucket
        self.defer = False
        if self.overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(param):
            if self.defer:
                return
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


class PeerAllReduce:
    """All-reduce (SUM, fp32, in place) of small vectors over peer memory: `csrc/peer_allreduce.cu` — one launch of one CTA per call in place of
    an NCCL all-reduce; built for the sync-BatchNorm statistics (C2: 2 x C floats, twice per BatchNorm and step, all on the critical path).
    Every rank owns one window (device memory shared over CUDA IPC; the 64-byte handles travel through the process group) and maps every
    peer's.  Every rank must make the same sequence of calls on its current stream; the call counter lives on the device, so the calls can be
    captured in a CUDA graph.  `check()` (host sync) raises if a peer ever failed to arrive within the kernel's bounded wait."""

    def __init__(self, capacity: int = 4096, device=None, timeout_ms: int = 0):
        import ctypes

        from . import _lib

        lib = _lib.load()
        self.world, self.rank, self.capacity = world_size(), rank(), int(capacity)
        nbytes = int(lib.tfasr_peer_window_bytes(self.world, self.capacity))
        if nbytes == 0:
            raise ValueError(f"PeerAllReduce: unsupported world size {self.world} / capacity {capacity}")
        self._owned, self._mapped, self._shm = None, [], []
        if _lib.EMULATED:
            ptrs = self._open_host_windows(nbytes)
            device = torch.device("cpu")
        else:
            device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            ptrs = self._open_ipc_windows(lib, nbytes, ctypes)
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        self._comm = _lib.PeerComm(self.table.data_ptr(), self.state.data_ptr(), self.rank, self.world, self.capacity, int(timeout_ms))
        if self.world > 1:
            dist.barrier()  # every window is zero-filled and mapped everywhere before the first push

    def _open_ipc_windows(self, lib, nbytes, ctypes):
        from . import _lib

        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.tfasr_peer_window_alloc(nbytes, ctypes.byref(own), handle), "tfasr_peer_window_alloc")
        self._owned = own.value
        import socket

        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, (socket.gethostname(), bytes(handle.raw)))
            if len({h for h, _ in handles}) != 1:  # CUDA IPC (and NVLink peer stores) end at the node boundary
                self.close()
                raise RuntimeError("PeerAllReduce: all ranks must run on one node (NVLink peer memory); use the NCCL exchange (ops.set_peer_bn(False))")
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(self._owned)
                continue
            p = ctypes.c_void_p()
            _lib.check(lib.tfasr_peer_window_open(ctypes.create_string_buffer(handles[r][1], 64), ctypes.byref(p)), "tfasr_peer_window_open")
            self._mapped.append(p.value)
            ptrs.append(p.value)
        return ptrs

    def _open_host_windows(self, nbytes):
        """CPU emulator backend (tests): the windows are POSIX shared-memory blocks, mapped at a different address in every process —
        exactly the situation of CUDA IPC mappings."""
        import ctypes
        from multiprocessing import shared_memory

        own = shared_memory.SharedMemory(create=True, size=nbytes)
        own.buf[:nbytes] = bytes(nbytes)
        names = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(names, own.name)
        else:
            names = [own.name]
        ptrs = []
        for r, name in enumerate(names):
            shm = own if r == self.rank else shared_memory.SharedMemory(name=name)
            self._shm.append(shm)
            ptrs.append(ctypes.addressof(ctypes.c_char.from_buffer(shm.buf)))
        self._own_shm = own
        return ptrs

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        from . import _lib

        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() > self.capacity:
            raise ValueError(f"PeerAllReduce.all_reduce: contiguous fp32 of at most {self.capacity} elements expected, got {t.dtype} x {t.numel()}")
        import ctypes

        lib = _lib.load()
        _lib.check(lib.tfasr_peer_allreduce_f32(ctypes.byref(self._comm), _lib.ptr(t), t.numel(), _lib.stream_ptr(t.device)), "tfasr_peer_allreduce_f32")
        return t

    def check(self):
        calls, late = (int(v) for v in self.state.tolist())
        if late:
            raise RuntimeError(f"PeerAllReduce: a peer did not arrive within the bounded wait (rank {self.rank}, {calls} calls completed)")
        return calls

    def close(self):
        from . import _lib

        if self._shm:
            import gc

            self.table = None
            gc.collect()
            for shm in self._shm:
                try:
                    shm.close()
                except BufferError:
                    pass
            try:
                self._own_shm.unlink()
            except FileNotFoundError:
                pass
            self._shm = []
            return
        lib = _lib.load()
        for p in self._mapped:
            lib.tfasr_peer_window_close(p, 0)
        self._mapped = []
        if self._owned:
            lib.tfasr_peer_window_close(self._owned, 1)
            self._owned = None


_PEER = None
_PEER_FAILED = False


def _lib_emulated() -> bool:
    from . import _lib

    return _lib.EMULATED


def small_all_reduce(t: torch.Tensor) -> torch.Tensor:
    """C2: in-place SUM of a small fp32 vector over the replicas — NCCL by default, the peer-memory kernel with ops.set_peer_bn(True)."""
    if world_size() <= 1:
        return t
    from . import ops

    if ops.peer_bn_active() and t.dtype == torch.float32 and t.is_contiguous():
        global _PEER, _PEER_FAILED
        if (_PEER is None or _PEER.world != world_size()) and not _PEER_FAILED and (t.is_cuda or _lib_emulated()):
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PeerAllReduce must be created before CUDA-graph capture (run one eager step first)")
            try:
                _PEER = PeerAllReduce(capacity=max(4096, t.numel()))
            except RuntimeError:  # ranks on several nodes, or no peer access: the NCCL exchange stays
                _PEER_FAILED = True
        if _PEER is not None and not _PEER_FAILED and t.numel() <= _PEER.capacity:
            return _PEER.all_reduce(t)
    dist.all_reduce(t)
    return t


def peer_all_reduce_check():
    """Host-side health check of the peer-memory exchange (None when it is not in use)."""
    return None if _PEER is None else _PEER.check()


def all_reduce_mean_scalar(t: torch.Tensor) -> torch.Tensor:
    """C3: loss/metric mean over replicas (base_model.py:257-261)."""
    if world_size() > 1:
        t = t.clone()
        dist.all_reduce(t)
        t /= world_size()
    return t
