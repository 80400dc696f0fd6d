"""CPU, world_size 2 (gloo) on the emulator backend: the FUSED Conformer path under data parallelism — batch sharding, the synchronized
BatchNorm statistics exchanged by fused.bn_train_fwd / bn_train_bwd around the BatchNorm kernels (C2), loss / num_replicas and summed
gradients (C1; base_model.py:166-173) — must reproduce the single-process full-batch gradients of the same fused path; and the
same step with the statistics exchanged through the peer-memory all-reduce kernel (ops.set_peer_bn) must agree with the gloo exchange."""
import os
import shutil
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available (emulator build)")


class _Patch:
    """monkeypatch.setattr for a spawned worker (no pytest fixture there; the process ends with the test).  In the pytest process itself
    `undo()` puts everything back."""

    def __init__(self):
        self._saved = []

    def setattr(self, obj, name, value):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def undo(self):
        for obj, name, value in reversed(self._saved):
            setattr(obj, name, value)
        self._saved = []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(B=4, N=8000, U=3, V=16):
    from tensorflowasr_b200 import schemas

    g = torch.Generator().manual_seed(0)
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.tensor([N, N - 1700, N - 300, N - 2500], dtype=torch.int32)[:B]
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    lab_len = torch.tensor([U, U - 1, U, U - 2], dtype=torch.int32)[:B]
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)

    def shard(lo, hi):
        return schemas.TrainData(schemas.TrainInput(sig[lo:hi], sig_len[lo:hi], pred[lo:hi], lab_len[lo:hi] + 1),
                                 schemas.TrainLabel(labels[lo:hi], lab_len[lo:hi]))

    return shard


def _fused_grads(data, world):
    import emu_backend
    from test_conformer_emulated_cpu import _model

    from tensorflowasr_b200 import ops

    patch = _Patch()
    try:
        emu_backend.install(patch)
        ops.set_precision("bf16")
        model = _model()
        model.train()
        losses = model.compute_losses(data, training=True)
        (losses.mean() / world).backward()
        return losses.detach(), {n: p.grad.detach().clone().float() for n, p in model.named_parameters() if p.grad is not None}
    finally:
        patch.undo()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensorflowasr_b200 import dist as tdist

    from tensorflowasr_b200 import ops as _ops

    lo, hi = tdist.shard_batch(4)
    _ops.set_peer_bn(False)  # first the gloo / NCCL-style exchange (the peer-memory kernel is the default route on GPUs)
    losses, grads = _fused_grads(_data()(lo, hi), world)
    # the same step once more with the sync-BatchNorm statistics exchanged by the peer-memory kernel (csrc/peer_allreduce.cu on the emulator;
    # the windows are shared-memory blocks mapped at different addresses in the two processes, as CUDA IPC mappings would be)
    from tensorflowasr_b200 import ops

    ops.set_peer_bn(True)
    try:
        losses_p, grads_p = _fused_grads(_data()(lo, hi), world)
        calls = tdist.peer_all_reduce_check()
    finally:
        ops.set_peer_bn(False)
    assert calls is not None and calls >= 6, calls  # 2 subsampling + 1 per block BatchNorms, forward and backward
    torch.testing.assert_close(losses_p, losses, rtol=1e-3, atol=1e-3)
    for n, g in grads.items():
        if g.norm() > 1e-5 * max(1.0, g.numel() ** 0.5):
            cos = float((grads_p[n] * g).sum() / (grads_p[n].norm() * g.norm() + 1e-30))
            assert cos > 0.999, (n, cos)  # not bit-identical: the kernels' float atomics commit in a different order from run to run
    tdist._PEER.close()
    for n in sorted(grads):
        dist.all_reduce(grads[n])  # gradients are SUMMED over replicas
    if rank == 0:  # plain numpy through the queue: torch's shared-memory hand-over needs the producer alive at receive time
        out.put((losses.numpy(), {n: g.numpy() for n, g in grads.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_fused_path_matches_single_process_full_batch():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    losses0, grads = out.get(timeout=600)
    losses0, grads = torch.from_numpy(losses0), {n: torch.from_numpy(g) for n, g in grads.items()}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    try:
        full_losses, full = _fused_grads(_data()(0, 4), 1)
    finally:
        from tensorflowasr_b200 import ops

        ops.set_precision("fp32")
    # rank 0's utterances see the same statistics as in the full batch: identical losses up to bf16 noise of the different launch geometry
    torch.testing.assert_close(losses0, full_losses[:2], rtol=2e-2, atol=2e-2)
    checked = 0
    for n, g in full.items():
        if g.norm() < 1e-5 * max(1.0, g.numel() ** 0.5):
            continue
        cos = float((grads[n] * g).sum() / (grads[n].norm() * g.norm() + 1e-30))
        assert cos > 0.99, (n, cos)
        assert abs(float(grads[n].norm() / g.norm()) - 1.0) < 0.1, (n, float(grads[n].norm() / g.norm()))
        checked += 1
    assert checked > 40, checked
