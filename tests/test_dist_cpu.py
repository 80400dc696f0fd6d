"""CPU, world_size 2 (gloo): the data-parallel host logic — batch sharding, the flat-buffer bucketed gradient all-reduce (C1) and the
synchronized batch-norm statistics (C2) — must reproduce single-process full-batch gradients (reference arrangement:
loss / num_replicas, gradients SUMMED; models/base_model.py:166-173)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(seed=0):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.nn.Parameter(torch.randn(8, 6, generator=g) * 0.3), torch.nn.Parameter(torch.randn(6, generator=g) * 0.1),
          torch.nn.Parameter(torch.rand(6, generator=g) + 0.5), torch.nn.Parameter(torch.randn(6, generator=g) * 0.1),
          torch.nn.Parameter(torch.randn(6, 3, generator=g) * 0.3)]
    return ws


def _forward(ws, x):
    from tensorflowasr_b200 import ops

    ops.set_precision("fp32")
    w1, b1, gamma, beta, w2 = ws
    h = ops.linear(x, w1, b1)
    mm, mv = torch.zeros(6), torch.ones(6)
    h = ops.batch_norm(h, gamma, beta, mm, mv, training=True)  # sync-BN: statistics over BOTH ranks' rows
    return (ops.linear(torch.nn.functional.silu(h), w2) ** 2).sum(-1)  # per-utterance "loss"


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensorflowasr_b200 import dist as tdist

    torch.manual_seed(123)
    x = torch.randn(8, 8)
    lo, hi = tdist.shard_batch(8)
    assert (lo, hi) == (rank * 4, rank * 4 + 4)
    ws = _model()
    red = tdist.GradientReducer(ws, bucket_bytes=64, overlap=True)  # tiny buckets: exercise several async all-reduces
    assert len(red.buckets) >= 3
    losses = _forward(ws, x[lo:hi])
    (losses.mean() / world).backward()  # base_model.py:166-173
    red.all_reduce()
    mean_loss = tdist.all_reduce_mean_scalar(losses.mean().detach())
    if rank == 0:
        out.put(([w.grad.clone().numpy() for w in ws], float(mean_loss)))  # by value: a torch tensor would travel as a file descriptor this process must outlive
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_single_process_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    grads, mean_loss = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(123)
    x = torch.randn(8, 8)
    ws = _model()
    losses = _forward(ws, x)
    losses.mean().backward()
    torch.testing.assert_close(torch.tensor(mean_loss), losses.mean().detach(), rtol=1e-5, atol=1e-6)
    for g, w in zip(grads, ws):
        torch.testing.assert_close(torch.from_numpy(g), w.grad, rtol=1e-4, atol=1e-6)
