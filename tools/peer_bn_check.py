"""Multi-GPU check + microbenchmark of the peer-memory all-reduce (csrc/peer_allreduce.cu) against NCCL, and of the headline step with it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/peer_bn_check.py \
        --out gpurun_out/peer_bn_check.json

1. correctness: 200 calls of random sizes, every rank's result == the NCCL all-reduce of the same inputs (bit-exact for 2 ranks; rank-order
   sums vs NCCL's order differ in the last bit for more, so the bound is 1e-6 relative there), eagerly and replayed from a CUDA graph;
2. latency: a CUDA graph of 36 dependent small all-reduces (the sync-BatchNorm pattern of one Conformer-S step: 288 floats each), NCCL vs
   peer kernel, CUDA-event time per replay, max over ranks;
3. health: PeerAllReduce.check() (no peer ever timed out).
Rank 0 prints / writes one JSON object."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _time_graph(fn, buf, reps=50):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn(buf)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(36):
            fn(buf)
            buf.mul_(1.0 / dist.get_world_size())  # a dependent kernel between the exchanges (as BatchNorm's finalize / apply are); keeps the values bounded
    for _ in range(5):
        g.replay()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--floats", type=int, default=288)
    args = ap.parse_args()
    from tensorflowasr_b200 import dist as tdist

    rank, world, dev = tdist.init_from_env()
    assert world > 1, "run under torchrun with 2+ GPUs"
    comm = tdist.PeerAllReduce(capacity=4096)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    sizes = torch.randint(1, 4097, (200,), generator=torch.Generator().manual_seed(7)).tolist()  # the same sequence on every rank
    worst = 0.0
    for n in sizes:
        x = torch.randn(n, device="cuda", generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        y = comm.all_reduce(x.clone())
        err = float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
        worst = max(worst, err)
    tol = 0.0 if world == 2 else 1e-6
    ok_eager = worst <= tol
    # every rank must hold bit-identical sums (replicas must not drift): compare against rank 0's
    x = torch.randn(4096, device="cuda", generator=g)
    y = comm.all_reduce(x.clone())
    y0 = y.clone()
    dist.broadcast(y0, 0)
    same = bool(torch.equal(y, y0))
    buf = torch.randn(args.floats, device="cuda", generator=g)
    t_nccl = _time_graph(lambda b: dist.all_reduce(b), buf.clone())
    t_peer = _time_graph(lambda b: comm.all_reduce(b), buf.clone())
    calls = comm.check()
    res = {"world": world, "eager_max_rel_err": worst, "eager_ok": ok_eager, "identical_across_ranks": same, "calls": calls,
           "graph36_ms": {"nccl": t_nccl, "peer": t_peer}, "us_per_exchange": {"nccl": t_nccl / 36 * 1e3, "peer": t_peer / 36 * 1e3},
           "floats": args.floats}
    if rank == 0:
        print(json.dumps(res))
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(res, f)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if not (ok_eager and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
