"""GPU check of the peer-memory all-reduce (csrc/peer_allreduce.cu, dist.PeerAllReduce) on ONE GPU: a world of one rank still
runs the whole protocol — push into the own window (a cudaMalloc'ed block with an IPC handle), release / acquire the own flag, sum — eagerly
and as a replayed CUDA graph (the call counter lives on the device).  The multi-GPU run is tools/peer_bn_check.py (torchrun, 2+ GPUs).

The single-rank protocol ran green on a B200 in round 2; the multi-GPU run is tools/peer_bn_check.py and tests/test_gpu_peer_multi.py (torchrun, 2+ GPUs).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_single_rank_world_runs_the_protocol_eagerly_and_in_a_graph():
    from tensorflowasr_b200 import dist as tdist

    comm = tdist.PeerAllReduce(capacity=1280)
    try:
        x = torch.randn(1280, device="cuda")
        ref = x.clone()
        for n in (288, 1280, 2, 641):
            y = x[:n].clone()
            comm.all_reduce(y)
            assert torch.equal(y, ref[:n])
        assert comm.check() == 4
        static = torch.zeros(288, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            comm.all_reduce(static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(6):
                comm.all_reduce(static)
        for k in range(3):
            static.copy_(ref[:288] + k)
            graph.replay()
            assert torch.equal(static, ref[:288] + k)
        assert comm.check() == 5 + 18
        with pytest.raises(ValueError):
            comm.all_reduce(torch.zeros(1281, device="cuda"))
    finally:
        comm.close()
