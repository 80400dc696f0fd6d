"""Two real GPUs: the peer-memory sync-BatchNorm exchange (csrc/peer_allreduce.cu over CUDA-IPC windows and NVLink stores) against NCCL —
every rank's sums equal NCCL's, identical on all ranks, eagerly and replayed from a CUDA graph of 36 dependent exchanges — by launching
tools/peer_bn_check.py under torchrun.  Skipped on boxes with one GPU (run it with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_peer_allreduce_matches_nccl_on_two_gpus(tmp_path):
    out = tmp_path / "peer.json"
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "peer_bn_check.py"), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["world"] == 2 and res["eager_ok"] and res["identical_across_ranks"] and res["eager_max_rel_err"] == 0.0
    assert res["calls"] > 200
    print(res)
