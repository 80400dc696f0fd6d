"""Mint tests/golden/conformer_s_config3_fp64.npz: the float64 CPU oracle (oracle/conformer.py, oracle/logmel.py) run ONCE at BASELINE
config-3 size — Conformer-S (16 blocks, d=144), B=32 x 10 s, U=64, V=1000, training mode (batch statistics), dropout 0, SpecAugment off —
forward + backward, so that the GPU suite can hold the benchmarked bf16 path against it without spending minutes of CPU time per run.

TEST INFRASTRUCTURE.  Inputs and weights are regenerated from seeds by tests/test_gpu_bf16_oracle.py (same helper: `full_size_case`),
so the fixture only stores what the oracle produced: per-utterance losses, logit statistics, and the gradients of a list of
parameters (large tensors as a strided sample).

    python tools/mint_full_size_golden.py [--batch 32] [--out tests/golden/conformer_s_config3_fp64.npz]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import conformer as oc  # noqa: E402
from oracle import logmel as ologmel  # noqa: E402

GRAD_NAMES = ["encoder.linear.kernel", "encoder.block_0.mhsa_module.mhsa.query.kernel", "encoder.block_7.conv_module.dw_conv.kernel",
              "encoder.block_0.ff_module_1.dense_1.bias", "encoder.block_15.ff_module_2.dense_2.kernel", "encoder.content_attention_bias",
              "encoder.positional_attention_bias", "encoder.subsampling.conv_0.kernel", "encoder.subsampling.conv_1.kernel",
              "encoder.block_8.conv_module.dw_bn.gamma", "encoder.block_3.conv_module.pw_conv_1.kernel", "prediction.lstm_0.recurrent_kernel",
              "prediction.embedding.embeddings", "joint.vocab.kernel", "joint.vocab.bias", "joint.enc.kernel", "joint.pred.bias",
              "encoder.block_15.ln.beta", "encoder.block_12.mhsa_module.mhsa.encoding.kernel", "encoder.block_5.mhsa_module.mhsa.attention_output.kernel"]
SAMPLE = 4096  # elements kept of a large gradient (evenly strided)


def full_size_case(B=32, seconds=10, U=64, V=1000, seed=20260923):
    """-> (ocfg, params fp32 dict, (sig, sig_len, pred, pred_len, labels, lab_len)) — deterministic on the CPU generator."""
    ocfg = oc.ConformerConfig(vocab_size=V)
    params = oc.init_params(ocfg, seed=seed % 1000, randomize_all=True)
    g = torch.Generator().manual_seed(seed)
    N = 16000 * seconds
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.full((B,), N, dtype=torch.int32)
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    lab_len = torch.full((B,), U, dtype=torch.int32)
    # a ragged tail: the last quarter of the batch is shorter (lengths still dominated by full utterances, like the bench)
    for i in range(B - B // 4, B):
        sig_len[i] = N - 4000 * (i - (B - B // 4) + 1)
        lab_len[i] = U - 3 * (i - (B - B // 4) + 1)
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)
    return ocfg, params, (sig, sig_len, pred, lab_len + 1, labels, lab_len)


def sample(t: torch.Tensor) -> np.ndarray:
    f = t.detach().reshape(-1).double().numpy()
    if f.size <= SAMPLE:
        return f
    idx = np.linspace(0, f.size - 1, SAMPLE).astype(np.int64)
    return f[idx]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=int, default=10)
    ap.add_argument("--labels", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "conformer_s_config3_fp64.npz"))
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    ocfg, params, host = full_size_case(a.batch, a.seconds, a.labels)
    sig, sig_len, pred, pred_len, labels, lab_len = host
    t0 = time.time()
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.double().clone() for k, v in params.items()}
    for k in GRAD_NAMES:
        p[k].requires_grad_(True)
    logits, llen = oc.transducer_forward(p, ocfg, torch.from_numpy(feats), torch.from_numpy(flen).long(), pred, pred_len.long(), training=True)
    print(f"forward {time.time() - t0:.1f} s, logits {tuple(logits.shape)}", flush=True)
    losses = oc.rnnt_loss_torch(logits, labels, llen.int(), lab_len)
    print(f"loss {time.time() - t0:.1f} s: {losses[:4].tolist()}", flush=True)
    losses.mean().backward()
    print(f"backward {time.time() - t0:.1f} s", flush=True)
    out = {"losses": losses.detach().numpy(), "logits_length": llen.numpy(),
           "logits_mean": np.float64(logits.detach().mean().item()), "logits_std": np.float64(logits.detach().std().item()),
           "logits_sample": sample(logits), "batch": np.int64(a.batch), "seconds": np.int64(a.seconds), "labels": np.int64(a.labels)}
    for k in GRAD_NAMES:
        out["grad:" + k] = sample(p[k].grad)
        out["gmax:" + k] = np.float64(p[k].grad.abs().max().item())
    np.savez_compressed(a.out, **out)
    print(f"wrote {a.out} ({os.path.getsize(a.out) / 1e6:.2f} MB) in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
