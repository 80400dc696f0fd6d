"""ContextNet (BASELINE config 4 family) on one B200: per-kernel timings of csrc/seqconv.cu with achieved GB/s, and the encoder train
step (fwd + bwd) on the fused route (fused_seq.py) against the library route.  CUDA events, L2 flushed between iterations.

    python tools/seqconv_bench.py [--batch 64] [--seconds 15] [--alpha 1.0] [--out gpurun_out/seqconv_bench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from microbench import peaks, time_ms  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", default="cuda", help="'cpu' = dry run on the emulator backend (tests/emu_backend.py must be installed by the caller)")
    args = ap.parse_args(argv)
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ctc_models, kernels, ops

    pk = peaks()
    dev = args.device
    on_gpu = dev.startswith("cuda")
    res = {"gpu": torch.cuda.get_device_name(0) if on_gpu else "cpu emulator (dry run)", "peaks": pk}
    B, T = args.batch, int(args.seconds * 100)

    def gbs(nbytes, ms):
        return {"ms": ms, "gbs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / pk["hbm_gbs"]}

    # ---- kernel level, C = 640 * alpha, K = 5
    C = int(640 * args.alpha) // 8 * 8
    x = torch.randn(B, T, C, device=dev).to(torch.bfloat16)
    w = torch.randn(5, C, device=dev) * 0.3
    dw = torch.zeros(5, C, device=dev)
    length = torch.full((B,), T, dtype=torch.int32, device=dev)
    act_bytes = B * T * C * 2
    for s in (1, 2):
        med, _ = time_ms(lambda: kernels.seq_dwconv_fwd(x, w, s))
        res[f"seq_dwconv_fwd_s{s}"] = gbs(act_bytes + act_bytes // s, med)
        dy = torch.randn(B, -(-T // s), C, device=dev).to(torch.bfloat16)
        med, _ = time_ms(lambda: kernels.seq_dwconv_bwd(dy, x, w, dw, s))
        res[f"seq_dwconv_bwd_s{s}"] = gbs(2 * act_bytes + 2 * (act_bytes // s), med)  # dx pass: read dy, write dx; dw pass: read dy, x
    med, _ = time_ms(lambda: kernels.seq_masked_mean(x, length))
    res["seq_masked_mean"] = gbs(act_bytes, med)
    H = C // 8
    W1, b1 = torch.randn(C, H, device=dev) / C ** 0.5, torch.zeros(H, device=dev)
    W2, b2 = torch.randn(H, C, device=dev) / H ** 0.5, torch.zeros(C, device=dev)
    m = kernels.seq_masked_mean(x, length)
    med, _ = time_ms(lambda: kernels.se_fc_fwd(m, W1, b1, W2, b2, 1))
    res["se_fc_fwd"] = {"ms": med}
    z1, sg = kernels.se_fc_fwd(m, W1, b1, W2, b2, 1)
    r = torch.randn_like(x)
    med, _ = time_ms(lambda: kernels.se_apply_fwd(x, sg, r, 1))
    res["se_apply_fwd"] = gbs(3 * act_bytes, med)
    med, _ = time_ms(lambda: kernels.se_apply_bwd(r, x, sg, r, 1))
    res["se_apply_bwd"] = gbs(4 * act_bytes, med)
    dres, ds = kernels.se_apply_bwd(r, x, sg, r, 1)
    g = [torch.zeros_like(t) for t in (W1, b1, W2, b2)]
    med, _ = time_ms(lambda: kernels.se_fc_bwd(ds, sg, z1, m, W1, W2, length, T, 1, *g))
    res["se_fc_bwd"] = {"ms": med}
    dm = kernels.se_fc_bwd(ds, sg, z1, m, W1, W2, length, T, 1, *g)
    med, _ = time_ms(lambda: kernels.se_dy(dres, sg, dm, length))
    res["se_dy"] = gbs(2 * act_bytes, med)
    del x, r, dres

    # ---- encoder train step, fused route vs library route
    ops.set_precision("bf16")
    cfg = tfasr.contextnet_config(vocab_size=1000, alpha=args.alpha)
    enc = ctc_models.ContextNetEncoder(blocks=cfg["encoder_blocks"], alpha=args.alpha, num_feature_bins=80, gen=torch.Generator().manual_seed(0)).to(dev)
    feats = torch.randn(B, T, 80, 1, device=dev)
    flen = torch.full((B,), T, dtype=torch.int32, device=dev)

    def step():
        for p in enc.parameters():
            p.grad = None
        y, _ = enc((feats, flen), training=True)
        y.float().square().mean().backward()

    for flag in (False, True):
        ops.set_fused_seq(flag)
        try:
            med, best = time_ms(step, iters=5, warmup=2, flush=False)
            res["encoder_step_fused" if flag else "encoder_step_library"] = {"ms": med, "ms_best": best, "utt_per_s": B / med * 1e3,
                                                                            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30 if on_gpu else None}
        except Exception as e:  # keep the other numbers if one route fails
            res["encoder_step_fused" if flag else "encoder_step_library"] = {"error": repr(e)[:300]}
        if on_gpu:
            torch.cuda.reset_peak_memory_stats()
    ops.set_fused_seq(False)
    print(json.dumps(res, indent=1))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
    return res


if __name__ == "__main__":
    main()
