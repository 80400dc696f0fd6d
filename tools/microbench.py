"""Per-kernel timings with CUDA events (L2 flushed between iterations) -> achieved GB/s vs MEASURED_PEAKS.json.

    python tools/microbench.py [--out gpurun_out/microbench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "fallback": True}


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    _flush.fill_(1)


def time_ms(fn, iters=10, warmup=3, flush=True):
    if not torch.cuda.is_available():  # CPU dry run of a tool on the emulator backend (tests): wall clock, no L2 to flush
        import time

        ts = []
        for _ in range(max(1, min(iters, 2))):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        return ts[len(ts) // 2], ts[0]
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from tensorflowasr_b200 import FeatureExtraction, _lib, rnnt_loss

    pk = peaks()
    res = {"peaks": pk, "gpu": torch.cuda.get_device_name(0)}
    lib = _lib.load()

    # log-mel: B x 10 s
    fe = FeatureExtraction()
    for B in (32, 256):
        sig = (0.1 * torch.randn(B, 160000, device="cuda")).clamp(-1, 1)
        lens = torch.full((B,), 160000, dtype=torch.int32, device="cuda")
        med, best = time_ms(lambda: fe((sig, lens)))
        nbytes = B * 960000
        res[f"logmel_B{B}"] = dict(ms=med, ms_best=best, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / pk["hbm_gbs"], utt_per_s=B / med * 1e3)

    # RNN-T loss: fwd (K1+K2), bwd (K3), both shapes
    for shape in ((32, 250, 65, 1000), (1, 743, 201, 1000)):
        B, T, U1, V = shape
        x = torch.randn(*shape, device="cuda")
        lab = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
        tl = torch.full((B,), T, dtype=torch.int32, device="cuda")
        ul = torch.full((B,), U1 - 1, dtype=torch.int32, device="cuda")
        ws_bytes = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        costs = torch.empty(B, device="cuda")
        grads = torch.empty_like(x)
        st = torch.cuda.current_stream().cuda_stream

        def fwd():
            _lib.check(lib.tfasr_rnnt_loss_fwd_f32(x.data_ptr(), lab.data_ptr(), tl.data_ptr(), ul.data_ptr(), B, T, U1, V, 0,
                                                   costs.data_ptr(), ws.data_ptr(), ws_bytes, st), "fwd")

        def bwd():
            _lib.check(lib.tfasr_rnnt_loss_bwd_f32(x.data_ptr(), lab.data_ptr(), tl.data_ptr(), ul.data_ptr(), B, T, U1, V, 0,
                                                   costs.data_ptr(), None, grads.data_ptr(), ws.data_ptr(), ws_bytes, st), "bwd")

        def both():
            fwd()
            bwd()

        n = B * T * U1 * V * 4
        mf, _ = time_ms(fwd)
        mb, _ = time_ms(bwd)
        mt, bt = time_ms(both)
        key = "rnnt_%dx%dx%dx%d" % shape
        res[key] = dict(fwd_ms=mf, bwd_ms=mb, total_ms=mt, total_ms_best=bt,
                        fwd_gbs=n / mf / 1e6, bwd_gbs=2 * n / mb / 1e6,
                        algorithmic_gbs=2 * n / mt / 1e6, algorithmic_frac=2 * n / mt / 1e6 / pk["hbm_gbs"],
                        practical_gbs=3 * n / mt / 1e6, practical_frac=3 * n / mt / 1e6 / pk["hbm_gbs"])
    print(json.dumps(res, indent=1))
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
