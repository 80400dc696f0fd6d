"""Greedy transducer decode on one B200 (Conformer-S sizes): the host-driven loop against the persistent kernel (csrc/decode.cu), on a
random-weight prediction / joint network and synthetic encoder outputs.  Reports ms per decode, iterations, us per iteration and the
L2-side bytes the persistent kernel streams per iteration.

    python tools/decode_bench.py [--batch 32] [--frames 250] [--threads 1024] [--out gpurun_out/decode_bench.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=250)
    ap.add_argument("--threads", type=int, default=1024)
    ap.add_argument("--blank-bias", type=float, default=4.0, help="added to the blank logit so that random weights emit a realistic token rate")
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", default="cuda", help="'cpu' = dry run on the emulator backend (tests/emu_backend.py must be installed by the caller)")
    args = ap.parse_args(argv)
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    ops.set_precision("fp32")
    dev = args.device
    sync = torch.cuda.synchronize if dev.startswith("cuda") else (lambda: None)
    model = tfasr.Conformer(**tfasr.conformer_small_config(vocab_size=1000), seed=1).to(dev).eval()
    with torch.no_grad():
        model.joint_net["vocab.bias"][0] += args.blank_bias
    B, T = args.batch, args.frames
    g = torch.Generator().manual_seed(0)
    enc = torch.randn(B, T, 144, generator=g).to(dev)
    enc_len = torch.full((B,), T, dtype=torch.int64, device=dev)
    prev = model.get_initial_tokens(B)
    states = model.get_initial_decoder_states(B)
    calls = [0]
    orig = model.call_next

    def counted(*a):
        calls[0] += 1
        if calls[0] > 8 * T:
            raise RuntimeError("host loop does not terminate with these random weights: raise --blank-bias")
        return orig(*a)

    model.call_next = counted
    model._encode_for_decode = lambda inputs: (enc, enc_len, [], prev, states)
    dummy = tfasr.schemas.PredictInput(torch.zeros(B, 1, device=dev), torch.ones(B, dtype=torch.int32, device=dev))
    res = {"gpu": torch.cuda.get_device_name(0) if dev.startswith("cuda") else "cpu emulator (dry run)", "batch": B, "frames": T}
    outs = {}
    for flag in (False, True):
        ops.set_persistent_decode(flag)
        times = []
        for it in range(3):
            calls[0] = 0
            sync()
            t0 = time.perf_counter()
            out = model.recognize_batch(dummy)
            sync()
            times.append((time.perf_counter() - t0) * 1e3)
        outs[flag] = out.tokens.cpu()
        key = "persistent" if flag else "host_loop"
        res[key] = {"ms": sorted(times)[1], "ms_all": times}
        if not flag:
            res["iterations"] = calls[0]
            res[key]["us_per_iteration"] = sorted(times)[1] * 1e3 / max(calls[0], 1)
    ops.set_persistent_decode(False)
    res["tokens_equal"] = bool((outs[False] == outs[True]).all())
    res["emitted_per_utterance"] = float((outs[False][:, 2:] != 0).sum(1).float().mean())
    res["persistent"]["us_per_iteration"] = res["persistent"]["ms"] * 1e3 / max(res["iterations"], 1)
    E = R = J = 320
    res["bytes_per_iteration_per_cta"] = 4 * ((E + R) * 4 * R + R * J + J * 1000)
    res["persistent"]["l2_gbs_all_ctas"] = res["bytes_per_iteration_per_cta"] * B / (res["persistent"]["us_per_iteration"] * 1e-6) / 1e9
    print(json.dumps(res, indent=1))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
    return res


if __name__ == "__main__":
    main()
