"""Kernel-level breakdown of one Conformer-S train step.

    python tools/profile_step.py            # torch.profiler table (top kernels by device time)
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --ncu   # cudaProfilerStart/Stop around ONE step for an ncu launch list
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--batch", type=int, default=bench.B_PER_GPU)
    ap.add_argument("--rows", type=int, default=45)
    args = ap.parse_args()
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    dev = torch.device("cuda", 0)
    ops.set_precision("bf16")
    model = tfasr.Conformer(**tfasr.conformer_small_config(vocab_size=bench.VOCAB), seed=42).to(dev)
    model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-4))
    model.train()
    data = bench.to_train_data([t.to(dev) for t in bench.synth_batch(args.batch, seed=1234)])
    for _ in range(3):
        model.train_step(data)
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.cudart().cudaProfilerStart()
        model.train_step(data)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            model.train_step(data)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=90))
    # device-side kernels only, every one of them, per step
    kern = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = kern.setdefault(ev.name, [0.0, 0, 1e30, 0.0])
            k[0] += ev.device_time
            k[1] += 1
            k[2] = min(k[2], ev.device_time)
            k[3] = max(k[3], ev.device_time)
    total = sum(v[0] for v in kern.values())
    print(f"\n== device kernels per step: {total / 2e3:.3f} ms over {sum(v[1] for v in kern.values()) // 2} launches")
    for name, (t, c, lo, hi) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
        print(f"{t / 2e3:8.3f} ms {c // 2:5d} x {t / c:8.1f} us [{lo:7.1f} .. {hi:7.1f}]  {name[:130]}")


if __name__ == "__main__":
    main()
