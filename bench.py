#!/usr/bin/env python
"""bench.py — utterances/sec of the Conformer-S RNN-T training step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port: TF is not installable), rank 0 only

One step = frontend (log-mel) -> Conformer-S encoder fwd -> prediction net -> joint -> RNN-T loss+grad -> backward ->
gradient all-reduce -> Adam, on `configs[2]` of BASELINE.json: batch 32 x 10 s per GPU (weak scaling), 80-mel, vocab 1000,
U = 64 labels, bf16 tensor-core operands / fp32 accumulation, dropout 0.1 on (training mode).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "utterances/sec (10 s, 80-mel) Conformer-S RNN-T training step"
UNIT = "utterances/s"
B_PER_GPU, N_SAMPLES, U_LABELS, VOCAB = 32, 160000, 64, 1000


def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        pk["_source"] = "measured"
        return pk
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


def synth_batch(batch, seed, device="cpu", pin=False):
    """SURVEY.md §8d synthetic inputs: 0.1*randn clipped audio, labels randint(1, V), predictions = [blank] + labels."""
    g = torch.Generator().manual_seed(seed)
    sig = (0.1 * torch.randn(batch, N_SAMPLES, generator=g)).clamp_(-1, 1)
    sig_len = torch.full((batch,), N_SAMPLES, dtype=torch.int32)
    labels = torch.randint(1, VOCAB, (batch, U_LABELS), generator=g, dtype=torch.int32)
    lab_len = torch.full((batch,), U_LABELS, dtype=torch.int32)
    pred = torch.cat([torch.zeros(batch, 1, dtype=torch.int32), labels], dim=1)
    pred_len = lab_len + 1
    ts = [sig, sig_len, pred, pred_len, labels, lab_len]
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device != "cpu":
        ts = [t.to(device) for t in ts]
    return ts


def to_train_data(ts):
    from tensorflowasr_b200 import schemas

    sig, sig_len, pred, pred_len, labels, lab_len = ts
    return schemas.TrainData(schemas.TrainInput(sig, sig_len, pred, pred_len), schemas.TrainLabel(labels, lab_len))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.15)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_reference_step_fn(batch):
    """The reference's CPU path, as an oracle port (TensorFlow/Keras are not installable in this image): the torch-CPU
    restatement in oracle/conformer.py of FeatureExtraction -> Conformer-S -> joint -> RNN-T loss, forward + backward + Adam,
    fp32, all host threads.  Returns (step_fn, cores)."""
    import numpy as np

    from oracle import conformer as oc
    from oracle import logmel

    cores = host_threads()
    cfg = oc.ConformerConfig(vocab_size=VOCAB)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "moving" not in k) for k, v in oc.init_params(cfg, seed=42).items()}
    leaves = [v for v in params.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-4, betas=(0.9, 0.98), eps=1e-9)
    sig, sig_len, pred, pred_len, labels, lab_len = synth_batch(batch, seed=1234)

    def step():
        feats, flen = logmel.feature_extraction(sig.numpy(), sig_len.numpy(), dtype=np.float32)
        logits, llen = oc.transducer_forward(params, cfg, torch.from_numpy(feats), torch.from_numpy(flen).long(), pred, pred_len.long(),
                                             training=True)
        losses = oc.rnnt_loss_torch(logits, labels, llen.int(), lab_len, fast=True)
        loss = losses.mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, cores


def host_threads():
    """CPU threads this process may really use: affinity mask and cgroup quota, not os.cpu_count() (a 128-thread OpenMP team on a
    box that grants fewer cores spends its time spinning: measured 8x slower than 8 threads on this workload)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def time_cpu(batch, steps, warmup):
    """-> (utterances/s, ms per step, threads used).  The thread count is the fastest of {8, 16, 32, all} that the box grants, chosen
    on one probe step each (the many small fp32 ops of this model do not scale past a few dozen threads)."""
    step, _ = cpu_reference_step_fn(batch)
    avail = host_threads()
    cands = sorted({min(c, avail) for c in (8, 16, 32, avail)})
    torch.set_num_threads(cands[0])
    for _ in range(max(1, warmup)):
        step()
    best, best_dt = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = c, dt
        if dt > 4 * best_dt:
            break  # more threads are only getting slower
    torch.set_num_threads(best)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, best


def time_cpu_subprocess(batch, steps, warmup, timeout_s=240):
    """The CPU arm in a fresh process (no CUDA context, own OpenMP pool) under a hard deadline; None on timeout."""
    import subprocess
    import sys

    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup), "--cpu-batch", str(batch)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env).stdout
    except subprocess.TimeoutExpired:
        return None
    for ln in reversed(out.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


def pick_cpu_batch(steps, warmup, budget_s=170.0):
    """The CPU arm's batch: the GPU arm's 32 utterances per step when the whole --steps/--warmup run fits `budget_s` on this host, else the
    largest of {16, 8, 4, 2, 1} that does (the arm must end within a few minutes).  Probe: one step at batch 2 after one warm-up step; the
    step time of this model grows about linearly with the batch (the lattice and the logits dominate)."""
    step, _ = cpu_reference_step_fn(2)
    torch.set_num_threads(min(host_threads(), 32))
    step()
    t0 = time.perf_counter()
    step()
    per_utt = (time.perf_counter() - t0) / 2
    n_steps = steps + max(1, warmup) + 4  # + the thread-count probe steps
    for b in (32, 16, 8, 4, 2, 1):
        if b * per_utt * n_steps <= budget_s:
            return b, per_utt
    return 1, per_utt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, how = args.cpu_batch, "as requested"
    if batch <= 0:
        batch, per_utt = pick_cpu_batch(args.steps, args.warmup)
        how = (f"largest batch of 32/16/8/4/2/1 whose whole run fits ~170 s on this host (probe: {per_utt:.2f} s per utterance and step); "
               "the GPU arm runs 32 per step")
    val, ms, cores = time_cpu(batch, args.steps, args.warmup)
    sample = (f"{batch} x 10 s utterance(s) per step ({how}; full Conformer-S train step, fp32 oracle port), {args.steps} steps after "
              f"{max(1, args.warmup)} warm-up + thread-count probe ({cores} threads chosen of {host_threads()} granted)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, batch_override=batch),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, batch_override=None):
    b = batch_override if batch_override is not None else B_PER_GPU * n_gpus
    return {"workload": "BASELINE.json configs[2]: Conformer-S Transducer (encoder-dim 144, 16 layers) RNN-T train step, batch=32x10 s per GPU",
            "global_batch": b, "audio_seconds": 10, "labels_per_utterance": U_LABELS, "vocab": VOCAB, "mel": 80,
            "parallelism": f"dp{n_gpus}", "dropout": 0.1, "step": "one CUDA graph replay per step (fwd+bwd+allreduce+Adam)",
            "l2_policy": ("no explicit flush: each step streams > 3 GB through a 126 MB L2 (the joint logits of 32 utterances alone are 1.04 GB as bf16 — "
                          "2.08 GB as f32 on the f32-logits route — written once, read and overwritten by their gradient, read twice more by the backward GEMMs)")}


# ---------------------------------------------------------------------------------------------------- B200 arm
def run_ours(args):
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner at communicator creation) are sent to
    # stderr for the duration of the run; the result line goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _run_ours(args, real_stdout)
    finally:
        real_stdout.flush()


def _run_ours(args, real_stdout):
    if args.config != "conformer":
        import bench_configs

        return bench_configs.run(args, real_stdout, {"ClockSampler": ClockSampler, "peaks": peaks, "host_threads": host_threads})
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import _lib, ops
    from tensorflowasr_b200 import dist as tdist

    rank, world, dev = tdist.init_from_env()
    assert dev.type == "cuda", "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(dev)
    lib = _lib.load()
    ops.set_precision("bf16")
    torch.manual_seed(1234 + rank)
    cfg = tfasr.conformer_small_config(vocab_size=VOCAB)
    model = tfasr.Conformer(**cfg, seed=42).to(dev)
    sched = tfasr.TransformerSchedule(dmodel=144, warmup_steps=10000, max_lr=0.05 / (144 ** 0.5), scale=2.0)
    model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=sched, beta_1=0.9, beta_2=0.98, epsilon=1e-9, weight_decay=1e-6))
    model.train()
    model.enable_cuda_graph(not args.no_graph)

    host = synth_batch(B_PER_GPU, seed=1234 + rank, pin=True)
    resident = to_train_data([t.to(dev) for t in host])

    def step_resident():
        return model.train_step(resident)["loss"]

    def step_e2e():
        data = to_train_data([t.to(dev, non_blocking=True) for t in host])  # H2D from pinned memory, every step
        losses = model.train_step(data)["loss"]
        return losses.mean().item()  # D2H read of the step's loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item() / steps

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    # per-kernel-group CUDA-event timings and launch counts: the same step run eagerly (event records cannot be read back from
    # a replayed graph); the kernels and their order are identical to the captured ones
    model.enable_cuda_graph(False)
    step_resident()
    lib.tfasr_launch_count_reset()
    ops.TIMERS.enable(True)
    for _ in range(3):
        torch.cuda._sleep(40_000_000)  # ~20 ms of device idle time: the host enqueues the whole eager step behind it, so the event pairs
        step_resident()                # around short kernels measure back-to-back device time, not host launch gaps
    kernel_ms = ops.TIMERS.report(3)
    ops.TIMERS.enable(False)
    launches = lib.tfasr_launch_count() // 3 * args.steps
    model.enable_cuda_graph(not args.no_graph)
    # end to end through the public training loop (BaseModel.fit, tensorflowasr_b200/training.py): every step copies its batch from pinned host
    # memory (on the loop's copy stream, one batch ahead) and reads its loss back to the host (one step late); fit returns after the last
    # loss arrived.  The unpipelined figure (copy -> step -> blocking read, one after the other) is kept beside it.
    def host_batches(n):
        for _ in range(n):
            yield to_train_data(host)

    e2e_io = {}

    def fit_e2e(n):
        h = model.fit(host_batches(n), epochs=1)
        assert len(h.steps["loss"]) == n and all(v == v for v in h.steps["loss"]), h.steps
        e2e_io.update(h2d=h.h2d_bytes // n, d2h=h.d2h_bytes // n)

    fit_e2e(3)
    ms_e2e = timed(lambda: fit_e2e(args.steps), 1) / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e_serial = timed(step_e2e, max(3, args.steps // 2))
    peer_calls = tdist.peer_all_reduce_check()  # opt-in peer-memory sync-BN exchange: raises if a peer ever missed the bounded wait

    if rank != 0:
        return
    pk = peaks()
    gb = B_PER_GPU * world
    h2d = sum(t.numel() * t.element_size() for t in host)
    # Rooflines.  `roofline`: the RNN-T loss+grad group (rnnt_lse_combine + rnnt_lattice + rnnt_grad_bf16_v2), the largest HBM-bound
    # piece of the step and the one the survey gives algorithmic bytes for (read the f32 logits once + write the gradient once =
    # 8 B per lattice node and vocabulary entry).  `roofline_joint_gemm`: the largest tcgen05 launch (joint logits: M = B T' (U+1) =
    # 520 000, K = 320, N = 1000); its 2.08 GB of f32 logits make it HBM-bound as well (t_hbm 367 us > t_tensor 234 us at the
    # measured peaks), so its fraction is reported against the HBM copy peak, with the tensor figure beside it.
    T_enc, U1, J = 250, U_LABELS + 1, 320
    rows = B_PER_GPU * T_enc * U1
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "roofline_traffic_r02.json")))
    except (OSError, ValueError):
        pass
    timing = "CUDA events on the launching stream around the launches, same step run eagerly right after the timed region"
    from tensorflowasr_b200 import ops as _ops

    # algorithmic bytes per lattice node and vocabulary entry: the survey's 8 B (f32 logits read once + gradient written once) on the
    # f32-logits route; 2 + 2 B with bf16 logits (ops.bf16_logits), where the tensors are half the size — the fraction is quoted on the
    # bytes the route has to move, the survey's count is kept beside it
    lb = 2 if _ops.bf16_logits_active() else 4
    alg_bytes = (lb + (2 if lb == 2 else 4)) * rows * VOCAB
    tkey = "rnnt_loss_grad_bf16_logits" if lb == 2 else "rnnt_loss_grad"
    rn_ms = kernel_ms.get("rnnt_fwd", 0.0) + kernel_ms.get("rnnt_bwd", 0.0)
    roofline = None
    if rn_ms > 0:
        ach = alg_bytes / (rn_ms * 1e-3) / 1e9
        gk_ms = kernel_ms.get("rnnt_bwd", 0.0)
        lat = "rnnt_lattice_smem_kernel" if _ops.lattice_v2_active() else "rnnt_lattice_kernel"
        roofline = {"kernel": (f"rnnt_lse_combine_kernel (log-softmax normaliser from the joint GEMM's row statistics) + {lat} + "
                               + ("rnnt_grad_bf16_io8_kernel (bf16 logits -> bf16 gradient in place)" if lb == 2 else "rnnt_grad_bf16_v2_kernel")
                               + " (RNN-T loss + gradient, one launch each per step)"),
                    "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": traffic.get(tkey, {}).get("bytes"), "traffic_source": traffic.get(tkey, {}).get("source"),
                    "peak_source": pk["_source"], "algorithmic_bytes_per_launch": alg_bytes, "logit_bytes": lb,
                    "survey_bytes_per_launch": 8 * rows * VOCAB,
                    "frac_on_survey_bytes": 8 * rows * VOCAB / (rn_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "ms_per_launch": rn_ms, "share_of_step": rn_ms / ms,
                    "gradient_kernel": {"ms": gk_ms, "frac": (alg_bytes / (gk_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if gk_ms > 0 else None,
                                        "note": "the byte-moving launch alone; the group also holds the latency-bound lattice wavefronts"},
                    "timing": timing}
    jg_ms = kernel_ms.get("joint_gemm_fwd", 0.0)
    roofline_joint = None
    if jg_ms > 0:
        jkey = "joint_gemm_bf16_logits" if lb == 2 else "joint_gemm"
        gemm_flops = 2.0 * rows * J * VOCAB
        gemm_bytes = 2.0 * rows * J + 2.0 * VOCAB * J + float(lb) * rows * VOCAB
        ach = gemm_bytes / (jg_ms * 1e-3) / 1e9
        tpeak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        roofline_joint = {"kernel": f"gemm_bf16_kernel<256> (joint logits = tanh(e+p) . Wv + bv, {'bf16' if lb == 2 else 'f32'} out through TMA stores)", "bound": "hbm",
                          "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                          "traffic": traffic.get(jkey, {}).get("bytes"), "traffic_source": traffic.get(jkey, {}).get("source"),
                          "algorithmic_bytes_per_launch": gemm_bytes, "algorithmic_flops_per_launch": gemm_flops,
                          "tensor_tflops": gemm_flops / (jg_ms * 1e-3) / 1e12, "tensor_frac_of_sustained_bf16": gemm_flops / (jg_ms * 1e-3) / 1e12 / tpeak,
                          "ms_per_launch": jg_ms, "share_of_step": jg_ms / ms, "peak_source": pk["_source"], "timing": timing}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch if args.cpu_batch > 0 else 4  # ~20 s of CPU work inside the default run; `--impl reference` sizes its own batch
        ref = time_cpu_subprocess(cb, 3, 1)
        sample = f"{cb} x 10 s utterance(s) per step (full train step, fp32 oracle port), 3 steps after 1 warm-up + thread-count probe"
        if ref is None:
            cpu = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample + " -- did not finish within 240 s"}
        else:
            cpu = {"value": ref["value"], "unit": UNIT, "cores": ref["cpu_baseline"]["cores"], "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": gb / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": gb / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": e2e_io.get("h2d", h2d), "d2h_bytes_per_step": e2e_io.get("d2h", 4),
                "ms_per_step": ms_e2e, "api": "BaseModel.fit over pinned host batches (copy stream one batch ahead, loss read one step late)",
                "unpipelined_ms_per_step": ms_e2e_serial},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_joint_gemm": roofline_joint, "cpu_baseline": cpu,
        "kernel_ms_per_step": kernel_ms,
    }
    if world > 1:
        line["config"]["sync_bn_exchange"] = "peer-memory kernel (csrc/peer_allreduce.cu)" if peer_calls else "nccl all-reduce per BatchNorm pass"
    print(json.dumps(line), file=real_stdout, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="utterances per CPU step (bounded sample of the workload); 0 = the largest of 32/16/8/4/2/1 that keeps the run within minutes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="conformer", choices=["conformer", "ds2", "contextnet", "decode"],
                    help="conformer = BASELINE.json configs[2], the headline (default); ds2 / contextnet / decode = configs[1] / [3] / [4] (bench_configs.py)")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
