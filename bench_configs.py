"""bench.py --config {ds2, contextnet, decode}: the other BASELINE.json configurations through the same contract as the headline bench
(one JSON line; device-resident `value`, `e2e` with host buffers and the copies inside the timed region, clocks, a roofline object, the CPU
oracle port timed beside it on a bounded sample).  The headline stays `bench.py` without --config (configs[2]).

  ds2         configs[1]  DeepSpeech2 CTC (deepspeech2/base.yml.j2: 2 x Conv2D + BN + ReLU, 5 x BiLSTM(512), FC 1024, V = 29), batch 8 x 10 s,
                          training step, 1 GPU
  contextnet  configs[3]  ContextNet transducer, alpha = 1.0 ("M"), bf16, batch 64 x 15 s per GPU, training step, NCCL gradient all-reduce
  decode      configs[4]  Conformer-L (17 x 512, 8 heads x 64, kernel 32; the paper's L — the reference ships only S) streaming greedy decode:
                          chunk 16 / history 64 attention mask over the utterance, memory_length null, exactly the reference's streaming config
                          (small-streaming.yml.j2:37-39); 1000 x 30 s in batches; beam search for transducers does not exist in the reference
                          (base_transducer.py:841-842), so greedy is what is measured
"""
from __future__ import annotations

import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def _synth_audio(batch, seconds, seed):
    g = torch.Generator().manual_seed(seed)
    n = 16000 * seconds
    sig = (0.1 * torch.randn(batch, n, generator=g)).clamp_(-1, 1)
    return sig, torch.full((batch,), n, dtype=torch.int32), g


def conformer_large_config(vocab_size=1000):
    import tensorflowasr_b200 as tfasr

    cfg = tfasr.conformer_small_config(vocab_size=vocab_size)
    cfg["speech_config"].pop("augmentation_config", None)
    sub = cfg["encoder_subsampling"]["config"]
    sub.update(filters=[512, 512], norms=["layer", "layer"])  # small-streaming.yml.j2:19 (layer norms in the streaming variant)
    cfg.update(encoder_dmodel=512, encoder_num_blocks=17, encoder_head_size=64, encoder_num_heads=8, encoder_kernel_size=32,
               encoder_convm_dw_norm_type="layer", encoder_chunk_size=16, encoder_history_size=64, encoder_memory_length=None,
               prediction_embed_dim=640, prediction_rnn_units=640, joint_dim=640)
    return cfg


def _build(name, dev, world):
    """-> (model, make_host_batch(seed) -> list of host tensors, to_data(list) -> TrainData | PredictInput, meta)"""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import schemas

    if name == "ds2":
        V, B, S, L = 29, 8, 10, 100
        model = tfasr.DeepSpeech2(**tfasr.deepspeech2_base_config(vocab_size=V), seed=42).to(dev)

        def host(seed):
            sig, sl, g = _synth_audio(B, S, seed)
            labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32)
            return [sig, sl, labels, torch.full((B,), L, dtype=torch.int32)]

        def data(ts):
            return schemas.TrainData(schemas.TrainInput(ts[0], ts[1], None, None), schemas.TrainLabel(ts[2], ts[3]))

        meta = dict(metric="utterances/sec (10 s, 160-bin spectrogram) DeepSpeech2 CTC training step", batch=B, seconds=S, train=True,
                    workload="BASELINE.json configs[1]: DeepSpeech2 CTC (base.yml.j2: 2 conv + 5 x BiLSTM(512) + FC 1024), synthetic 16 kHz batch=8 x 10 s",
                    extra={"labels_per_utterance": L, "vocab": V})
    elif name == "contextnet":
        V, B, S, U = 1000, 64, 15, 64
        cfg = tfasr.contextnet_config(vocab_size=V, alpha=1.0)
        model = tfasr.ContextNet(**cfg, seed=42).to(dev)

        def host(seed):
            sig, sl, g = _synth_audio(B, S, seed)
            labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
            ll = torch.full((B,), U, dtype=torch.int32)
            pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)
            return [sig, sl, pred, ll + 1, labels, ll]

        def data(ts):
            return schemas.TrainData(schemas.TrainInput(ts[0], ts[1], ts[2], ts[3]), schemas.TrainLabel(ts[4], ts[5]))

        meta = dict(metric="utterances/sec (15 s, 80-mel) ContextNet-M RNN-T training step", batch=B, seconds=S, train=True,
                    workload="BASELINE.json configs[3]: ContextNet-M (alpha 1.0) Transducer bf16, batch=64 x 15 s per GPU, NCCL grad all-reduce",
                    extra={"labels_per_utterance": U, "vocab": V})
    elif name == "decode":
        V, B, S = 1000, 50, 30
        model = tfasr.Conformer(**conformer_large_config(V), seed=42).to(dev)
        with torch.no_grad():
            model.joint_net["vocab.bias"][0] += 3.0  # random weights: a realistic blank rate (else every frame emits its token cap)

        def host(seed):
            sig, sl, _ = _synth_audio(B, S, seed)
            return [sig, sl]

        def data(ts):
            return schemas.PredictInput(ts[0], ts[1])

        meta = dict(metric="utterances/sec (30 s, 80-mel) Conformer-L streaming (chunk 16 / history 64) greedy decode", batch=B, seconds=S, train=False,
                    workload="BASELINE.json configs[4]: Conformer-L streaming chunked greedy decode, 1000 x 30 s synthetic in batches of 50",
                    extra={"utterances": 1000, "vocab": V, "beam": "none in the reference for transducers (base_transducer.py:841-842)"})
    else:
        raise ValueError(name)
    return model, host, data, meta


def _matmul_flops_per_utt(model, meta):
    """Dense-contraction flops of ONE forward pass per utterance, from the weight shapes: 2 * in * out per frame the layer sees.  A lower
    bound of the step's work (elementwise / recurrence gate math not counted); a training step is 3x."""
    frames = meta["seconds"] * 100
    total = 0.0
    for n, p in model.named_parameters():
        if p.dim() < 2:
            continue
        nn = n.replace("__", ".")
        t = frames
        if "prediction" in nn or "predict_net" in nn:
            t = meta["extra"].get("labels_per_utterance", 64) + 1
        elif "joint" in nn:
            continue  # counted below (needs T' x U)
        elif "encoder" in nn:
            red = getattr(model, "time_reduction_factor", 4)
            t = frames // red if "subsampling" not in nn and "conv_0" not in nn else frames // 2
        total += 2.0 * p.numel() * t
    return total


def run(args, real_stdout, shared):
    """`shared`: helpers of bench.py (ClockSampler, peaks, host_threads)."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import _lib, ops
    from tensorflowasr_b200 import dist as tdist

    rank, world, dev = tdist.init_from_env()
    assert dev.type == "cuda", "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(dev)
    lib = _lib.load()
    ops.set_precision("bf16")
    ops.set_fused_seq(True)
    torch.manual_seed(1234 + rank)
    model, host_fn, to_data, meta = _build(args.config, dev, world)
    B = meta["batch"]
    host = [t.pin_memory() for t in host_fn(1234 + rank)]
    resident = to_data([t.to(dev) for t in host])
    if meta["train"]:
        model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-4, beta_1=0.9, beta_2=0.98, epsilon=1e-9, weight_decay=1e-6))
        model.train()
        if args.config == "ds2":
            args.no_graph = True  # the 512-unit BiLSTM runs on cuDNN's packed-sequence path, whose length bookkeeping syncs with the host
        model.enable_cuda_graph(not args.no_graph)

        def step_resident():
            return model.train_step(resident)["loss"]

        def step_e2e():
            return model.train_step(to_data([t.to(dev, non_blocking=True) for t in host]))["loss"].mean().item()

        d2h = 4
    else:
        model.eval()

        def step_resident():
            return model.recognize(resident).tokens

        def step_e2e():
            return model.recognize(to_data([t.to(dev, non_blocking=True) for t in host])).tokens.cpu()

        d2h = None

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
        out = step_resident()
    sampler = shared["ClockSampler"](dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    lib.tfasr_launch_count_reset()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = lib.tfasr_launch_count()
    if meta["train"] and not args.no_graph:  # replays do not pass through the launchers: count one eager step
        model.enable_cuda_graph(False)
        step_resident()
        lib.tfasr_launch_count_reset()
        step_resident()
        launches = lib.tfasr_launch_count() * args.steps
        model.enable_cuda_graph(True)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if d2h is None:
        d2h = int(out.numel() * out.element_size())
    if rank != 0:
        return
    pk = shared["peaks"]()
    gb = B * world
    h2d = sum(t.numel() * t.element_size() for t in host)
    fwd = _matmul_flops_per_utt(model, meta)
    flops_step = fwd * B * (3.0 if meta["train"] else 1.0)
    tpeak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach = flops_step / (ms * 1e-3) / 1e12
    roofline = {"kernel": "whole step (dense contractions from the weight shapes; joint / elementwise / recurrent gate math not counted)", "bound": "tensor",
                "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": None, "peak_source": pk["_source"],
                "algorithmic_flops_per_step": flops_step, "timing": "CUDA events around the timed steps"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port(args.config, meta)
    cfg = {"workload": meta["workload"], "global_batch": gb, "audio_seconds": meta["seconds"], "parallelism": f"dp{world}",
           "step": ("one CUDA graph replay per step (fwd+bwd+allreduce+Adam)" if meta["train"] and not args.no_graph else "eager launches"),
           "l2_policy": "inputs + activations per step exceed the 126 MB L2 (audio alone: %.0f MB)" % (h2d / 1e6)}
    cfg.update(meta["extra"])
    line = {"metric": meta["metric"], "value": gb / ms * 1e3, "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": gb / ms_e2e * 1e3, "unit": "utterances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    if args.config == "decode":
        line["utterances_1000_seconds"] = 1000.0 / (gb / ms_e2e * 1e3)
    print(json.dumps(line), file=real_stdout, flush=True)


def cpu_port(name, meta, budget_s=25.0):
    """The oracle port of the same workload on the host cores, on a bounded sample (a few utterances, forward [+ backward]); None for
    workloads the oracle does not restate end to end."""
    import numpy as np

    from oracle import logmel

    threads = min(len(os.sched_getaffinity(0)), 32)
    torch.set_num_threads(threads)
    try:
        if name == "decode":
            from oracle import conformer as oc

            cfg = oc.ConformerConfig(vocab_size=1000, dmodel=512, num_blocks=17, head_size=64, num_heads=8, kernel_size=32, dw_norm_type="layer",
                                     subsampling_norms=("layer", "layer"), chunk_size=16, history_size=64, embed_dim=640, rnn_units=640, joint_dim=640)
            params = oc.init_params(cfg, seed=42)
            params["joint.vocab.bias"][0] += 3.0
            sig, sl, _ = _synth_audio(1, meta["seconds"], 7)
            t0 = time.perf_counter()
            feats, flen = logmel.feature_extraction(sig.numpy(), sl.numpy(), dtype=np.float32)
            with torch.no_grad():
                enc, elen = oc.conformer_encoder(torch.from_numpy(feats), torch.from_numpy(flen).long(), params, cfg, training=False)
                oc.greedy_single(enc, int(elen[0]), params, cfg)
            dt = time.perf_counter() - t0
            return {"value": 1.0 / dt, "unit": "utterances/s", "cores": threads, "kind": "port",
                    "sample": "1 x 30 s utterance: log-mel + Conformer-L streaming encoder + greedy walk (oracle port, fp32), once"}
    except Exception as e:  # the oracle's config surface may not cover the variant: report that instead of failing the bench
        return {"value": None, "unit": "utterances/s", "cores": threads, "kind": "port", "sample": f"oracle port unavailable for this config: {e!r}"[:300]}
    return {"value": None, "unit": "utterances/s", "cores": threads, "kind": "port",
            "sample": "no end-to-end oracle port of this model's training step (oracle/other_models.py restates the forward only); see the headline bench"}
