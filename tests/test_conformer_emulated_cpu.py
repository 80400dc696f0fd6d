"""CPU: one Conformer-transducer training step — frontend, SpecAugment off, Conv2dSubsampling, one Conformer block (FFN, relative-position
MHSA, conv module, FFN, LN), prediction net (cluster LSTM), joint, RNN-T loss and the whole backward — through the PRODUCT's fused Python path
(fused.py, encoder.py, transducer.py, kernels.py unchanged) with every hand-written kernel except the tcgen05 GEMM running as its own source
on the CPU emulator (tests/emu_backend.py).  Checks: losses against the fp32 library route, every parameter gradient by cosine similarity
(bf16 noise changes magnitudes by percents, a wiring mistake changes directions), and that switching on every second-generation kernel
(conv1 / im2col / joint / dwconv weight gradient / RNN-T gradient v2, fused log-sum-exp) leaves losses and gradients where they were."""
import shutil

import pytest
import torch

import emu_backend

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available (emulator build)")


def _model(V=16, blocks=1):
    import tensorflowasr_b200 as tfasr

    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)
    cfg.update(encoder_num_blocks=blocks, encoder_dmodel=16, encoder_head_size=8, encoder_num_heads=2, encoder_dropout=0.0, prediction_embed_dim=16,
               prediction_rnn_units=64, joint_dim=16, encoder_kernel_size=7)
    cfg["encoder_subsampling"]["config"]["filters"] = [16, 16]
    model = tfasr.Conformer(**cfg, seed=3)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("bias", "beta")):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif n.endswith("gamma"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return model


def _batch(V=16, B=2, N=11200, U=4):
    from tensorflowasr_b200 import schemas

    g = torch.Generator().manual_seed(0)
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.tensor([N, N - 2300], dtype=torch.int32)
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    lab_len = torch.tensor([U, U - 2], dtype=torch.int32)
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)
    return schemas.TrainData(schemas.TrainInput(sig, sig_len, pred, lab_len + 1), schemas.TrainLabel(labels, lab_len))


def _run(model, data, state0):
    model.load_state_dict(state0)
    for p in model.parameters():
        p.grad = None
    model.train()
    losses = model.compute_losses(data, training=True)
    losses.mean().backward()
    return losses.detach().clone(), {n: p.grad.detach().clone().float() for n, p in model.named_parameters() if p.grad is not None}


def _cos(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def test_conformer_train_step_on_emulated_kernels(monkeypatch):
    from tensorflowasr_b200 import ops

    emu = emu_backend.install(monkeypatch)
    model, data = _model(blocks=2), _batch()
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # + the batched position projection: one GEMM for all blocks, attention kernels on column slices, one shared gradient buffer
    v2 = (ops.set_conv1_v2, ops.set_im2col_v2, ops.set_joint_v2, ops.set_joint_reduce_v2, ops.set_dwconv_wgrad_v2, ops.set_rnnt_grad_v2,
          ops.set_fused_lse, ops.set_pos_batched)
    saved = ops.switch_state()
    try:
        for s in v2:
            s(False)
        ops.set_bf16_logits(False)  # the head's round-2 defaults come back on for the fourth configuration below
        ops.set_lattice_v2(False)
        ops.set_precision("fp32")
        ops.set_fused(False)
        ref_loss, ref_g = _run(model, data, state0)  # library route, fp32: the reference
        ops.set_precision("bf16")
        ops.set_fused(True)
        emu.calls.clear()
        loss1, g1 = _run(model, data, state0)  # fused route, first-generation kernels (what bench.py runs)
        calls1 = dict(emu.calls)
        for s in v2:
            s(True)
        emu.calls.clear()
        loss2, g2 = _run(model, data, state0)  # fused route, every second-generation kernel + fused log-sum-exp
        calls2 = dict(emu.calls)
        loss3, g3 = _run(model, data, state0)  # once more on the same packs: packed gradient sums and the shared dpos buffer start clean
        ops.set_bf16_logits(True)
        ops.set_lattice_v2(True)
        emu.calls.clear()
        loss4, g4 = _run(model, data, state0)  # + bf16 logits (combine / gradient kernels on the 2-byte tensor, gradient in place at pitch V) + shared-memory wavefronts
        calls4 = dict(emu.calls)
    finally:
        ops.restore_switches(saved)
        ops.set_precision("fp32")
    second_gen = ("tfasr_conv1_3x3s2_fwd_v2", "tfasr_conv1_3x3s2_wgrad_v2", "tfasr_im2col_3x3s2_v2", "tfasr_col2im_3x3s2_v2", "tfasr_joint_tanh_fwd_v2",
                  "tfasr_joint_reduce_bwd_v2", "tfasr_dwconv_causal_wgrad_v2", "tfasr_dwconv_causal_bwd_dx", "tfasr_rnnt_loss_bwd_bf16_v2",
                  "tfasr_rnnt_loss_fwd_rowstats_f32", "tfasr_rel_attention_fwd_ld", "tfasr_rel_attention_bwd_ld")
    assert not any(k in calls1 for k in second_gen), calls1  # every switch off: the first-generation kernels ...
    assert all(calls2.get(k, 0) > 0 for k in second_gen), calls2  # ... and every switch routes to its second-generation kernel
    for k in ("tfasr_logmel_f32", "tfasr_layernorm_fwd", "tfasr_layernorm_bwd", "tfasr_rel_attention_fwd", "tfasr_rel_attention_bwd", "tfasr_lstm_rec_fwd",
              "tfasr_lstm_rec_bwd", "tfasr_bn_stats", "tfasr_bn_bwd_apply", "tfasr_dwconv_causal_fwd", "tfasr_glu_bwd", "tfasr_pack_weights",
              "tfasr_rnnt_loss_fwd_f32", "tfasr_rnnt_loss_bwd_bf16", "tfasr_conv1_3x3s2_fwd", "tfasr_im2col_3x3s2", "tfasr_joint_tanh_fwd"):
        assert calls1.get(k, 0) > 0, (k, sorted(calls1))  # the step really ran on the hand-written kernels
    assert calls2["tfasr_rel_attention_fwd_ld"] == 2 and "tfasr_rel_attention_fwd" not in calls2 and "tfasr_rel_attention_bwd" not in calls2
    assert calls4.get("tfasr_rnnt_loss_bwd_bf16_io", 0) > 0, calls4
    assert calls4.get("tfasr_rnnt_lse_from_rowstats_bf16", 0) > 0 and calls4.get("tfasr_rnnt_lattice_smem_f32", 0) > 0, calls4
    assert "tfasr_rnnt_loss_bwd_bf16_v2" not in calls4 and "tfasr_rnnt_loss_fwd_rowstats_f32" not in calls4
    torch.testing.assert_close(loss4, loss2, rtol=5e-3, atol=5e-3)
    assert torch.isfinite(loss1).all() and torch.isfinite(loss2).all()
    torch.testing.assert_close(loss3, loss2, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(loss1, ref_loss, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(loss2, loss1, rtol=5e-3, atol=5e-3)
    checked = 0
    for n, g in ref_g.items():
        if n not in g1:  # conv biases in front of a batch norm: identically zero, not computed by the fused route
            assert n.endswith("bias"), n
            continue
        if g.norm() < 1e-5 * max(1.0, g.numel() ** 0.5) or any(k in n for k in ("key__bias", "encoding__bias", "dw_conv__bias")):
            continue  # analytically zero gradients are rounding noise
        assert _cos(g1[n], g) > 0.95, (n, _cos(g1[n], g))
        assert _cos(g2[n], g1[n]) > 0.995, (n, _cos(g2[n], g1[n]))
        assert _cos(g3[n], g2[n]) > 0.995 and abs(float(g3[n].norm() / g2[n].norm()) - 1.0) < 0.05, n
        assert _cos(g4[n], g2[n]) > 0.995 and abs(float(g4[n].norm() / g2[n].norm()) - 1.0) < 0.05, (n, _cos(g4[n], g2[n]))
        checked += 1
    assert checked > 40, checked


def test_train_step_with_flat_adamw_learns_on_emulated_kernels(monkeypatch):
    """BaseModel.fit -> train_step end to end (forward, backward, the one-launch flat AdamW, gradient zeroing) on the emulator backend: the loss
    of a fixed batch goes down step after step; the loop reports every step's loss to the callbacks and the History (tensorflow_asr/scripts/train.py:113-121)."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    emu = emu_backend.install(monkeypatch)
    try:
        ops.set_precision("bf16")
        model, data = _model(), _batch()
        model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=2e-3))
        model.train()
        seen = []

        class Log:
            def on_train_batch_end(self, step, logs):
                seen.append((step, logs["loss"]))

            def on_epoch_end(self, epoch, logs):
                seen.append(("epoch", epoch, logs["loss"]))

        hist = model.fit([data, data], epochs=1, callbacks=[Log()])
        losses = hist.steps["loss"]
    finally:
        ops.set_precision("fp32")
    assert len(losses) == 2 and [s[0] for s in seen] == [0, 1, "epoch"] and seen[0][1] == losses[0]
    assert abs(hist.history["loss"][0] - sum(losses) / 2) < 1e-6 and hist.epoch == [0]
    assert emu.calls.get("tfasr_adamw_flat_f32", 0) == 2 and emu.calls.get("tfasr_gemm_bf16", 0) > 50
    assert losses[1] < losses[0], losses


def _grads_close(ref_g, g, skip, min_checked, cos_min=0.95):
    checked = 0
    for n, r in ref_g.items():
        if any(k in n for k in skip) or n not in g or r.norm() < 1e-5 * max(1.0, r.numel() ** 0.5):
            continue
        assert _cos(g[n], r) > cos_min, (n, _cos(g[n], r))
        checked += 1
    assert checked >= min_checked, checked


def test_contextnet_transducer_fused_route_on_emulated_kernels(monkeypatch):
    """ContextNet transducer: log-mel -> fused_seq encoder route (depthwise / squeeze-excite kernels, BatchNorm kernels, GEMM
    contract) -> prediction net -> joint -> RNN-T loss, forward + backward, against the fp32 library route."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    emu = emu_backend.install(monkeypatch)
    blocks = [dict(nlayers=1, kernel_size=5, filters=32, strides=1, residual=False, activation="silu", padding="causal"),
              dict(nlayers=2, kernel_size=5, filters=32, strides=2, residual=True, activation="silu", padding="causal"),
              dict(nlayers=1, kernel_size=3, filters=48, strides=2, residual=True, activation="silu", padding="causal")]
    V = 16
    cfg = tfasr.contextnet_config(vocab_size=V, alpha=0.5)
    cfg.update(encoder_blocks=blocks, prediction_embed_dim=16, prediction_rnn_units=64, joint_dim=16)
    model = tfasr.ContextNet(**cfg, seed=4)
    data = _batch(V=V)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    try:
        ops.set_precision("fp32")
        ops.set_fused(False)
        ref_loss, ref_g = _run(model, data, state0)
        ops.set_precision("bf16")
        ops.set_fused(True)
        ops.set_fused_seq(True)
        emu.calls.clear()
        loss, g = _run(model, data, state0)
    finally:
        ops.set_fused_seq(False)
        ops.set_fused(True)
        ops.set_precision("fp32")
    for k in ("tfasr_seq_dwconv_fwd", "tfasr_seq_dwconv_bwd", "tfasr_seq_masked_mean", "tfasr_se_fc_fwd", "tfasr_se_fc_bwd", "tfasr_se_apply_bwd", "tfasr_se_dy",
              "tfasr_bn_stats", "tfasr_bn_bwd_apply"):
        assert emu.calls.get(k, 0) > 0, (k, sorted(emu.calls))
    torch.testing.assert_close(loss, ref_loss, rtol=4e-2, atol=4e-2)
    _grads_close(ref_g, g, skip=("conv__bias",), min_checked=30)


def test_deepspeech2_ctc_fused_route_on_emulated_kernels(monkeypatch):
    """DeepSpeech2: spectrogram frontend -> fused_seq conv module (general im2col / col2im, BatchNorm + relu kernels) ->
    BiLSTM on the cluster LSTM kernels -> fc (bias + relu epilogue contract) -> vocab Dense -> CTC loss kernel, forward + backward."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops, schemas

    emu = emu_backend.install(monkeypatch)
    cfg = tfasr.deepspeech2_base_config(vocab_size=12)
    cfg.update(conv_kernels=[[5, 7], [3, 5]], conv_strides=[[2, 2], [1, 2]], conv_filters=[8, 8], rnn_nlayers=1, rnn_units=64, fc_units=24, rnn_dropout=0.0,
               fc_dropout=0.0)
    cfg["speech_config"].update(num_feature_bins=24)
    model = tfasr.DeepSpeech2(**cfg, seed=2)
    model.compile(optimizer=torch.optim.SGD(model.parameters(), lr=0.0))
    g0 = torch.Generator().manual_seed(1)
    B, N, L = 2, 6400, 4
    sig = (0.1 * torch.randn(B, N, generator=g0)).clamp(-1, 1)
    data = schemas.TrainData(schemas.TrainInput(sig, torch.tensor([N, N - 1500], dtype=torch.int32), None, None),
                             schemas.TrainLabel(torch.randint(1, 12, (B, L), generator=g0, dtype=torch.int32), torch.tensor([L, 2], dtype=torch.int32)))
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    try:
        ops.set_precision("fp32")
        ref_loss, ref_g = _run(model, data, state0)
        ops.set_precision("bf16")
        ops.set_fused_seq(True)
        emu.calls.clear()
        loss, g = _run(model, data, state0)
    finally:
        ops.set_fused_seq(False)
        ops.set_precision("fp32")
    for k in ("tfasr_conv_im2col", "tfasr_conv_col2im", "tfasr_bn_apply", "tfasr_lstm_rec_fwd", "tfasr_lstm_rec_bwd", "tfasr_ctc_loss_f32"):
        assert emu.calls.get(k, 0) > 0, (k, sorted(emu.calls))
    torch.testing.assert_close(loss, ref_loss, rtol=4e-2, atol=4e-2)
    _grads_close(ref_g, g, skip=("conv2d__bias",), min_checked=12, cos_min=0.93)


def test_feature_extraction_api_reproduces_testflac_goldens(monkeypatch):
    """tensorflowasr_b200.FeatureExtraction (the product's Python class + the log-mel kernel source on the emulator) on the reference's
    tests/test.flac: default layer arguments, the 16 000-sample `padding` variant of tests/test_layers.py:37, and the 160-bin spectrogram of
    the DeepSpeech2 configs — against tests/golden/logmel_testflac.npz."""
    import os

    import numpy as np

    import tensorflowasr_b200 as tfasr

    emu_backend.install(monkeypatch)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "logmel_testflac.npz"))
    sig = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None, :]
    n = torch.tensor([sig.shape[1]], dtype=torch.int32)
    for kwargs, feat_key, len_key in ((dict(), "features", "features_length"), (dict(padding=16000), "features_padded", "features_padded_length"),
                                      (dict(feature_type="spectrogram", num_feature_bins=160), "spectrogram160", "features_length")):
        feats, flen = tfasr.FeatureExtraction(**kwargs)((sig, n), training=False)
        want = torch.from_numpy(g[feat_key])
        assert feats.shape == want.shape and flen.tolist() == g[len_key].tolist()
        err = (feats - want).abs().max().item()
        assert err < 2e-3, (feat_key, err)  # north_star: log-mel within 1e-3 relative (values span 1 .. 14 in magnitude)


def test_greedy_decode_tokens_bit_exact_vs_oracle_on_emulator(monkeypatch):
    """The GPU suite's greedy-token gate (north_star: bit-exact token indices) on the CPU: log-mel kernel (emulated) -> Conformer encoder (fp32
    route) -> both reference decode loops, host-driven AND through the persistent decode kernel (its own source on the emulator), against the
    float64 oracle's restatement of base_transducer.py:496-712."""
    import numpy as np

    import tensorflowasr_b200 as tfasr
    from oracle import conformer as oc
    from oracle import logmel as ologmel
    from tensorflowasr_b200 import ops, schemas

    emu_backend.install(monkeypatch)
    ops.set_precision("fp32")
    V = 32
    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)
    cfg.update(encoder_num_blocks=2, encoder_dropout=0.0)
    model = tfasr.Conformer(**cfg, seed=1)
    ocfg = oc.ConformerConfig(vocab_size=V, num_blocks=2)
    params = oc.init_params(ocfg, seed=5, randomize_all=True)
    params["joint.vocab.bias"][0] += -0.6  # make random weights emit labels, not only blanks (as tests/test_gpu_model.py does)
    model.load_named_weights(params)
    model.eval()
    g = torch.Generator().manual_seed(0)
    B, N = 3, 9600
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.tensor([N, N - 2500, N // 2], dtype=torch.int32)
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.double() for k, v in params.items()}
    enc, elen = oc.conformer_encoder(torch.from_numpy(feats), torch.from_numpy(flen).long(), p, ocfg, training=False)
    ref, margin = oc.greedy_batch(enc, elen, p, ocfg)
    ref1, margin1 = oc.greedy_single(enc[:1], int(elen[0]), p, ocfg)
    assert margin > 1e-3 and margin1 > 1e-3, "pick other weights: the oracle's own arg-max is a near-tie"
    assert int((ref[:, 2:] != 0).sum()) >= 3
    try:
        for flag in (False, True):
            ops.set_persistent_decode(flag)
            out = model.recognize(schemas.PredictInput(sig, sig_len))
            assert out.tokens.shape == ref.shape and out.tokens.long().tolist() == ref.tolist(), ("persistent" if flag else "host loop")
            out1 = model.recognize(schemas.PredictInput(sig[:1], sig_len[:1]))
            assert out1.tokens[0].tolist() == ref1, ("persistent" if flag else "host loop", "single")
    finally:
        ops.set_persistent_decode(False)


def test_smoke_parity_check_on_emulator(monkeypatch):
    """What __graft_entry__.smoke() asserts on cuda:0, on the CPU: log-mel kernel -> Conformer (fp32 route) -> joint -> RNN-T loss kernels against
    the float64 oracle, 1e-3 relative."""
    import numpy as np

    import tensorflowasr_b200 as tfasr
    from oracle import conformer as oc
    from oracle import logmel as ologmel
    from tensorflowasr_b200 import ops, schemas

    emu = emu_backend.install(monkeypatch)
    ops.set_precision("fp32")
    torch.manual_seed(0)
    V, B, N, U = 64, 2, 16000, 5
    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)
    cfg.update(encoder_num_blocks=2, encoder_dropout=0.0)
    model = tfasr.Conformer(**cfg, seed=7)
    ocfg = oc.ConformerConfig(vocab_size=V, num_blocks=2)
    params = oc.init_params(ocfg, seed=11, randomize_all=True)
    model.load_named_weights(params)
    sig = (0.1 * torch.randn(B, N)).clamp(-1, 1)
    sig_len = torch.tensor([N, N - 3000], dtype=torch.int32)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    lab_len = torch.tensor([U, U - 2], dtype=torch.int32)
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)
    data = schemas.TrainData(schemas.TrainInput(sig, sig_len, pred, lab_len + 1), schemas.TrainLabel(labels, lab_len))
    losses = model.compute_losses(data, training=False)
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p64 = {k: v.double() for k, v in params.items()}
    logits, llen = oc.transducer_forward(p64, ocfg, torch.from_numpy(feats), torch.from_numpy(flen).long(), pred, (lab_len + 1).long())
    ref = oc.rnnt_loss_torch(logits, labels, llen.int(), lab_len)
    rel = ((losses.detach().double() - ref).abs() / ref.abs()).max().item()
    assert rel < 1e-3, (losses.tolist(), ref.tolist())
    assert emu.calls.get("tfasr_logmel_f32", 0) == 1 and emu.calls.get("tfasr_rnnt_loss_fwd_f32", 0) + emu.calls.get("tfasr_rnnt_loss_f32", 0) >= 1, emu.calls


def test_decode_bench_tool_dry_run_on_emulator(monkeypatch):
    """tools/decode_bench.py (round-2 measurement of the persistent decode kernel) end to end on the emulator backend: the tool's own
    check — host loop and persistent kernel emit the same tokens — must hold, and the script must not die on a Python error on the GPU box."""
    import os
    import sys

    emu_backend.install(monkeypatch)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import decode_bench

    res = decode_bench.main(["--batch", "2", "--frames", "5", "--device", "cpu"])
    assert res["tokens_equal"] and res["iterations"] >= 4 and res["persistent"]["ms"] > 0
