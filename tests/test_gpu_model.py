"""GPU parity: Conformer transducer (encoder / prediction / joint / loss head / train step) vs the CPU oracle.
fp32 mode: outputs within 1e-3 relative of the float64 oracle; greedy tokens bit-exact vs the oracle's greedy walk."""
import numpy as np
import pytest
import torch

from oracle import conformer as oc
from oracle import logmel as ologmel

pytestmark = pytest.mark.gpu


def _setup(num_blocks=2, V=48, dropout=0.0, dw_norm="batch", seed=3, sub_norms=("batch", "batch"), chunk=None, hist=None, blank_bias=0.0):
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    ops.set_precision("fp32")
    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)  # SpecAugment is RNG-dependent: parity runs with it off
    cfg.update(encoder_num_blocks=num_blocks, encoder_dropout=dropout, encoder_convm_dw_norm_type=dw_norm, encoder_chunk_size=chunk,
               encoder_history_size=hist)
    cfg["encoder_subsampling"]["config"]["norms"] = list(sub_norms)
    model = tfasr.Conformer(**cfg, seed=1).cuda()
    ocfg = oc.ConformerConfig(vocab_size=V, num_blocks=num_blocks, dw_norm_type=dw_norm, subsampling_norms=tuple(sub_norms), chunk_size=chunk,
                              history_size=hist)
    params = oc.init_params(ocfg, seed=seed, randomize_all=True)
    params["joint.vocab.bias"][0] += blank_bias  # decoding tests: make random weights emit labels, not only blanks
    model.load_named_weights(params)
    return model, ocfg, params


def _batch(B=3, N=12000, U=6, V=48, seed=0, ragged=True):
    from tensorflowasr_b200 import schemas

    g = torch.Generator().manual_seed(seed)
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.full((B,), N, dtype=torch.int32)
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    lab_len = torch.full((B,), U, dtype=torch.int32)
    if ragged:
        sig_len[1:] = torch.tensor([N - 2500, N // 2][: B - 1], dtype=torch.int32)
        lab_len[1:] = torch.tensor([U - 2, 1][: B - 1], dtype=torch.int32)
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32), labels], 1)
    data = schemas.TrainData(schemas.TrainInput(sig.cuda(), sig_len.cuda(), pred.cuda(), (lab_len + 1).cuda()),
                             schemas.TrainLabel(labels.cuda(), lab_len.cuda()))
    return data, (sig, sig_len, pred, lab_len + 1, labels, lab_len)


def _oracle_losses(params, ocfg, host, training, dtype=torch.float64, requires_grad=()):
    sig, sig_len, pred, pred_len, labels, lab_len = host
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.to(dtype).clone() for k, v in params.items()}
    for k in requires_grad:
        p[k].requires_grad_(True)
    logits, llen = oc.transducer_forward(p, ocfg, torch.from_numpy(feats).to(dtype), torch.from_numpy(flen).long(), pred, pred_len.long(),
                                         training=training)
    return oc.rnnt_loss_torch(logits, labels, llen.int(), lab_len), logits, llen, p


@pytest.mark.parametrize("variant", ["batch", "layer", "streaming"])
def test_forward_logits_vs_oracle(variant):
    from tensorflowasr_b200 import schemas

    kw = {}
    if variant == "layer":
        kw = dict(dw_norm="layer", sub_norms=("layer", "layer"))
    if variant == "streaming":
        kw = dict(dw_norm="layer", sub_norms=("layer", "layer"), chunk=4, hist=8)
    model, ocfg, params = _setup(**kw)
    data, host = _batch()
    model.eval()
    out = model(data.inputs, training=False)
    ref_losses, ref_logits, ref_len, _ = _oracle_losses(params, ocfg, host, training=False)
    assert out.logits_length.tolist() == ref_len.tolist()
    got = out.logits.detach().cpu().double()
    # padded encoder frames are not comparable in general? they are: the oracle reproduces the reference's unmasked arithmetic
    np.testing.assert_allclose(got.numpy(), ref_logits.numpy(), rtol=1e-3, atol=2e-3)
    # the drop-in route (RnntLoss on the returned logits) and the fused training head give the same per-utterance losses
    from tensorflowasr_b200 import RnntLoss

    dropin = RnntLoss(blank=0).call(data.labels, schemas.TrainOutput(out.logits, out.logits_length)).detach().cpu().double()
    fused = model.compute_losses(data, training=False).detach().cpu().double()
    np.testing.assert_allclose(fused.numpy(), ref_losses.numpy(), rtol=1e-3)
    np.testing.assert_allclose(dropin.numpy(), ref_losses.numpy(), rtol=1e-3)


def test_training_mode_losses_and_gradients_vs_oracle():
    """Training mode (batch statistics in BN), dropout 0: per-utterance losses and parameter gradients vs oracle autograd (float64)."""
    model, ocfg, params = _setup()
    data, host = _batch()
    model.train()
    losses = model.compute_losses(data, training=True)
    losses.mean().backward()
    names = ["encoder.linear.kernel", "encoder.block_0.mhsa_module.mhsa.query.kernel", "encoder.block_1.conv_module.dw_conv.kernel",
             "encoder.block_0.ff_module_1.dense_1.bias", "encoder.content_attention_bias", "encoder.positional_attention_bias",
             "encoder.subsampling.conv_0.kernel", "encoder.block_1.conv_module.dw_bn.gamma", "prediction.lstm_0.recurrent_kernel",
             "prediction.embedding.embeddings", "joint.vocab.kernel", "joint.vocab.bias", "joint.enc.kernel", "joint.pred.bias",
             "encoder.block_1.ln.beta", "encoder.block_0.mhsa_module.mhsa.encoding.kernel"]
    ref_losses, _, _, p = _oracle_losses(params, ocfg, host, training=True, requires_grad=names)
    ref_losses.mean().backward()
    np.testing.assert_allclose(losses.detach().cpu().double().numpy(), ref_losses.detach().numpy(), rtol=1e-3)
    w = model.named_weights()
    for n in names:
        g, r = w[n].grad.detach().cpu().double(), p[n].grad
        scale = r.abs().max().item()
        err = (g - r).abs().max().item()
        assert err <= 5e-3 * scale + 1e-7, (n, err, scale)


def test_bf16_mode_close_to_fp32_and_train_step_learns():
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    model, ocfg, params = _setup(dropout=0.1)
    data, host = _batch(ragged=False)
    model.eval()
    l32 = model.compute_losses(data, training=False).detach()
    ops.set_precision("bf16")
    try:
        l16 = model.compute_losses(data, training=False).detach()
        torch.testing.assert_close(l16, l32, rtol=3e-2, atol=3e-2)
        model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=2e-3))
        model.train()
        first = None
        for _ in range(12):
            cur = model.train_step(data)["loss"].mean().item()
            first = cur if first is None else first
        assert np.isfinite(cur) and cur < first, (first, cur)
    finally:
        ops.set_precision("fp32")


def test_greedy_decode_tokens_bit_exact_vs_oracle():
    """Greedy transducer decode, both reference loops (recognize_batch base_transducer.py:496-575, recognize_single :577-712):
    token indices must equal the float64 oracle's restatement of the same loops EXACTLY (north_star: bit-exact greedy tokens)."""
    from tensorflowasr_b200 import schemas

    model, ocfg, params = _setup(V=32, seed=5, blank_bias=-0.6)
    data, host = _batch(V=32, B=3, N=16000, ragged=True)
    sig, sig_len = host[0], host[1]
    model.eval()
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.double() for k, v in params.items()}
    enc, elen = oc.conformer_encoder(torch.from_numpy(feats), torch.from_numpy(flen).long(), p, ocfg, training=False)
    out = model.recognize(schemas.PredictInput(sig.cuda(), sig_len.cuda()))
    ref, margin = oc.greedy_batch(enc, elen, p, ocfg)
    assert out.tokens.shape == ref.shape
    if margin > 1e-3:  # a near-tie in the oracle's own argmax would make fp32-vs-fp64 disagreement legitimate
        assert out.tokens.cpu().long().tolist() == ref.tolist()
    # batch-1 path
    out1 = model.recognize(schemas.PredictInput(sig[:1].cuda(), sig_len[:1].cuda()))
    feats1, flen1 = ologmel.feature_extraction(sig[:1].numpy().astype(np.float64), sig_len[:1].numpy(), dtype=np.float64)
    enc1, elen1 = oc.conformer_encoder(torch.from_numpy(feats1), torch.from_numpy(flen1).long(), p, ocfg, training=False)
    ref1, margin1 = oc.greedy_single(enc1, int(elen1[0]), p, ocfg)
    if margin1 > 1e-3:
        assert out1.tokens[0].cpu().tolist() == ref1


def _load_rand(model, seed):
    g = torch.Generator().manual_seed(seed)
    ws = {}
    for n, t in model.named_weights().items():
        if "moving_variance" in n:
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif n.endswith("gamma"):
            v = torch.randn(t.shape, generator=g) * 0.1 + 1
        elif n.endswith(("bias", "beta", "moving_mean")):
            v = torch.randn(t.shape, generator=g) * 0.1
        else:
            v = t.detach().cpu().clone()
        ws[n] = v
    model.load_named_weights(ws)
    return ws


def test_deepspeech2_ctc_model_vs_oracle():
    """BASELINE config 2 family: spectrogram frontend kernel -> DeepSpeech2 encoder -> vocab Dense -> CTC loss kernel; greedy CTC tokens
    bit-exact vs the oracle's tf.nn.ctc_greedy_decoder restatement."""
    import tensorflowasr_b200 as tfasr
    from oracle import ctc as octc
    from oracle import other_models as om
    from tensorflowasr_b200 import ops, schemas

    ops.set_precision("fp32")
    cfg = tfasr.deepspeech2_base_config(vocab_size=29)
    cfg.update(rnn_nlayers=2, rnn_units=32, fc_units=48, rnn_dropout=0.0, fc_dropout=0.0)
    model = tfasr.DeepSpeech2(**cfg, seed=2).cuda()
    ws = _load_rand(model, 9)
    B, N, L = 3, 16000, 12
    g = torch.Generator().manual_seed(0)
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    sig_len = torch.tensor([N, N - 3100, N // 2], dtype=torch.int32)
    labels = torch.randint(1, 29, (B, L), generator=g, dtype=torch.int32)
    lab_len = torch.tensor([L, 7, 3], dtype=torch.int32)
    x = schemas.TrainInput(sig.cuda(), sig_len.cuda(), None, None)
    model.eval()
    out = model(x, training=False)
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), feature_type="spectrogram", num_feature_bins=160, dtype=np.float64)
    p = {k: v.double() for k, v in ws.items()}
    ocfg = dict(conv_kernels=[(11, 41), (11, 21)], conv_strides=[(2, 2), (1, 2)], rnn_nlayers=2, rnn_units=32, fc_nlayers=1)
    ref, rl = om.deepspeech2_forward(p, ocfg, torch.from_numpy(feats), torch.from_numpy(flen).long(), training=False)
    assert out.logits_length.tolist() == rl.tolist()
    for b in range(B):
        n = int(rl[b])
        np.testing.assert_allclose(out.logits[b, :n].detach().cpu().double().numpy(), ref[b, :n].numpy(), rtol=1e-3, atol=2e-3)
    loss = tfasr.CtcLoss().call(schemas.TrainLabel(labels.cuda(), lab_len.cuda()), out)
    rc, _ = octc.ctc_loss(ref.numpy(), labels.numpy(), rl.numpy(), lab_len.numpy())
    np.testing.assert_allclose(loss.detach().cpu().numpy(), rc, rtol=1e-3)
    toks = model.recognize(schemas.PredictInput(sig.cuda(), sig_len.cuda())).tokens.cpu()
    want = om.ctc_greedy(ref, rl)
    for b in range(B):
        assert [t for t in toks[b].tolist()][: len(want[b])] == want[b] and all(t == 0 for t in toks[b].tolist()[len(want[b]):])
    # one training step runs and is finite
    ops.set_precision("bf16")
    try:
        model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-3))
        model.train()
        l0 = model.train_step(schemas.TrainData(x, schemas.TrainLabel(labels.cuda(), lab_len.cuda())))["loss"]
        assert torch.isfinite(l0).all()
    finally:
        ops.set_precision("fp32")


def test_contextnet_transducer_vs_oracle():
    """BASELINE config 4 family (ContextNet transducer): encoder + fused loss head vs the oracle, fp32."""
    import tensorflowasr_b200 as tfasr
    from oracle import other_models as om
    from tensorflowasr_b200 import ops, schemas

    ops.set_precision("fp32")
    blocks = [dict(nlayers=1, kernel_size=5, filters=64, strides=1, residual=False, activation="silu", padding="causal"),
              dict(nlayers=3, kernel_size=5, filters=64, strides=2, residual=True, activation="silu", padding="causal"),
              dict(nlayers=2, kernel_size=5, filters=96, strides=2, residual=True, activation="silu", padding="causal")]
    V = 40
    cfg = tfasr.contextnet_config(vocab_size=V, alpha=0.5)
    cfg.update(encoder_blocks=blocks, prediction_embed_dim=32, prediction_rnn_units=32, joint_dim=48)
    model = tfasr.ContextNet(**cfg, seed=4).cuda()
    ws = _load_rand(model, 11)
    data, host = _batch(B=3, N=12000, U=5, V=V)
    model.eval()
    losses = model.compute_losses(data, training=False).detach().cpu().double()
    sig, sig_len, pred, pred_len, labels, lab_len = host
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.double() for k, v in ws.items()}
    enc, elen = om.contextnet_forward(p, [dict(b, filters=int(b["filters"] * 0.5)) for b in blocks], torch.from_numpy(feats),
                                      torch.from_numpy(flen).long(), training=False)
    pr = oc.prediction_net(pred, pred_len.long(), p, oc.ConformerConfig(vocab_size=V, embed_dim=32, rnn_units=32))
    logits = oc.joint_net(enc, pr, p)
    ref = oc.rnnt_loss_torch(logits, labels, elen.int(), lab_len)
    np.testing.assert_allclose(losses.numpy(), ref.numpy(), rtol=1e-3)
