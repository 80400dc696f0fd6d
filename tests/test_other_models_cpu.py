"""CPU: host logic of the DeepSpeech2 / ContextNet encoders (library path) vs the oracle restatements (oracle/other_models.py)."""
import torch

from oracle import other_models as om


def _rand_params(mod, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for n, t in mod.named().items():
        if "moving_variance" in n:
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif n.endswith("gamma"):
            v = torch.randn(t.shape, generator=g) * 0.1 + 1
        elif n.endswith(("bias", "beta", "moving_mean")):
            v = torch.randn(t.shape, generator=g) * 0.1
        else:
            v = t.detach().clone()
        out[n] = v
    mod.load_named(out)
    return out


def test_deepspeech2_encoder_and_greedy_vs_oracle():
    from tensorflowasr_b200 import ops
    from tensorflowasr_b200.ctc_models import CtcDecoder, DeepSpeech2Encoder

    ops.set_precision("fp32")
    cfg = dict(conv_kernels=[(5, 7), (3, 5)], conv_strides=[(2, 2), (1, 2)], rnn_nlayers=2, rnn_units=8, fc_nlayers=1)
    enc = DeepSpeech2Encoder(conv_kernels=cfg["conv_kernels"], conv_strides=cfg["conv_strides"], conv_filters=(4, 6), rnn_nlayers=2, rnn_units=8,
                             fc_nlayers=1, fc_units=12, rnn_dropout=0.0, fc_dropout=0.0, num_feature_bins=20, gen=torch.Generator().manual_seed(1))
    dec = CtcDecoder(9, enc.dmodel, gen=torch.Generator().manual_seed(2))
    p = {f"encoder.{k}": v for k, v in _rand_params(enc, 3).items()}
    p.update({f"decoder.{k}": v for k, v in _rand_params(dec, 4).items()})
    torch.manual_seed(0)
    x = torch.randn(3, 31, 20, 1)
    length = torch.tensor([31, 20, 7])
    for training in (False, True):
        with torch.no_grad():
            y, l = enc((x, length), training=training)
            logits, _ = dec((y, l))
            ref, rl = om.deepspeech2_forward(p, cfg, x, length, training=training)
        assert l.tolist() == rl.tolist() == [16, 10, 4]
        for b in range(3):
            n = int(rl[b])
            torch.testing.assert_close(logits[b, :n], ref[b, :n], rtol=1e-4, atol=1e-4)


def test_contextnet_encoder_vs_oracle():
    from tensorflowasr_b200 import ops
    from tensorflowasr_b200.ctc_models import ContextNetEncoder

    ops.set_precision("fp32")
    blocks = [dict(nlayers=1, kernel_size=5, filters=16, strides=1, residual=False, activation="silu"),
              dict(nlayers=3, kernel_size=5, filters=16, strides=2, residual=True, activation="silu"),
              dict(nlayers=2, kernel_size=3, filters=32, strides=1, residual=True, activation="silu")]
    enc = ContextNetEncoder(blocks=blocks, alpha=0.5, num_feature_bins=12, gen=torch.Generator().manual_seed(5))
    p = {f"encoder.{k}": v for k, v in _rand_params(enc, 6).items()}
    torch.manual_seed(1)
    x = torch.randn(3, 29, 12, 1)
    length = torch.tensor([29, 18, 5])
    for training in (False, True):
        with torch.no_grad():
            y, l = enc((x, length), training=training)
            ref, rl = om.contextnet_forward(p, blocks_scaled(blocks, 0.5), x, length, training=training)
        assert l.tolist() == rl.tolist() == [15, 9, 3]
        torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)


def blocks_scaled(blocks, alpha):
    return [dict(b, filters=int(b["filters"] * alpha)) for b in blocks]
