"""CPU: the autograd wiring of fused_seq.py (ContextNet / DeepSpeech2 fused encoder routes) against the library
route of the same encoder — outputs, lengths, moving statistics and every parameter gradient.  The seqconv kernels run as their real
source on the CPU emulator; GEMM / BatchNorm launchers are torch stand-ins of their contracts (tests/mock_kernels.py).  The GPU twin of
this test is tests/test_gpu_seqconv.py."""
import shutil

import pytest
import torch

import mock_kernels

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available (emulator build)")


def _randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith(("bias", "beta")):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif n.endswith("gamma"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def _run_routes(enc, feats, length, seed=0, lib_precision="bf16"):
    from tensorflowasr_b200 import ops

    state0 = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    out = {}
    for flag in (False, True):
        ops.set_fused_seq(flag)
        ops.set_precision("bf16" if flag else lib_precision)
        enc.load_state_dict(state0)
        for p in enc.parameters():
            p.grad = None
        y, ylen = enc((feats, length), training=True)
        proj = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 1))
        (y.float() ** 2 * proj.abs()).sum().backward()
        grads = {n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None}
        moving = {k: v.detach().clone() for k, v in enc.state_dict().items() if "moving" in k}
        out[flag] = (y.detach().float(), torch.as_tensor(ylen).tolist(), grads, moving)
    return out


def _compare(out, zero_grad_keys, min_checked, tol_y, tol_g):
    y0, l0, g0, m0 = out[False]
    y1, l1, g1, m1 = out[True]
    assert l0 == l1 and y0.shape == y1.shape
    assert (y1 - y0).abs().max().item() <= tol_y * y0.abs().max().item()
    for k in m0:
        assert (m1[k] - m0[k]).abs().max().item() <= tol_y * (m0[k].abs().max().item() + 1e-3), k
    checked = 0
    for n, g in g0.items():
        scale = g.abs().max().item()
        if any(k in n for k in zero_grad_keys):  # bias in front of a batch norm: identically zero, not computed by the fused route
            continue
        assert n in g1, f"fused route produced no gradient for {n}"
        if tol_g is not None:
            err = (g1[n].float() - g.float()).abs().max().item()
            assert err <= tol_g * scale + 1e-6, (n, err, scale)
        checked += 1
    assert checked >= min_checked, checked


# "ideal": fp32 storage + torch implementations of the kernel contracts -> the fused route must reproduce the fp32 library route to
#          ~1e-3 (outputs, moving statistics, every parameter gradient): any wiring mistake is an O(1) difference.
# "emulated": bf16 storage with the REAL seqconv kernels on the CPU emulator -> outputs / statistics within bf16 noise of the fp32 route
#          (gradients of this small ill-conditioned problem carry 10-30 % bf16 noise on either route, so only their presence is checked).
MODES = {"ideal": dict(ideal=True, tol_y=2e-3, tol_g=5e-3), "emulated": dict(ideal=False, tol_y=6e-2, tol_g=None)}


@pytest.mark.parametrize("mode", list(MODES))
def test_contextnet_encoder_fused_route_wiring(monkeypatch, mode):
    from tensorflowasr_b200 import ctc_models, ops

    cfg = MODES[mode]
    mock_kernels.install(monkeypatch, ideal=cfg["ideal"])
    ops.set_precision("bf16")
    try:
        blocks = [dict(nlayers=1, kernel_size=5, filters=32, strides=1, residual=False, activation="silu", padding="causal"),
                  dict(nlayers=2, kernel_size=5, filters=32, strides=2, residual=True, activation="silu", padding="causal"),
                  dict(nlayers=2, kernel_size=3, filters=48, strides=1, residual=True, activation="relu", padding="causal")]
        enc = ctc_models.ContextNetEncoder(blocks=blocks, alpha=0.5, num_feature_bins=24, gen=torch.Generator().manual_seed(3))
        _randomize(enc, 5)
        B, T = 4, 64
        feats = torch.randn(B, T, 24, 1, generator=torch.Generator().manual_seed(1))
        length = torch.tensor([T, T - 13, 40, 6], dtype=torch.int32)
        assert enc._fusable(24, True)
        out = _run_routes(enc, feats, length, lib_precision="fp32")
        _compare(out, ("conv__bias", "conv.bias"), 30, cfg["tol_y"], cfg["tol_g"])
    finally:
        ops.set_fused_seq(False)
        ops.set_precision("fp32")


@pytest.mark.parametrize("mode", list(MODES))
def test_deepspeech2_encoder_fused_route_wiring(monkeypatch, mode):
    from tensorflowasr_b200 import ctc_models, ops

    cfg = MODES[mode]
    mock_kernels.install(monkeypatch, ideal=cfg["ideal"])
    ops.set_precision("bf16")
    try:
        enc = ctc_models.DeepSpeech2Encoder(conv_kernels=[[5, 7], [3, 5]], conv_strides=[[2, 2], [1, 2]], conv_filters=[8, 16], rnn_nlayers=1,
                                            rnn_units=12, rnn_dropout=0.0, fc_nlayers=1, fc_units=20, fc_dropout=0.0, num_feature_bins=21,
                                            gen=torch.Generator().manual_seed(2))
        _randomize(enc, 7)
        B, T = 2, 30
        feats = torch.randn(B, T, 21, 1, generator=torch.Generator().manual_seed(4))
        length = torch.tensor([T, 17], dtype=torch.int32)
        assert enc._fusable(True)
        out = _run_routes(enc, feats, length, lib_precision="fp32")
        _compare(out, ("conv2d__bias", "conv2d.bias"), 12, cfg["tol_y"], cfg["tol_g"])
    finally:
        ops.set_fused_seq(False)
        ops.set_precision("fp32")


def test_bilstm_reversal_matches_packed_bidirectional(monkeypatch):
    """fused_seq.bilstm = forward LSTM + (reverse valid frames -> LSTM -> reverse back): with a unidirectional torch LSTM standing in for the
    cluster kernel it must equal the packed-sequence bidirectional LSTM of the library route (Keras Bidirectional mask semantics)."""
    from tensorflowasr_b200 import ctc_models, fused_seq, ops

    ops.set_precision("fp32")
    enc = ctc_models.DeepSpeech2Encoder(conv_kernels=[[3, 3]], conv_strides=[[1, 1]], conv_filters=[8], rnn_nlayers=1, rnn_units=10, fc_nlayers=0,
                                        num_feature_bins=4, gen=torch.Generator().manual_seed(0))
    B, T, D = 3, 9, 32
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(1))
    length = torch.tensor([9, 4, 1], dtype=torch.int32)

    def uni_lstm(xx, ln, kernel, recurrent_kernel, bias):
        R = recurrent_kernel.shape[0]
        ws = [kernel.float().t().contiguous(), recurrent_kernel.float().t().contiguous(), bias.float(), torch.zeros_like(bias).float()]
        packed = torch.nn.utils.rnn.pack_padded_sequence(xx.float(), ln.to(torch.int64).clamp(min=1), batch_first=True, enforce_sorted=False)
        h0 = xx.new_zeros(1, xx.shape[0], R)
        out, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes, (h0, h0.clone()), ws, True, 1, 0.0, False, False)
        out = torch.nn.utils.rnn.PackedSequence(out, packed.batch_sizes, packed.sorted_indices, packed.unsorted_indices)
        return torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=xx.shape[1])[0]

    monkeypatch.setattr(fused_seq, "lstm", uni_lstm)
    ws = [tuple(enc[f"rnn_module.block_0.blstm.{d}_lstm.{n}"] for n in ("kernel", "recurrent_kernel", "bias")) for d in ("forward", "backward")]
    with torch.no_grad():
        got = fused_seq.bilstm(x, length, ws[0], ws[1])
        ref = enc._bilstm(x, length, 0)
    assert got.shape == ref.shape == (B, T, 20)
    torch.testing.assert_close(got, ref.float(), rtol=1e-5, atol=1e-6)
