"""GPU parity of the ContextNet / DeepSpeech2 encoder-body kernels (csrc/seqconv.cu) and of the fused modules
built on them (fused_seq.py) against plain PyTorch fp32 references of the same ops on the same bf16-rounded inputs, and of the fused
encoders against the library route of the same model (losses + every parameter gradient).

First run on a B200 in round 2 (the round-1 failures of test_seq_dwconv_kernels_vs_torch were a test bug: a non-contiguous dy handed to the C ABI).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=1e-2):
    scale = ref.abs().max().item() + 1e-6
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol * scale, (err, scale)


def _act(z, act):
    return F.silu(z) if act == 1 else (F.relu(z) if act == 2 else z)


# ------------------------------------------------------------------------------------------------ kernel level
@pytest.mark.parametrize("cfg", [(3, 150, 80, 5, 1), (2, 151, 320, 5, 2), (4, 64, 640, 5, 1), (1, 33, 8, 3, 3), (2, 90, 264, 11, 1)])
def test_seq_dwconv_kernels_vs_torch(cfg):
    from tensorflowasr_b200 import kernels

    B, T, C, K, s = cfg
    torch.manual_seed(T)
    x = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(K, C, device="cuda") * 0.3
    y = kernels.seq_dwconv_fwd(x, w, s)
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv1d(F.pad(xr.transpose(1, 2), (K - 1, 0)), wr.t().unsqueeze(1), None, stride=s, groups=C).transpose(1, 2)
    assert y.shape == ref.shape
    _close(y, ref)
    dy = torch.randn(ref.shape, device="cuda").to(torch.bfloat16)  # contiguous [B, To, C] (randn_like would inherit ref's transposed strides)
    ref.backward(dy.float())
    dw = torch.zeros(K, C, device="cuda")
    dx = kernels.seq_dwconv_bwd(dy, x, w, dw, s)
    _close(dx, xr.grad)
    _close(dw, wr.grad, 5e-3)
    dw2 = torch.zeros(K, C, device="cuda")
    assert kernels.seq_dwconv_bwd(dy, x, w, dw2, s, need_dx=False) is None
    _close(dw2, wr.grad, 5e-3)


def test_masked_mean_and_se_kernels_vs_torch():
    from tensorflowasr_b200 import kernels

    torch.manual_seed(0)
    for (B, T, C, H, act) in [(5, 187, 320, 40, 1), (3, 40, 640, 80, 1), (2, 33, 24, 3, 2)]:
        y = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        res = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        length = torch.tensor([T, T // 3, 0, T + 9, 1][:B], dtype=torch.int32, device="cuda")
        W1 = torch.randn(C, H, device="cuda") / C ** 0.5
        b1 = torch.randn(H, device="cuda") * 0.1
        W2 = torch.randn(H, C, device="cuda") / H ** 0.5
        b2 = torch.randn(C, device="cuda") * 0.1
        yr = y.float().requires_grad_(True)
        rr = res.float().requires_grad_(True)
        ps = [t.clone().requires_grad_(True) for t in (W1, b1, W2, b2)]
        n = length.clamp(0, T)
        mask = (torch.arange(T, device="cuda")[None, :] < n[:, None]).float()[..., None]
        m_ref = (yr * mask).sum(1) / n.clamp(min=1)[:, None]
        s_ref = torch.sigmoid(_act(m_ref @ ps[0] + ps[1], act) @ ps[2] + ps[3])
        out_ref = _act(yr * s_ref[:, None, :] + rr, act)
        m = kernels.seq_masked_mean(y, length)
        _close(m, m_ref, 1e-4)
        z1, s = kernels.se_fc_fwd(m, W1, b1, W2, b2, act)
        _close(s, s_ref, 1e-4)
        out = kernels.se_apply_fwd(y, s, res, act)
        _close(out, out_ref)
        dout = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        out_ref.backward(dout.float())
        dres, ds = kernels.se_apply_bwd(dout, y, s, res, act)
        _close(dres, rr.grad)
        g = [torch.zeros_like(t) for t in (W1, b1, W2, b2)]
        dm = kernels.se_fc_bwd(ds, s, z1, m, W1, W2, length, T, act, *g)
        dy = kernels.se_dy(dres, s, dm, length)
        _close(dy, yr.grad)
        for got, p in zip(g, ps):
            _close(got, p.grad, 1e-2)


@pytest.mark.parametrize("cfg", [(2, 100, 160, 1, 11, 41, 2, 2, "same", 32), (2, 50, 80, 32, 11, 21, 1, 2, "same", 32), (1, 21, 20, 16, 3, 3, 2, 2, "causal", 24)])
def test_conv_lowering_matches_conv2d(cfg):
    """im2col + GEMM == conv2d; col2im + GEMM == its data gradient; X^T dY GEMM == its weight gradient."""
    from tensorflowasr_b200 import kernels

    B, T, Fq, Cin, kh, kw, sh, sw, padding, Cout = cfg
    torch.manual_seed(kh * kw)
    x = torch.randn(B, T, Fq, Cin, device="cuda")
    if Cin > 1:
        x = x.to(torch.bfloat16)
    w = (torch.randn(kh, kw, Cin, Cout, device="cuda") / (kh * kw * Cin) ** 0.5).to(torch.bfloat16)
    g = kernels.conv_geom(B, T, Fq, Cin, kh, kw, sh, sw, padding)
    col = kernels.conv_im2col(x.contiguous(), g)
    kreal = kh * kw * Cin
    wp = torch.zeros(g.Kp, Cout, device="cuda", dtype=torch.bfloat16)
    wp[:kreal] = w.reshape(kreal, Cout)
    z = kernels.gemm(col, wp.t().contiguous(), out_dtype=torch.float32)
    xr = x.to(torch.bfloat16).float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    xp = F.pad(xr.permute(0, 3, 1, 2), (g.pl, max((g.Fo - 1) * sw + kw - Fq - g.pl, 0), g.pt, max((g.To - 1) * sh + kh - T - g.pt, 0)))
    ref = F.conv2d(xp, wr.permute(3, 2, 0, 1), stride=(sh, sw)).permute(0, 2, 3, 1)
    assert ref.shape == (B, g.To, g.Fo, Cout)
    _close(z.view(ref.shape), ref, 1e-2)
    dz = torch.randn(B * g.To * g.Fo, Cout, device="cuda").to(torch.bfloat16)
    ref.backward(dz.float().view(ref.shape))
    dw = kernels.gemm(col, dz, mn_major=True, out_dtype=torch.float32)
    _close(dw[:kreal].view(kh, kw, Cin, Cout), wr.grad, 1e-2)
    if Cin % 8 == 0:
        da = kernels.conv_col2im(kernels.gemm(dz, wp), g)
        _close(da, xr.grad, 2e-2)


# ------------------------------------------------------------------------------------------------ module level
def test_sepconv_bn_and_se_block_modules_vs_torch():
    from tensorflowasr_b200 import fused, fused_seq, ops

    torch.manual_seed(5)
    B, T, C, N, K, s, act = 4, 77, 80, 128, 5, 2, 1
    x = torch.randn(B, T, C, device="cuda").to(torch.bfloat16).requires_grad_(True)
    dw_w = torch.nn.Parameter(torch.randn(K, C, device="cuda") * 0.3)
    pw = torch.nn.Parameter(torch.randn(C, N, device="cuda") / C ** 0.5)
    bias = torch.nn.Parameter(torch.randn(N, device="cuda") * 0.1)
    gamma = torch.nn.Parameter(torch.rand(N, device="cuda") + 0.5)
    beta = torch.nn.Parameter(torch.randn(N, device="cuda") * 0.1)
    mm, mv = torch.zeros(N, device="cuda"), torch.ones(N, device="cuda")
    pack = fused.WeightPack(pw, bias)
    y = fused_seq.sepconv_bn(x, dw_w, pack, (gamma, beta, mm, mv), s, act, True)
    # reference: same bf16 roundings at the module's HBM boundaries (depthwise output, GEMM output)
    xr = x.detach().float().requires_grad_(True)
    dwr, pwr, gr, br = (p.detach().clone().requires_grad_(True) for p in (dw_w, pw, gamma, beta))
    d = F.conv1d(F.pad(xr.transpose(1, 2), (K - 1, 0)), dwr.t().unsqueeze(1), None, stride=s, groups=C).transpose(1, 2)
    z = d @ pwr.to(torch.bfloat16).float() + bias.detach()
    mean, var = z.mean((0, 1)), z.var((0, 1), unbiased=False)
    ref = _act((z - mean) * torch.rsqrt(var + ops.BN_EPS) * gr + br, act)
    _close(y, ref, 3e-2)
    _close(mm, mean.detach() * (1 - ops.BN_MOMENTUM), 3e-2)
    dy = torch.randn_like(ref).to(torch.bfloat16)
    y.backward(dy)
    ref.backward(dy.float())
    _close(x.grad, xr.grad, 5e-2)
    _close(dw_w.grad, dwr.grad, 5e-2)
    _close(pw.grad, pwr.grad, 5e-2)
    _close(gamma.grad, gr.grad, 3e-2)
    _close(beta.grad, br.grad, 3e-2)
    # squeeze-excite tail
    H = N // 8
    length = torch.tensor([39, 20, 0, 39], dtype=torch.int32, device="cuda")
    ps = [torch.nn.Parameter(t) for t in (torch.randn(N, H, device="cuda") / N ** 0.5, torch.randn(H, device="cuda") * 0.1,
                                          torch.randn(H, N, device="cuda") / H ** 0.5, torch.randn(N, device="cuda") * 0.1)]
    To = y.shape[1]
    yy = torch.randn(B, To, N, device="cuda").to(torch.bfloat16).requires_grad_(True)
    res = torch.randn(B, To, N, device="cuda").to(torch.bfloat16).requires_grad_(True)
    out = fused_seq.se_block(yy, res, length, *ps, act)
    yr, rr = yy.detach().float().requires_grad_(True), res.detach().float().requires_grad_(True)
    pr = [p.detach().clone().requires_grad_(True) for p in ps]
    n = length.clamp(0, To)
    mask = (torch.arange(To, device="cuda")[None, :] < n[:, None]).float()[..., None]
    se = torch.sigmoid(_act(((yr * mask).sum(1) / n.clamp(min=1)[:, None]) @ pr[0] + pr[1], act) @ pr[2] + pr[3])
    oref = _act(yr * se[:, None, :] + rr, act)
    _close(out, oref)
    do = torch.randn_like(oref).to(torch.bfloat16)
    out.backward(do)
    oref.backward(do.float())
    _close(yy.grad, yr.grad, 2e-2)
    _close(res.grad, rr.grad, 2e-2)
    for p, q in zip(ps, pr):
        _close(p.grad, q.grad, 2e-2)


# ------------------------------------------------------------------------------------------------ model level: fused route vs library route
def _compare_routes(model, data, skip=()):
    from tensorflowasr_b200 import ops

    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("bias", "beta")):
                p.add_(0.05 * torch.randn_like(p))
            if n.endswith("gamma"):
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    res = {}
    # reference: the library route in fp32 (cuDNN / cuBLAS, TF32 off); candidate: the hand-written kernels in bf16.  Tolerance: bf16
    # activations (2^-8 relative) through ~10 normalised layers, gradients that are sums of O(1000) signed terms: 10 % of the tensor's
    # largest entry element-wise and cosine similarity >= 0.99 (a wrong formula fails both by a wide margin; two bf16 routes compared with
    # each other, as this test did in round 1, double the noise and tripped on a single BatchNorm gamma at 9.1 %)
    for flag in (False, True):
        ops.set_precision("bf16" if flag else "fp32")
        ops.set_fused_seq(flag)
        for p in model.parameters():
            p.grad = None
        losses = model.compute_losses(data, training=True)
        losses.mean().backward()
        res[flag] = (losses.detach().float().clone(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None})
    torch.testing.assert_close(res[True][0], res[False][0], rtol=3e-2, atol=3e-2)
    checked = 0
    for n, g in res[False][1].items():
        scale = g.abs().max().item()
        if n not in res[True][1] or scale < 1e-6 or any(k in n for k in skip):
            continue  # biases in front of a batch norm: identically zero (pure rounding noise on the library route)
        got = res[True][1][n]
        err = (got - g).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.reshape(-1), g.reshape(-1), dim=0).item()
        assert err <= 1e-1 * scale + 1e-6 and cos >= 0.99, (n, err, scale, cos)
        checked += 1
    return checked


def test_contextnet_fused_route_matches_library_route():
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    from test_gpu_fused import _mini_batch

    ops.set_precision("bf16")
    try:
        blocks = [dict(nlayers=1, kernel_size=5, filters=64, strides=1, residual=False, activation="silu", padding="causal"),
                  dict(nlayers=3, kernel_size=5, filters=64, strides=2, residual=True, activation="silu", padding="causal"),
                  dict(nlayers=2, kernel_size=5, filters=96, strides=2, residual=True, activation="silu", padding="causal")]
        V = 40
        cfg = tfasr.contextnet_config(vocab_size=V, alpha=0.5)
        cfg.update(encoder_blocks=blocks, prediction_embed_dim=32, prediction_rnn_units=64, joint_dim=48)
        model = tfasr.ContextNet(**cfg, seed=4).cuda()
        assert model.encoder._fusable(80, True)
        checked = _compare_routes(model, _mini_batch(B=4, N=12000, U=5, V=V), skip=("conv__bias", ".conv.bias"))
        assert checked > 40
    finally:
        ops.set_fused_seq(False)
        ops.set_precision("fp32")


def test_deepspeech2_fused_route_matches_library_route():
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops, schemas

    ops.set_precision("bf16")
    try:
        cfg = tfasr.deepspeech2_base_config(vocab_size=29)
        cfg.update(rnn_nlayers=2, rnn_units=64, fc_units=48, rnn_dropout=0.0, fc_dropout=0.0)
        model = tfasr.DeepSpeech2(**cfg, seed=2).cuda()
        model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-3))
        B, N, L = 3, 16000, 12
        g = torch.Generator().manual_seed(0)
        sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
        sig_len = torch.tensor([N, N - 3100, N // 2], dtype=torch.int32)
        labels = torch.randint(1, 29, (B, L), generator=g, dtype=torch.int32)
        lab_len = torch.tensor([L, 7, 3], dtype=torch.int32)
        data = schemas.TrainData(schemas.TrainInput(sig.cuda(), sig_len.cuda(), None, None), schemas.TrainLabel(labels.cuda(), lab_len.cuda()))
        checked = _compare_routes(model, data, skip=("conv2d__bias", "conv2d.bias"))
        assert checked > 20
    finally:
        ops.set_fused_seq(False)
        ops.set_precision("fp32")
