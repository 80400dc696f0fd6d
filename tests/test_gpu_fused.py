"""GPU parity: fused row-wise kernels vs plain PyTorch fp32 references of the same ops (same bf16-rounded inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=1e-2):
    scale = ref.abs().max().item() + 1e-6
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("shape", [(8000, 144), (2080, 320), (37, 576), (5, 1000), (1, 8)])
@pytest.mark.parametrize("xdtype", [torch.bfloat16, torch.float32])
def test_layernorm_fwd_bwd(shape, xdtype):
    from tensorflowasr_b200 import kernels

    M, D = shape
    torch.manual_seed(M)
    x = (torch.randn(M, D, device="cuda") * 2 + 0.5).to(xdtype)
    gamma = torch.randn(D, device="cuda") * 0.2 + 1
    beta = torch.randn(D, device="cuda") * 0.1
    y, mean, rstd = kernels.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-3)
    _close(y, yr)
    dy = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    dres = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    yr.backward(dy.float())
    dgamma = torch.zeros(D, device="cuda")
    dbeta = torch.zeros(D, device="cuda")
    dx = kernels.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=dres)
    _close(dx, xr.grad + dres.float())
    _close(dgamma, gr.grad, 5e-3)
    _close(dbeta, br.grad, 5e-3)


def test_scale_mask_colsum_and_glu_bwd():
    from tensorflowasr_b200 import kernels

    torch.manual_seed(0)
    M, N = 999, 144
    x = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    add = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    cs = torch.zeros(N, device="cuda")
    out = kernels.scale_mask_colsum(x, scale=0.5, colsum=cs, add=add)
    ref = 0.5 * x.float() + add.float()
    _close(out, ref)
    _close(cs, ref.sum(0), 1e-3)  # the kernel sums the fp32 values before the bf16 rounding of `out`
    # dropout mask must equal the mask the GEMM epilogue generates for the same (seed, salt, N)
    seed = torch.tensor([77], dtype=torch.int64, device="cuda")
    a = torch.randn(M, 64, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, 64, device="cuda").to(torch.bfloat16)
    fwd = kernels.gemm(a, b, dropout_p=0.3, dropout_seed=seed, dropout_salt=5, out_dtype=torch.float32)
    ones = torch.ones(M, N, device="cuda").to(torch.bfloat16)
    mask = kernels.scale_mask_colsum(ones, dropout_p=0.3, seed=seed, salt=5)
    assert torch.equal(fwd != 0, mask != 0)
    _close(mask[mask != 0], torch.full_like(mask[mask != 0], 1 / 0.7).float())
    # GLU backward
    D = 144
    aux = torch.randn(M, 2 * D, device="cuda").to(torch.bfloat16)
    dg = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    cs2 = torch.zeros(2 * D, device="cuda")
    daux = kernels.glu_bwd(aux, dg, colsum=cs2)
    ar = aux.float().requires_grad_(True)
    (ar[:, 0::2] * torch.sigmoid(ar[:, 1::2])).backward(dg.float())
    _close(daux, ar.grad)
    _close(cs2, ar.grad.sum(0), 1e-3)


def _mini_model(num_blocks=2, dropout=0.0, V=40, seed=1):
    import tensorflowasr_b200 as tfasr

    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)  # SpecAugment is RNG-dependent: parity runs with it off
    cfg.update(encoder_num_blocks=num_blocks, encoder_dropout=dropout)
    return tfasr.Conformer(**cfg, seed=seed).cuda()


def _mini_batch(B=4, N=8000, U=5, V=40, seed=0):
    from tensorflowasr_b200 import schemas

    g = torch.Generator().manual_seed(seed)
    sig = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1).cuda()
    sig_len = torch.full((B,), N, dtype=torch.int32).cuda()
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32).cuda()
    lab_len = torch.full((B,), U, dtype=torch.int32).cuda()
    pred = torch.cat([torch.zeros(B, 1, dtype=torch.int32).cuda(), labels], 1)
    return schemas.TrainData(schemas.TrainInput(sig, sig_len, pred, lab_len + 1), schemas.TrainLabel(labels, lab_len))


def test_fused_modules_match_library_path():
    """Same weights, same batch, bf16: the fused-kernel path and the torch-library path must agree (losses + gradients)."""
    from tensorflowasr_b200 import ops

    ops.set_precision("bf16")
    try:
        model = _mini_model()
        # perturb norms / biases so every term matters
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith(("bias", "beta")):
                    p.add_(0.05 * torch.randn_like(p))
                if n.endswith("gamma"):
                    p.add_(0.1 * torch.randn_like(p))
        data = _mini_batch()
        model.train()
        res = {}
        for flag in (False, True):
            ops.set_fused(flag)
            for p in model.parameters():
                p.grad = None
            losses = model.compute_losses(data, training=True)
            losses.mean().backward()
            res[flag] = (losses.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
        torch.testing.assert_close(res[True][0], res[False][0], rtol=2e-2, atol=2e-2)
        checked = 0
        for n, g in res[False][1].items():
            if n not in res[True][1]:  # conv biases in front of a batch norm: identically zero, not computed by the fused path
                assert n.endswith("bias") and ("conv_subsampling" in n or "dw_conv" in n), n
                continue
            g2 = res[True][1][n]
            scale = g.abs().max().item()
            # analytically-zero gradients are pure rounding noise: softmax is invariant to key / position biases, BatchNorm to the dw-conv bias
            if scale < 1e-6 or any(k in n for k in ("key__bias", "encoding__bias", "dw_conv__bias")):
                continue
            err = (g2 - g).abs().max().item()
            assert err <= 6e-2 * scale + 1e-6, (n, err, scale)
            checked += 1
        assert checked > 50
    finally:
        ops.set_fused(True)
        ops.set_precision("fp32")


def test_cuda_graph_train_step_matches_eager():
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    ops.set_precision("bf16")
    try:
        data = _mini_batch()
        out = {}
        for use_graph in (False, True):
            model = _mini_model(dropout=0.0, seed=3)
            model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-3))
            model.train()
            model.enable_cuda_graph(use_graph)
            ls = []
            for _ in range(6):
                ls.append(model.train_step(data)["loss"].mean().item())
            out[use_graph] = ls
        # graph mode is step-equivalent to eager mode: the warm-up steps before capture are undone (parameters, Adam moments and step
        # count, BatchNorm moving statistics, schedule position), so the trajectories agree step by step (bf16 atomics: not bit for bit)
        for a, b in zip(out[True], out[False]):
            assert a == pytest.approx(b, rel=2e-2), (out[True], out[False])
        assert out[True][-1] < out[True][0] and out[False][-1] < out[False][0]
    finally:
        ops.set_precision("fp32")


def test_fit_loop_with_prefetch_matches_train_step_loop():
    """BaseModel.fit (training.py: host batches from pinned memory on a copy stream one batch ahead, losses read one step late, graph replay)
    against the plain loop `train_step(device batch)`: same per-step losses, different batches per step (so a stale or swapped buffer shows),
    every byte accounted."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops, training

    ops.set_precision("bf16")
    try:
        dev_batches = [_mini_batch(seed=s) for s in (0, 1, 2, 3, 4)]
        host_batches = [training._map(lambda t: t.cpu().pin_memory(), d) for d in dev_batches]
        out = {}
        for mode in ("loop", "fit"):
            model = _mini_model(dropout=0.0, seed=3)
            model.compile(optimizer=tfasr.make_optimizer(model, learning_rate=1e-3))
            model.train()
            model.enable_cuda_graph(True)
            if mode == "loop":
                out[mode] = [model.train_step(d)["loss"].mean().item() for d in dev_batches]
            else:
                seen = []

                class Log:
                    def on_train_batch_end(self, step, logs):
                        seen.append(step)

                hist = model.fit(host_batches, epochs=1, callbacks=[Log()])
                out[mode] = hist.steps["loss"]
                assert seen == [0, 1, 2, 3, 4]
                assert hist.h2d_bytes == sum(t.numel() * t.element_size() for d in host_batches for t in training._tensors(d))
                assert hist.d2h_bytes == 4 * 5
        for a, b in zip(out["fit"], out["loop"]):
            assert a == pytest.approx(b, rel=2e-2), out
        assert len(set(round(v, 3) for v in out["fit"])) > 1  # the batches differ: so do their losses
    finally:
        ops.set_precision("fp32")


def test_transducer_head_kernels_vs_torch():
    """joint tanh / GEMM / RNN-T loss (bf16 in-place gradient) / reductions vs a plain PyTorch fp32 evaluation of the same head."""
    from tensorflowasr_b200 import fused, rnnt_loss

    torch.manual_seed(0)
    B, T, U1, J, V = 3, 21, 8, 64, 40
    e = (torch.randn(B, T, J, device="cuda") * 0.5).to(torch.bfloat16)
    p = (torch.randn(B, U1, J, device="cuda") * 0.5).to(torch.bfloat16)
    W = torch.nn.Parameter(torch.randn(J, V, device="cuda") * 0.2)
    bias = torch.nn.Parameter(torch.randn(V, device="cuda") * 0.1)
    labels = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    tl = torch.tensor([21, 15, 9], device="cuda", dtype=torch.int32)
    ul = torch.tensor([7, 4, 0], device="cuda", dtype=torch.int32)
    e1, p1 = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
    costs = fused.transducer_head(e1, p1, fused.WeightPack(W, bias), labels, tl, ul)
    up = torch.tensor([1.0, 0.5, 2.0], device="cuda")
    (costs * up).sum().backward()
    # reference: fp32 torch ops on the same bf16-rounded operands + the (separately validated) fp32 loss kernels
    e2, p2 = e.float().requires_grad_(True), p.float().requires_grad_(True)
    W2 = W.detach().to(torch.bfloat16).float().requires_grad_(True)
    b2 = bias.detach().clone().requires_grad_(True)
    hh = torch.tanh(e2[:, :, None] + p2[:, None]).to(torch.bfloat16).float() + 0 * (e2[:, :, None] + p2[:, None])
    hh = torch.tanh(e2[:, :, None] + p2[:, None])
    logits = hh @ W2 + b2
    ref = rnnt_loss(logits, tl, labels, ul)
    (ref * up).sum().backward()
    torch.testing.assert_close(costs, ref.detach(), rtol=5e-3, atol=5e-2)
    _close(e1.grad, e2.grad, 3e-2)
    _close(p1.grad, p2.grad, 3e-2)
    _close(W.grad, W2.grad, 3e-2)
    _close(bias.grad, b2.grad, 3e-2)


def test_subsampling_kernels_vs_torch():
    """conv1 direct / im2col / col2im / BN+swish fwd+bwd vs PyTorch fp32 ops."""
    import torch.nn.functional as F

    from tensorflowasr_b200 import kernels

    torch.manual_seed(0)
    B, T, Fq, C = 2, 37, 21, 16
    x = torch.randn(B, T, Fq, device="cuda")
    w = torch.randn(3, 3, 1, C, device="cuda") * 0.3
    bias = torch.randn(C, device="cuda") * 0.1
    y = kernels.conv1_fwd(x, w, bias)
    ref = F.conv2d(F.pad(x[:, None], (2, 0, 2, 0)), w.permute(3, 2, 0, 1), bias, stride=2).permute(0, 2, 3, 1)
    _close(y, ref)
    dy = torch.randn(ref.shape, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(3, 3, 1, C, device="cuda")
    kernels.conv1_wgrad(x, dy, dw)
    wr = w.clone().requires_grad_(True)
    F.conv2d(F.pad(x[:, None], (2, 0, 2, 0)), wr.permute(3, 2, 0, 1), bias, stride=2).permute(0, 2, 3, 1).backward(dy.float())
    _close(dw, wr.grad, 2e-3)
    # im2col + GEMM == conv2d ; col2im == its data gradient
    a = torch.randn(B, 19, 11, C, device="cuda").to(torch.bfloat16)
    w2 = (torch.randn(3, 3, C, C, device="cuda") * 0.1)
    col = kernels.im2col(a)
    y2 = col.float() @ w2.reshape(9 * C, C).to(torch.bfloat16).float()
    ar = a.float().requires_grad_(True)
    ref2 = F.conv2d(F.pad(ar.permute(0, 3, 1, 2), (2, 0, 2, 0)), w2.to(torch.bfloat16).float().permute(3, 2, 0, 1), None, stride=2).permute(0, 2, 3, 1)
    _close(y2.view_as(ref2), ref2, 2e-3)
    dy2 = torch.randn(ref2.shape, device="cuda").to(torch.bfloat16)
    ref2.backward(dy2.float())
    dcol = (dy2.float().view(-1, C) @ w2.reshape(9 * C, C).to(torch.bfloat16).float().t()).to(torch.bfloat16)
    da = kernels.col2im(dcol, tuple(a.shape))
    _close(da, ar.grad, 2e-2)
    # BN + swish training fwd / bwd
    from tensorflowasr_b200 import fused

    R, C2 = 5000, 144
    xb = (torch.randn(R, C2, device="cuda") * 2 + 0.3).to(torch.bfloat16)
    gamma = torch.nn.Parameter(torch.randn(C2, device="cuda") * 0.2 + 1)
    beta = torch.nn.Parameter(torch.randn(C2, device="cuda") * 0.1)
    mm, mv = torch.zeros(C2, device="cuda"), torch.ones(C2, device="cuda")
    yb, saved = fused.bn_train_fwd(xb, gamma, beta, mm, mv)
    xr = xb.float().requires_grad_(True)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    mean, var = xr.mean(0), xr.var(0, unbiased=False)
    yr = F.silu((xr - mean) * torch.rsqrt(var + 1e-3) * g2 + b2)
    _close(yb, yr)
    _close(mm, 0.01 * mean.detach(), 1e-2)
    dyb = torch.randn(R, C2, device="cuda").to(torch.bfloat16)
    yr.backward(dyb.float())
    dxb = fused.bn_train_bwd(dyb, xb, saved, gamma, beta)
    _close(dxb, xr.grad, 2e-2)
    _close(gamma.grad, g2.grad, 5e-3)
    _close(beta.grad, b2.grad, 5e-3)


@pytest.mark.parametrize("cfg", [(3, 50, 4, 36, None, None), (2, 37, 2, 16, 4, 8), (1, 250, 4, 36, None, None), (2, 64, 2, 64, None, None),
                                 (2, 129, 3, 36, 16, 32), (1, 300, 2, 36, None, None), (2, 9, 2, 36, None, None), (3, 250, 4, 36, 32, 64)])
@pytest.mark.parametrize("tc", [False, True])
def test_rel_attention_kernels_vs_torch(cfg, tc):
    """Fused relative-position attention fwd/bwd vs the torch restatement (ops.rel_attention) in fp32 on the same bf16 inputs.
    tc: the forward on the tcgen05 kernel (csrc/attention_tc.cu) — the backward then runs on ITS probabilities."""
    from tensorflowasr_b200 import kernels, ops

    ops.set_attention_tc(tc)
    B, T, N, H, chunk, hist = cfg
    NH, R = N * H, 2 * T - 1
    torch.manual_seed(T)
    qkv = (torch.randn(B * T, 3 * NH, device="cuda") * 0.7).to(torch.bfloat16)
    pos = (torch.randn(B * R, NH, device="cuda") * 0.7).to(torch.bfloat16)
    u = torch.randn(N, H, device="cuda") * 0.2
    v = torch.randn(N, H, device="cuda") * 0.2
    length = torch.tensor([T, max(1, T // 2), max(1, T - 3)][:B], device="cuda", dtype=torch.int32)
    ctx, probs = kernels.rel_attention_fwd(qkv, pos, u, v, length, B, T, N, H, chunk, hist)
    ops.set_precision("fp32")
    q3 = qkv.float().view(B, T, 3 * NH).clone().requires_grad_(True)
    p3 = pos.float().view(B, R, NH).clone().requires_grad_(True)
    ur, vr = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
    ref = ops.rel_attention(q3[..., :NH], q3[..., NH:2 * NH], q3[..., 2 * NH:], p3, ur, vr, length, N, H, chunk_size=chunk, history_size=hist)
    _close(ctx.view(B, T, NH), ref, 2e-2)
    dctx = (torch.randn(B * T, NH, device="cuda") * 0.5).to(torch.bfloat16)
    ref.backward(dctx.float().view(B, T, NH))
    du = torch.zeros(N, H, device="cuda")
    dv = torch.zeros(N, H, device="cuda")
    dqkv, dpos = kernels.rel_attention_bwd(qkv, pos, u, v, length, B, T, N, H, probs, dctx, du, dv, chunk, hist, ctx=ctx)
    _close(dqkv.view(B, T, 3 * NH), q3.grad, 3e-2)
    _close(dpos.view(B, R, NH), p3.grad, 3e-2)
    _close(du, ur.grad, 3e-2)
    _close(dv, vr.grad, 3e-2)


def test_rel_attention_tc_long_streaming_forward_vs_torch():
    """Conformer-L streaming decode shape (BASELINE configs[4]): T' = 750, 8 heads x 64, chunk 16 / history 64 — only the tcgen05 kernel runs it."""
    from tensorflowasr_b200 import kernels, ops

    B, T, N, H, chunk, hist = 2, 750, 8, 64, 16, 64
    NH, R = N * H, 2 * T - 1
    torch.manual_seed(7)
    qkv = (torch.randn(B * T, 3 * NH, device="cuda") * 0.7).to(torch.bfloat16)
    pos = (torch.randn(B * R, NH, device="cuda") * 0.7).to(torch.bfloat16)
    u, v = torch.randn(N, H, device="cuda") * 0.2, torch.randn(N, H, device="cuda") * 0.2
    length = torch.tensor([T, T - 100], device="cuda", dtype=torch.int32)
    assert ops.attention_tc_active(T)  # "auto" takes the tcgen05 kernel beyond T' = 256
    ctx, probs = kernels.rel_attention_fwd(qkv, pos, u, v, length, B, T, N, H, chunk, hist)
    ops.set_precision("fp32")
    q3, p3 = qkv.float().view(B, T, 3 * NH), pos.float().view(B, R, NH)
    ref = ops.rel_attention(q3[..., :NH], q3[..., NH:2 * NH], q3[..., 2 * NH:], p3, u, v, length, N, H, chunk_size=chunk, history_size=hist)
    for b in range(B):
        n = int(length[b])
        _close(ctx.view(B, T, NH)[b, :n], ref[b, :n], 2e-2)
    torch.testing.assert_close(probs.float().sum(-1)[0], torch.ones(N, T, device="cuda"), rtol=2e-2, atol=2e-2)


def test_dwconv_kernels_vs_torch():
    import torch.nn.functional as F

    from tensorflowasr_b200 import kernels

    torch.manual_seed(1)
    B, T, C, K = 3, 150, 144, 31
    x = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(K, C, device="cuda") * 0.2
    bias = torch.randn(C, device="cuda") * 0.1
    y = kernels.dwconv_fwd(x, w, bias)
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = F.conv1d(F.pad(xr.transpose(1, 2), (K - 1, 0)), wr.t().unsqueeze(1), br, groups=C).transpose(1, 2)
    _close(y, ref)
    dy = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
    ref.backward(dy.float())
    dw = torch.zeros(K, C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx = kernels.dwconv_bwd(dy, x, w, dw, db)
    _close(dx, xr.grad)
    _close(dw, wr.grad, 5e-3)
    _close(db, br.grad, 5e-3)


def test_flat_adamw_matches_torch_adamw():
    """One-launch AdamW over flat buffers (csrc/param_ops.cu) vs torch.optim.AdamW, incl. the l2-regularizer gradient and zeroing."""
    from tensorflowasr_b200.optim import FlatAdamW

    torch.manual_seed(0)
    shapes = [(37, 5), (144,), (3, 7, 11), (1,)]
    ps = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    l2 = 1e-3
    opt = FlatAdamW(ps, lr=3e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-2)
    topt = torch.optim.AdamW(ref, lr=3e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-2)
    mask = torch.zeros(opt.flat_p.numel(), dtype=torch.uint8, device="cuda")
    mask[: ps[0].numel()] = 1  # only the first tensor carries the regularizer (offset 0)
    opt.set_l2(mask, l2)
    for it in range(5):
        gs = [torch.randn_like(p) for p in ps]
        for p, r, g in zip(ps, ref, gs):
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            r.grad = g.clone()
        ref[0].grad.add_(ref[0].detach(), alpha=2 * l2)
        opt.step(zero_grad=True)
        topt.step()
        for p, r in zip(ps, ref):
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=2e-6), (it, (p - r).abs().max())
            assert float(p.grad.abs().max()) == 0.0
    # parameters are views of one buffer
    assert ps[1].data_ptr() == opt.flat_p.data_ptr() + 4 * opt.offsets[1] and opt.offsets[1] % 64 == 0


def test_pack_plan_matches_eager_packs():
    """Batched weight-pack launch (Q|K|V concatenation, GLU interleave, transposes) vs the per-tensor torch packing."""
    from tensorflowasr_b200 import fused

    model = _mini_model(num_blocks=1)
    enc = model.encoder
    plan = fused.PackPlan()
    pre = "block_0"
    lazy = {
        "ff": fused.WeightPack(enc[pre + ".ff_module_1.dense_1.kernel"], enc[pre + ".ff_module_1.dense_1.bias"], plan=plan),
        "mhsa": fused.MhsaPack(enc, pre + ".mhsa_module", plan),
        "glu": fused.GluPack(enc[pre + ".conv_module.pw_conv_1.kernel"], enc[pre + ".conv_module.pw_conv_1.bias"], plan),
    }
    plan.run()
    torch.cuda.synchronize()
    eager = {
        "ff": fused.WeightPack(enc[pre + ".ff_module_1.dense_1.kernel"], enc[pre + ".ff_module_1.dense_1.bias"]),
        "mhsa": fused.MhsaPack(enc, pre + ".mhsa_module"),
        "glu": fused.GluPack(enc[pre + ".conv_module.pw_conv_1.kernel"], enc[pre + ".conv_module.pw_conv_1.bias"]),
    }
    pairs = [(lazy["ff"], eager["ff"]), (lazy["mhsa"].qkv, eager["mhsa"].qkv), (lazy["mhsa"].pos, eager["mhsa"].pos),
             (lazy["mhsa"].out, eager["mhsa"].out), (lazy["glu"], eager["glu"])]
    for a, b in pairs:
        assert torch.equal(a.w, b.w) and torch.equal(a.wt, b.wt)
    assert torch.equal(lazy["mhsa"].qkv._bias, eager["mhsa"].qkv._bias)
    assert torch.equal(lazy["glu"]._bias, eager["glu"]._bias)
    # the table follows the parameters when their storage moves
    with torch.no_grad():
        p = enc[pre + ".ff_module_1.dense_1.kernel"]
        p.data = (p.data * 2).clone()
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(lazy["ff"].w, (p.detach()).to(torch.bfloat16))


@pytest.mark.parametrize("cfg", [(32, 65, 320, 320), (5, 7, 64, 64), (33, 3, 48, 128)])
def test_lstm_cluster_kernels_vs_torch(cfg):
    """Cluster LSTM (csrc/lstm.cu) forward / backward vs torch's LSTM cell recurrence in fp32 (keras gate order i, f, c, o = torch's
    i, f, g, o), masked outputs, all parameter gradients."""
    from tensorflowasr_b200 import fused, kernels

    B, U, E, R = cfg
    assert kernels.lstm_supported(R)
    torch.manual_seed(B * 1000 + U)
    x = (torch.randn(B, U, E, device="cuda") * 0.5).to(torch.bfloat16)
    kernel = torch.nn.Parameter(torch.randn(E, 4 * R, device="cuda") * (1.0 / E ** 0.5))
    rk = torch.nn.Parameter(torch.randn(R, 4 * R, device="cuda") * (1.0 / R ** 0.5))
    bias = torch.nn.Parameter(torch.randn(4 * R, device="cuda") * 0.1)
    length = torch.randint(1, U + 1, (B,), device="cuda", dtype=torch.int32)
    length[0] = U
    xg = x.clone().requires_grad_(True)
    y = fused.lstm(xg, length, kernel, rk, bias)
    # fp32 reference on the same bf16 inputs
    xr = x.float().requires_grad_(True)
    kr, rr, br = kernel.detach().clone().requires_grad_(True), rk.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    h = torch.zeros(B, R, device="cuda")
    c = torch.zeros(B, R, device="cuda")
    outs = []
    for t in range(U):
        pre = xr[:, t] @ kr + h @ rr + br
        i, f, g, o = pre.split(R, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    mask = (torch.arange(U, device="cuda")[None, :] < length[:, None]).float()[..., None]
    ref = torch.stack(outs, dim=1) * mask
    _close(y, ref, 3e-2)
    dy = (torch.randn(B, U, R, device="cuda") * 0.5).to(torch.bfloat16)
    y.backward(dy)
    ref.backward(dy.float())
    _close(xg.grad, xr.grad, 5e-2)
    _close(kernel.grad, kr.grad, 5e-2)
    _close(rk.grad, rr.grad, 5e-2)
    _close(bias.grad, br.grad, 5e-2)
