"""GPU parity of the fused log-sum-exp path (first step): softmax row statistics from the joint GEMM's epilogue (gemm feature
class 8) + csrc/rnnt_combine.cu in place of the K1 pass over the logits.

First run on a B200 in round 2: green; the fused log-sum-exp, the conv1 / im2col / rnnt-grad second generation and the batched position
projection became defaults; the dwconv-wgrad and joint-reduce second generations measured slower and stay off."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(4097, 1000, 320), (300, 256, 64), (129, 40, 320), (1000, 1024, 128), (77, 520, 64)])
def test_gemm_row_statistics_give_logsumexp(shape):
    from tensorflowasr_b200 import kernels

    M, N, K = shape
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * (3.0 / K ** 0.5)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    parts = 2 * ((N + 255) // 256)
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    out = kernels.gemm(a, b, bias=bias, out_dtype=torch.float32, rowstats=stats)
    ref = a.float() @ b.float().t() + bias
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert torch.isfinite(stats[..., 1]).all()
    m = stats[..., 0].max(1, keepdim=True).values
    lse = (m + torch.log((stats[..., 1] * torch.exp(stats[..., 0] - m)).sum(1, keepdim=True)))[:, 0]
    torch.testing.assert_close(lse, torch.logsumexp(out, dim=1), rtol=1e-5, atol=1e-4)
    # the slices partition the row: their sum-exp terms add up to the row's
    torch.testing.assert_close((stats[..., 1] * torch.exp(stats[..., 0] - m)).sum(1), torch.exp(out - m).sum(1), rtol=1e-4, atol=1e-5)


def test_transducer_head_costs_and_gradients_match_with_fused_lse():
    from tensorflowasr_b200 import fused, ops

    torch.manual_seed(2)
    B, T, U1, J, V = 3, 37, 9, 64, 1000
    e = (torch.randn(B, T, J, device="cuda") * 0.5).to(torch.bfloat16)
    p = (torch.randn(B, U1, J, device="cuda") * 0.5).to(torch.bfloat16)
    labels = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    tl = torch.tensor([37, 20, 5], device="cuda", dtype=torch.int32)
    ul = torch.tensor([8, 3, 0], device="cuda", dtype=torch.int32)
    res = {}
    try:
        ops.set_bf16_logits(False)  # this test isolates the normaliser: f32 logits on both sides
        for flag in (False, True):
            ops.set_fused_lse(flag)
            w = torch.nn.Parameter((torch.randn(J, V, generator=torch.Generator().manual_seed(5)) / J ** 0.5).cuda())
            bvec = torch.nn.Parameter(torch.randn(V, generator=torch.Generator().manual_seed(6)).cuda() * 0.1)
            er, pr = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
            costs = fused.transducer_head(er, pr, fused.WeightPack(w, bvec), labels, tl, ul)
            costs.sum().backward()
            res[flag] = (costs.detach().clone(), er.grad.clone(), pr.grad.clone(), w.grad.clone(), bvec.grad.clone())
    finally:
        ops.set_fused_lse(False)
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-5, atol=1e-4)
    for a, b in zip(res[True][1:], res[False][1:]):
        assert (a.float() - b.float()).abs().max().item() <= 2e-3 * (b.float().abs().max().item() + 1e-6)


@pytest.mark.parametrize("shape", [(4097, 1000, 320), (300, 256, 64), (129, 40, 320)])
def test_gemm_row_statistics_of_bf16_output_describe_the_stored_values(shape):
    """bf16 logits (ops.bf16_logits): the epilogue rounds first and takes the statistics of what it stores."""
    from tensorflowasr_b200 import kernels

    M, N, K = shape
    torch.manual_seed(M + N + 1)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * (3.0 / K ** 0.5)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    parts = 2 * ((N + 255) // 256)
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    out = kernels.gemm(a, b, bias=bias, out_dtype=torch.bfloat16, rowstats=stats)
    ref = a.float() @ b.float().t() + bias
    assert (out.float() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()
    m = stats[..., 0].max(1, keepdim=True).values
    lse = (m + torch.log((stats[..., 1] * torch.exp(stats[..., 0] - m)).sum(1, keepdim=True)))[:, 0]
    torch.testing.assert_close(lse, torch.logsumexp(out.float(), dim=1), rtol=1e-5, atol=1e-4)


def test_transducer_head_with_bf16_logits():
    """The bf16-logits head against (a) the exact float32 RNN-T loss of the SAME rounded logits (tfasr_rnnt_loss_f32 through losses.rnnt_loss,
    itself pinned to the oracle in test_gpu_losses.py) — costs to 1e-5, the head's own claim — and (b) the f32-logits head: costs within
    2e-3 relative (the rounding of ~T+U log-probabilities), gradients within 2e-2 of the maximum and cosine >= 0.999."""
    from tensorflowasr_b200 import fused, kernels, losses, ops

    torch.manual_seed(12)
    B, T, U1, J, V = 4, 41, 11, 64, 1000
    e = (torch.randn(B, T, J, device="cuda") * 0.5).to(torch.bfloat16)
    p = (torch.randn(B, U1, J, device="cuda") * 0.5).to(torch.bfloat16)
    labels = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    tl = torch.tensor([41, 30, 7, 1], device="cuda", dtype=torch.int32)
    ul = torch.tensor([10, 4, 0, 2], device="cuda", dtype=torch.int32)
    weights = torch.tensor([1.0, 0.5, 2.0, 1.5], device="cuda")
    res = {}
    saved = ops.switch_state()
    try:
        ops.set_fused_lse(True)
        for flag in (False, True):
            ops.set_bf16_logits(flag)
            assert ops.bf16_logits_active() == flag
            w = torch.nn.Parameter((torch.randn(J, V, generator=torch.Generator().manual_seed(5)) / J ** 0.5).cuda())
            bvec = torch.nn.Parameter(torch.randn(V, generator=torch.Generator().manual_seed(6)).cuda() * 0.1)
            er, pr = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
            pack = fused.WeightPack(w, bvec)
            costs = fused.transducer_head(er, pr, pack, labels, tl, ul)
            (costs * weights).sum().backward()
            res[flag] = (costs.detach().clone(), er.grad.clone(), pr.grad.clone(), w.grad.clone(), bvec.grad.clone())
            if flag:
                h = kernels.joint_tanh_fwd(e, p)
                stored = kernels.gemm(h.view(-1, J), pack.wt, bias=pack.bias, out_dtype=torch.bfloat16).view(B, T, U1, V)
                exact = losses.rnnt_loss(stored.float(), tl, labels, ul)
    finally:
        ops.restore_switches(saved)
    torch.testing.assert_close(res[True][0], exact, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(res[True][0], res[False][0], rtol=2e-3, atol=1e-3)
    for a, b in zip(res[True][1:], res[False][1:]):
        a, b = a.float().reshape(-1), b.float().reshape(-1)
        assert (a - b).abs().max().item() <= 2e-2 * (b.abs().max().item() + 1e-6)
        assert torch.nn.functional.cosine_similarity(a, b, dim=0).item() >= 0.999


@pytest.mark.parametrize("bf16_logits", [False, True])
def test_transducer_head_with_shared_memory_wavefronts(bf16_logits):
    """csrc/rnnt_lattice_v2.cu through the training head against the first-generation wavefront kernel: same costs (1e-5), same gradients;
    ragged lengths incl. an empty utterance and one with more labels than frames' worth of slack."""
    from tensorflowasr_b200 import fused, ops

    torch.manual_seed(13)
    B, T, U1, J, V = 5, 57, 65, 64, 1000
    e = (torch.randn(B, T, J, device="cuda") * 0.5).to(torch.bfloat16)
    p = (torch.randn(B, U1, J, device="cuda") * 0.5).to(torch.bfloat16)
    labels = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    tl = torch.tensor([57, 30, 7, 1, 0], device="cuda", dtype=torch.int32)
    ul = torch.tensor([64, 33, 0, 2, 5], device="cuda", dtype=torch.int32)
    weights = torch.tensor([1.0, 0.5, 2.0, 1.5, 1.0], device="cuda")
    res = {}
    saved = ops.switch_state()
    try:
        ops.set_fused_lse(True)
        ops.set_bf16_logits(bf16_logits)
        for flag in (False, True):
            ops.set_lattice_v2(flag)
            w = torch.nn.Parameter((torch.randn(J, V, generator=torch.Generator().manual_seed(5)) / J ** 0.5).cuda())
            bvec = torch.nn.Parameter(torch.randn(V, generator=torch.Generator().manual_seed(6)).cuda() * 0.1)
            er, pr = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
            costs = fused.transducer_head(er, pr, fused.WeightPack(w, bvec), labels, tl, ul)
            finite = torch.isfinite(costs)
            (costs[finite] * weights[finite]).sum().backward()
            res[flag] = (costs.detach().clone(), er.grad.clone(), pr.grad.clone(), w.grad.clone(), bvec.grad.clone())
    finally:
        ops.restore_switches(saved)
    assert torch.isinf(res[True][0][4]) and torch.isinf(res[False][0][4])  # the empty utterance
    torch.testing.assert_close(res[True][0][:4], res[False][0][:4], rtol=1e-5, atol=1e-4)
    for a, b in zip(res[True][1:], res[False][1:]):
        assert (a.float() - b.float()).abs().max().item() <= 2e-3 * (b.float().abs().max().item() + 1e-6)


def test_conv1_v2_matches_first_generation_kernels():
    """csrc/conv1.cu (opt-in, TFASR_CONV1_V2) against the verified conv_bn.cu kernels: identical contract."""
    from tensorflowasr_b200 import kernels, ops

    torch.manual_seed(4)
    for (B, T, F, C) in [(4, 1000, 80, 144), (2, 37, 13, 16), (1, 1, 1, 8)]:
        x = torch.randn(B, T, F, device="cuda")
        w = torch.randn(3, 3, 1, C, device="cuda") * 0.3
        bias = torch.randn(C, device="cuda") * 0.1
        res = {}
        try:
            for flag in (False, True):
                ops.set_conv1_v2(flag)
                y = kernels.conv1_fwd(x, w, bias)
                dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).to(torch.bfloat16)
                dw = torch.zeros_like(w)
                kernels.conv1_wgrad(x, dy, dw)
                res[flag] = (y.float(), dw)
        finally:
            ops.set_conv1_v2(False)
        assert (res[True][0] - res[False][0]).abs().max().item() <= 1e-2 * res[False][0].abs().max().item()
        assert (res[True][1] - res[False][1]).abs().max().item() <= 2e-3 * res[False][1].abs().max().item() + 1e-6


def test_rnnt_grad_v2_matches_first_generation_kernel():
    """csrc/rnnt_grad_v2.cu (opt-in, TFASR_RNNT_GRAD_V2) against tfasr_rnnt_loss_bwd_bf16 through the training head: same costs, same
    gradients (in-place bf16 over the logits) incl. the vocab-bias gradient; also with the fused log-sum-exp forward."""
    from tensorflowasr_b200 import fused, ops

    torch.manual_seed(7)
    B, T, U1, J, V = 4, 41, 11, 64, 1000
    e = (torch.randn(B, T, J, device="cuda") * 0.5).to(torch.bfloat16)
    p = (torch.randn(B, U1, J, device="cuda") * 0.5).to(torch.bfloat16)
    labels = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    tl = torch.tensor([41, 30, 7, 1], device="cuda", dtype=torch.int32)
    ul = torch.tensor([10, 4, 0, 2], device="cuda", dtype=torch.int32)
    res = {}
    try:
        ops.set_bf16_logits(False)  # the f32-logits gradient kernels are the subject here
        for key, (g2, lse) in {"base": (False, False), "grad_v2": (True, False), "both": (True, True)}.items():
            ops.set_rnnt_grad_v2(g2)
            ops.set_fused_lse(lse)
            w = torch.nn.Parameter((torch.randn(J, V, generator=torch.Generator().manual_seed(5)) / J ** 0.5).cuda())
            bvec = torch.nn.Parameter(torch.randn(V, generator=torch.Generator().manual_seed(6)).cuda() * 0.1)
            er, pr = e.clone().requires_grad_(True), p.clone().requires_grad_(True)
            costs = fused.transducer_head(er, pr, fused.WeightPack(w, bvec), labels, tl, ul)
            (costs * torch.tensor([1.0, 0.5, 2.0, 1.5], device="cuda")).sum().backward()
            res[key] = (costs.detach().clone(), er.grad.clone(), pr.grad.clone(), w.grad.clone(), bvec.grad.clone())
    finally:
        ops.set_rnnt_grad_v2(False)
        ops.set_fused_lse(False)
    for key in ("grad_v2", "both"):
        torch.testing.assert_close(res[key][0], res["base"][0], rtol=1e-5, atol=1e-4)
        for a, b in zip(res[key][1:], res["base"][1:]):
            assert (a.float() - b.float()).abs().max().item() <= 2e-3 * (b.float().abs().max().item() + 1e-6), key


def test_joint_v2_matches_first_generation_kernels():
    """csrc/joint_v2.cu (opt-in, TFASR_JOINT_V2) against joint.cu: identical contracts."""
    from tensorflowasr_b200 import kernels, ops

    torch.manual_seed(8)
    for (B, T, U1, J) in [(4, 250, 65, 320), (2, 7, 11, 40), (3, 1, 1, 8)]:
        e = torch.randn(B, T, J, device="cuda").to(torch.bfloat16)
        p = torch.randn(B, U1, J, device="cuda").to(torch.bfloat16)
        dz = torch.randn(B, T, U1, J, device="cuda").to(torch.bfloat16)
        res = {}
        try:
            for flag in (False, True):
                ops.set_joint_v2(flag)
                res[flag] = (kernels.joint_tanh_fwd(e, p).float(),) + tuple(kernels.joint_reduce_bwd(dz))
        finally:
            ops.set_joint_v2(False)
        ref_h = torch.tanh(e.float()[:, :, None, :] + p.float()[:, None, :, :])
        assert (res[True][0] - ref_h).abs().max().item() <= 1e-2
        assert (res[True][0] - res[False][0]).abs().max().item() <= 1e-2
        torch.testing.assert_close(res[True][1], dz.float().sum(2), rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(res[True][2], dz.float().sum(1), rtol=1e-4, atol=1e-3)


def test_im2col_v2_matches_first_generation_kernels():
    """csrc/subsample_v2.cu (opt-in, TFASR_IM2COL_V2) against conv_bn.cu's im2col / col2im: identical contracts (bit-exact copy / same sums)."""
    from tensorflowasr_b200 import kernels, ops

    torch.manual_seed(9)
    for (B, T1, F1, C) in [(4, 500, 40, 144), (2, 7, 5, 24), (1, 1, 1, 8)]:
        a = torch.randn(B, T1, F1, C, device="cuda").to(torch.bfloat16)
        res = {}
        try:
            for flag in (False, True):
                ops.set_im2col_v2(flag)
                col = kernels.im2col(a)
                dcol = torch.randn(col.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).to(torch.bfloat16)
                res[flag] = (col, kernels.col2im(dcol, (B, T1, F1, C)).float())
        finally:
            ops.set_im2col_v2(False)
        assert torch.equal(res[True][0], res[False][0])
        assert (res[True][1] - res[False][1]).abs().max().item() <= 1e-2 * (res[False][1].abs().max().item() + 1e-6)


def test_dwconv_wgrad_v2_matches_first_generation_kernel():
    """csrc/dwconv_wgrad_v2.cu + tfasr_dwconv_causal_bwd_dx (opt-in, TFASR_DWCONV_WGRAD_V2) against tfasr_dwconv_causal_bwd."""
    from tensorflowasr_b200 import kernels, ops

    torch.manual_seed(10)
    for (B, T, C, K) in [(32, 250, 144, 31), (3, 37, 24, 5), (1, 3, 8, 7)]:
        x = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        dy = torch.randn(B, T, C, device="cuda").to(torch.bfloat16)
        w = torch.randn(K, C, device="cuda") * 0.2
        res = {}
        try:
            for flag in (False, True):
                ops.set_dwconv_wgrad_v2(flag)
                dw, db = torch.zeros(K, C, device="cuda"), torch.zeros(C, device="cuda")
                dx = kernels.dwconv_bwd(dy, x, w, dw, db)
                res[flag] = (dx.float(), dw, db)
        finally:
            ops.set_dwconv_wgrad_v2(False)
        assert torch.equal(res[True][0], res[False][0])
        for i in (1, 2):
            assert (res[True][i] - res[False][i]).abs().max().item() <= 2e-3 * (res[False][i].abs().max().item() + 1e-6)


def test_strided_position_slices_match_the_contiguous_attention_kernels():
    """tfasr_rel_attention_fwd_ld / bwd_ld on a column slice of a wide position buffer == the contiguous entry points, bit for bit (the same
    kernels, only the row stride differs); columns outside the slice of the shared dpos buffer stay untouched."""
    from tensorflowasr_b200 import kernels

    torch.manual_seed(11)
    B, T, N, H, blocks, blk = 4, 250, 4, 36, 16, 5
    NH, R = N * H, 2 * T - 1
    qkv = (torch.randn(B * T, 3 * NH, device="cuda") * 0.7).to(torch.bfloat16)
    pos_all = (torch.randn(B * R, blocks * NH, device="cuda") * 0.7).to(torch.bfloat16)
    sl = slice(blk * NH, (blk + 1) * NH)
    pos = pos_all[:, sl].contiguous()
    u, v = torch.randn(N, H, device="cuda") * 0.3, torch.randn(N, H, device="cuda") * 0.3
    length = torch.tensor([250, 180, 250, 97], dtype=torch.int32, device="cuda")
    dctx = torch.randn(B * T, NH, device="cuda").to(torch.bfloat16)
    ctx0, p0 = kernels.rel_attention_fwd(qkv, pos, u, v, length, B, T, N, H)
    ctx1, p1 = kernels.rel_attention_fwd(qkv, pos_all[:, sl], u, v, length, B, T, N, H)
    assert torch.equal(ctx0, ctx1) and torch.equal(p0, p1)
    du0, dv0, du1, dv1 = (torch.zeros(N, H, device="cuda") for _ in range(4))
    dq0, dp0 = kernels.rel_attention_bwd(qkv, pos, u, v, length, B, T, N, H, p0, dctx, du0, dv0, ctx=ctx0)
    shared = torch.full((B * R, blocks * NH), 7.0, device="cuda", dtype=torch.bfloat16)
    dq1, _ = kernels.rel_attention_bwd(qkv, pos_all[:, sl], u, v, length, B, T, N, H, p1, dctx, du1, dv1, ctx=ctx1, dpos_out=shared[:, sl])
    assert torch.equal(dq0, dq1) and torch.equal(dp0, shared[:, sl])
    torch.testing.assert_close(du0, du1, rtol=1e-4, atol=1e-4)  # atomics: order differs from run to run
    torch.testing.assert_close(dv0, dv1, rtol=1e-4, atol=1e-4)
    other = torch.ones(blocks * NH, dtype=torch.bool, device="cuda")
    other[sl] = False
    assert bool((shared[:, other] == 7.0).all())


def test_batched_position_projection_matches_the_per_block_route():
    """ops.set_pos_batched (one GEMM for every block's position projection, PosAllPack) against the per-block route: losses and every
    parameter gradient of a 3-block Conformer transducer step, bf16 fused route."""
    import tensorflowasr_b200 as tfasr
    from tensorflowasr_b200 import ops

    from test_gpu_model import _batch

    res = {}
    try:
        ops.set_precision("bf16")
        for flag in (False, True):
            ops.set_pos_batched(flag)
            cfg = tfasr.conformer_small_config(vocab_size=48)
            cfg["speech_config"].pop("augmentation_config", None)
            cfg.update(encoder_num_blocks=3, encoder_dropout=0.0)
            model = tfasr.Conformer(**cfg, seed=1).cuda()
            model.train()
            data, _ = _batch(V=48)
            losses = model.compute_losses(data, training=True)
            losses.mean().backward()
            torch.cuda.synchronize()
            res[flag] = (losses.detach().float(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None})
    finally:
        ops.set_pos_batched(False)
        ops.set_precision("fp32")
    torch.testing.assert_close(res[True][0], res[False][0], rtol=5e-3, atol=5e-3)
    assert set(res[True][1]) == set(res[False][1])
    checked = 0
    for n, g in res[False][1].items():
        if g.norm() < 1e-5 * max(1.0, g.numel() ** 0.5) or any(k in n for k in ("key__bias", "encoding__bias", "dw_conv__bias")):
            continue
        g2 = res[True][1][n]
        cos = float((g2 * g).sum() / (g2.norm() * g.norm() + 1e-30))
        assert cos > 0.995, (n, cos)
        checked += 1
    assert checked > 60, checked
