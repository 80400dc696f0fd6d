"""GPU parity of the one-kernel FFModule forward (csrc/ffn_fused.cu; conformer.py:101-109) against (a) a plain PyTorch fp32 reference of the
same op on the same bf16-rounded inputs and (b) the three-launch route it replaces (LayerNorm kernel + two tcgen05 GEMMs), whose rounding
points it shares — (b) must agree to a bf16 ulp, dropout masks included (same counter hash, same (row, column) indexing)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(M, H, D=144, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    x = (torch.randn(M, D, device=dev) * 1.3 + 0.2).to(torch.bfloat16)
    g = 1 + 0.2 * torch.randn(D, device=dev)
    b = 0.1 * torch.randn(D, device=dev)
    w1 = (torch.randn(D, H, device=dev) / D ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(H, D, device=dev) / H ** 0.5).to(torch.bfloat16)
    b1, b2 = 0.2 * torch.randn(H, device=dev), 0.2 * torch.randn(D, device=dev)
    return x, g, b, w1, b1, w2, b2


@pytest.mark.parametrize("shape", [(8000, 576), (100, 128), (128, 64), (33000, 576), (1, 576)])
@pytest.mark.parametrize("drop", [0.0, 0.1])
def test_ffn_fwd_matches_unfused_route_and_torch(shape, drop):
    from tensorflowasr_b200 import kernels

    M, H = shape
    x, g, b, w1, b1, w2, b2 = _case(M, H, seed=M + H)
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    w1t, w2t = w1.t().contiguous(), w2.t().contiguous()
    out, xn, z, h, mean, rstd = kernels.ffn_fwd(x, g, b, w1t, b1, w2t, b2, 0.5, drop, seed, 11, 12)
    # (b) the route it replaces
    a_ref, mean_ref, rstd_ref = kernels.layernorm_fwd(x, g, b, 1e-3)
    z_ref = torch.empty((M, H), dtype=torch.bfloat16, device="cuda")
    h_ref = kernels.gemm(a_ref, w1t, bias=b1, activation="swish", aux=z_ref, dropout_p=drop, dropout_seed=seed, dropout_salt=11)
    out_ref = kernels.gemm(h_ref, w2t, bias=b2, residual=x, residual_scale=0.5, dropout_p=drop, dropout_seed=seed, dropout_salt=12)
    torch.testing.assert_close(mean, mean_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rstd, rstd_ref, rtol=1e-4, atol=1e-5)
    for name, got, ref in (("xn", xn, a_ref), ("z", z, z_ref), ("h", h, h_ref), ("out", out, out_ref)):
        gf, rf = got.float(), ref.float()
        # one bf16 ulp of slack (2^-7 relative) wherever an fp32 value sat on a rounding boundary; dropout decisions must be identical
        assert ((gf == 0) == (rf == 0)).float().mean().item() > 0.9999, name
        torch.testing.assert_close(gf, rf, rtol=1.6e-2, atol=1e-2, msg=lambda m: f"{name}: {m}")
        assert (gf != rf).float().mean().item() < 0.02, (name, (gf != rf).float().mean().item())
    # (a) plain fp32 torch
    if drop == 0.0:
        xf = x.float()
        ln = torch.nn.functional.layer_norm(xf, (xf.shape[1],), g, b, 1e-3)
        ref = xf + 0.5 * (torch.nn.functional.silu(ln @ w1.float() + b1) @ w2.float() + b2)
        scale = ref.abs().max().item()
        assert (out.float() - ref).abs().max().item() <= 2e-2 * scale


@pytest.mark.parametrize("shape", [(8000, 576), (100, 128), (1, 64)])
@pytest.mark.parametrize("drop", [0.0, 0.1])
def test_ffn_bwd_chain_matches_unfused_route(shape, drop):
    """dy / dz / da of the one-launch gradient chain vs scale_mask_colsum + the swish' GEMM + the dgrad GEMM (same rounding points)."""
    from tensorflowasr_b200 import kernels

    M, H = shape
    _, _, _, w1, _, w2, _ = _case(M, H, seed=M)
    torch.manual_seed(M + 1)
    dout = torch.randn(M, 144, device="cuda").to(torch.bfloat16)
    z = (1.5 * torch.randn(M, H, device="cuda")).to(torch.bfloat16)
    seed = torch.tensor([777], dtype=torch.int64, device="cuda")
    dy, dz, da = kernels.ffn_bwd(dout, z, w2.contiguous(), w1.contiguous(), 0.5, drop, seed, 21, 22)
    dy_ref = kernels.scale_mask_colsum(dout, scale=0.5, dropout_p=drop, seed=seed, salt=22)
    dz_ref = kernels.gemm(dy_ref, w2.contiguous(), mul=z, mul_mode="swish_grad", dropout_p=drop, dropout_seed=seed, dropout_salt=21)
    da_ref = kernels.gemm(dz_ref, w1.contiguous())
    for name, got, ref in (("dy", dy, dy_ref), ("dz", dz, dz_ref), ("da", da, da_ref)):
        gf, rf = got.float(), ref.float()
        assert ((gf == 0) == (rf == 0)).float().mean().item() > 0.9999, name
        torch.testing.assert_close(gf, rf, rtol=1.6e-2, atol=1e-2 * max(1.0, rf.abs().max().item()), msg=lambda m: f"{name}: {m}")
        assert (gf != rf).float().mean().item() < 0.03, (name, (gf != rf).float().mean().item())


def test_ffn_module_autograd_matches_unfused_module():
    """fused.ff_module with the one-kernel forward vs the same module on the three-launch forward: outputs and every gradient."""
    from tensorflowasr_b200 import fused, ops

    M, H, D = 1000, 576, 144
    x0, g, b, w1, b1, w2, b2 = _case(M, H, seed=5)
    res = {}
    for flag in (False, True):
        ops.set_fused_ffn(flag)
        params = [torch.nn.Parameter(t.detach().float().clone()) for t in (g, b, w1, b1, w2, b2)]
        p1, p2 = fused.WeightPack(params[2], params[3]), fused.WeightPack(params[4], params[5])
        x = x0.clone().requires_grad_(True)
        y = fused.ff_module(x, params[0], params[1], p1, p2, 0.5, 0.0, 3, None)
        y.float().square().sum().backward()
        torch.cuda.synchronize()
        res[flag] = [y.detach().float(), x.grad.float()] + [p.grad.float() for p in params]
    for a, r in zip(res[True], res[False]):
        torch.testing.assert_close(a, r, rtol=2e-2, atol=2e-2 * r.abs().max().item())
