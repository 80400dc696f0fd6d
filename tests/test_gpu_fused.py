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
