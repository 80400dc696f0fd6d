"""GPU parity: the tcgen05 GEMM (through the C ABI) vs a plain PyTorch fp32 reference of the same op on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, mn=False):
    return (a.float().t() @ b.float()) if mn else (a.float() @ b.float().t())


def _close(got, ref, tol=2e-2):
    scale = ref.abs().max().item() + 1e-6
    err = (got.float() - ref).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("shape", [(128, 64, 64), (8000, 144, 144), (8000, 576, 144), (8000, 144, 576), (8000, 288, 144), (8000, 144, 2880),
                                   (300, 320, 144), (2080, 320, 320), (1, 8, 8), (129, 1000, 320), (4097, 1000, 320), (777, 432, 200)])
def test_tn_gemm_plain(shape):
    from tensorflowasr_b200 import kernels

    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    ref = _ref(a, b)
    out32 = kernels.gemm(a, b, out_dtype=torch.float32)
    torch.cuda.synchronize()
    _close(out32, ref, 2e-3)  # same bf16 inputs, fp32 accumulation: only summation-order noise
    out16 = kernels.gemm(a, b)
    _close(out16, ref, 1e-2)


def test_epilogues():
    from tensorflowasr_b200 import kernels

    torch.manual_seed(0)
    M, N, K = 1000, 288, 144
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    z = _ref(a, b) + bias
    # bias + swish, pre-activation saved
    aux = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out = kernels.gemm(a, b, bias=bias, activation="swish", aux=aux)
    _close(out, z * torch.sigmoid(z))
    _close(aux, z)
    # GLU on interleaved column pairs
    out = kernels.gemm(a, b, bias=bias, glu=True)
    _close(out, z[:, 0::2] * torch.sigmoid(z[:, 1::2]))
    # residual + scale, f32 out
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    out = kernels.gemm(a, b, bias=bias, residual=res, residual_scale=0.5, out_dtype=torch.float32)
    _close(out, res.float() + 0.5 * z, 2e-3)
    # backward epilogues: * swish'(mul), * (1 - mul^2), column sums
    mulz = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    sg = torch.sigmoid(mulz.float())
    colsum = torch.zeros(N, device="cuda")
    out = kernels.gemm(a, b, mul=mulz, mul_mode="swish_grad", out_dtype=torch.float32, colsum=colsum)
    ref = _ref(a, b) * (sg * (1 + mulz.float() * (1 - sg)))
    _close(out, ref, 2e-3)
    _close(colsum, ref.sum(0), 2e-3)
    th = torch.tanh(torch.randn(M, N, device="cuda")).to(torch.bfloat16)
    out = kernels.gemm(a, b, mul=th, mul_mode="tanh_grad", out_dtype=torch.float32)
    _close(out, _ref(a, b) * (1 - th.float() ** 2), 2e-3)
    # dropout: deterministic in (seed, index), right keep rate, kept values scaled
    seed = torch.tensor([123], dtype=torch.int64, device="cuda")
    o1 = kernels.gemm(a, b, dropout_p=0.25, dropout_seed=seed, dropout_salt=7, out_dtype=torch.float32)
    o2 = kernels.gemm(a, b, dropout_p=0.25, dropout_seed=seed, dropout_salt=7, out_dtype=torch.float32)
    seed += 1
    o3 = kernels.gemm(a, b, dropout_p=0.25, dropout_seed=seed, dropout_salt=7, out_dtype=torch.float32)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    kept = o1 != 0
    assert abs(kept.float().mean().item() - 0.75) < 0.01
    _close(o1[kept], (_ref(a, b) / 0.75)[kept], 2e-3)


@pytest.mark.parametrize("shape", [(144, 576, 8000), (576, 144, 8000), (320, 1000, 16640), (144, 144, 333), (2880, 144, 8000), (64, 64, 64)])
def test_mn_major_gemm_for_weight_gradients(shape):
    """dW[K_in, N_out] = X^T . dY with X [rows, K_in], dY [rows, N_out] — no transposes materialised."""
    from tensorflowasr_b200 import kernels

    M, N, K = shape  # output [M, N], contraction K (= number of rows of the sources)
    torch.manual_seed(K)
    x = torch.randn(K, M, device="cuda").to(torch.bfloat16)
    dy = (torch.randn(K, N, device="cuda") / K ** 0.5).to(torch.bfloat16)
    out = kernels.gemm(x, dy, mn_major=True, out_dtype=torch.float32)
    _close(out, _ref(x, dy, mn=True), 2e-3)
    # split-K with atomic accumulation into an existing f32 buffer (how weight gradients land in .grad)
    acc = torch.ones(M, N, device="cuda")
    kernels.gemm(x, dy, mn_major=True, out=acc, accumulate=True, k_splits=8)
    _close(acc - 1.0, _ref(x, dy, mn=True), 2e-3)
