"""CPU stand-ins for `tensorflowasr_b200.kernels` launchers — TEST INFRASTRUCTURE ONLY (tests/test_fused_seq_cpu.py).

Purpose: exercise the autograd wiring of fused_seq.py (argument order, shapes, which gradient lands where, accumulate semantics, pack
refresh) on the CPU.  The seqconv entry points run the REAL kernel source on the CPU emulator (tests/emu) with host pointers; the
tcgen05 GEMM and the BatchNorm kernels — verified on the GPU by tests/test_gpu_gemm.py / test_gpu_fused.py — are replaced by torch CPU
code that follows their documented contracts (bf16 outputs, fp32 accumulation and statistics).
"""
from __future__ import annotations

import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

BF = torch.bfloat16


def _act(z, act):
    if act in (1, "swish"):
        return z * torch.sigmoid(z)
    if act in (2, "relu"):
        return torch.relu(z)
    if act in (3, "tanh"):
        return torch.tanh(z)
    return z


def _act_grad(z, act):
    if act == 1:
        g = torch.sigmoid(z)
        return g * (1 + z * (1 - g))
    if act == 2:
        return (z > 0).float()
    return torch.ones_like(z)


# ------------------------------------------------------------------------------------------------ GEMM / BatchNorm contracts (torch CPU)
def gemm(a, b, *, out=None, out_dtype=None, bias=None, activation=None, aux=None, residual=None, residual_scale=1.0, mul=None, mul_mode=None,
         colsum=None, glu=False, alpha=1.0, dropout_p=0.0, dropout_seed=None, dropout_salt=0, mn_major=False, accumulate=False, k_splits=1):
    assert a.dtype == BF and b.dtype == BF and aux is None and mul is None and colsum is None and not glu and dropout_p == 0.0
    c = (a.float().t() @ b.float()) if mn_major else (a.float() @ b.float().t())
    c = c * alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == c.shape[1]
        c = c + bias
    c = _act(c, activation)
    if residual is not None:
        c = residual.float() + residual_scale * c
    if out is None:
        return c.to(BF if out_dtype is None else out_dtype)
    assert out.shape == c.shape
    if accumulate:
        assert out.dtype == torch.float32
        out.add_(c)
    else:
        out.copy_(c.to(out.dtype))
    return out


BN_COUNT_ON_DEVICE = -2.0


def _with_count(sums, R, with_count):
    return torch.cat([sums.reshape(-1), torch.tensor([float(R)])]) if with_count else sums


def _split_count(sums, count, C):
    if count == BN_COUNT_ON_DEVICE:
        return sums.reshape(-1)[:2 * C].view(2, C), sums.reshape(-1)[2 * C]
    return sums.reshape(-1)[:2 * C].view(2, C), count


def bn_stats(x2d, with_count=False):
    x = x2d.float()
    return _with_count(torch.stack([x.sum(0), (x * x).sum(0)]), x2d.shape[0], with_count)


def bn_finalize(sums, gamma, beta, moving_mean, moving_var, count, eps, momentum):
    sums, count = _split_count(sums, count, gamma.numel())
    mean = sums[0] / count
    var = (sums[1] / count - mean * mean).clamp(min=0)
    rstd = torch.rsqrt(var + eps)
    scale = gamma.float() * rstd
    moving_mean.mul_(momentum).add_(mean, alpha=1 - momentum)
    moving_var.mul_(momentum).add_(var, alpha=1 - momentum)
    return torch.stack([scale, beta.float() - mean * scale, mean, rstd])


def bn_apply(x2d, scale, shift, act=1):
    return _act(x2d.float() * scale + shift, act).to(BF)


def _bn_d(dy, x2d, scale, shift, act):
    return dy.float() * _act_grad(x2d.float() * scale + shift, act)


def bn_bwd_stats(dy, x2d, scale, shift, mean, rstd, act=1, with_count=False):
    d = _bn_d(dy, x2d, scale, shift, act)
    return _with_count(torch.stack([d.sum(0), (d * (x2d.float() - mean) * rstd).sum(0)]), x2d.shape[0], with_count)


def bn_bwd_finalize(sums, dgamma, dbeta):
    sums = sums.reshape(-1)[:2 * dgamma.numel()].view(2, -1)
    dgamma.add_(sums[1])
    dbeta.add_(sums[0])


def bn_bwd_apply(dy, x2d, scale, shift, mean, rstd, sums, count, act=1):
    sums, count = _split_count(sums, count, x2d.shape[1])
    d = _bn_d(dy, x2d, scale, shift, act)
    xh = (x2d.float() - mean) * rstd
    return (scale * (d - sums[0] / count - xh * sums[1] / count)).to(BF)


def scale_mask_colsum(x, scale=1.0, dropout_p=0.0, seed=None, salt=0, colsum=None, add=None, write_out=True):
    assert dropout_p == 0.0
    out = x.float() * scale + (add.float() if add is not None else 0.0)
    if colsum is not None:
        colsum.add_(out.to(BF).float().sum(0))
    return out.to(BF) if write_out else None


def colsum(x2d, out):
    out.add_(x2d.float().sum(0))


def lstm_supported(R):
    return False


def pack_plan_run(plan):
    """PackPlan.run on the CPU: the scatter tfasr_pack_weights performs (csrc/param_ops.cu)."""
    for (prm, off, K, N, sld, dw, ldw, dwt, ldwt, c0, cs, fl) in plan.entries:
        src = prm.detach().reshape(-1)[off // 4:][: (K - 1) * sld + N].float()
        src = torch.stack([src[k * sld: k * sld + N] for k in range(K)])
        cols = c0 + cs * torch.arange(N)
        if dw is not None:
            dw[:K, cols] = src.to(dw.dtype)
        if dwt is not None:
            dwt[cols, :K] = src.t().to(dwt.dtype)


# ------------------------------------------------------------------------------------------------ seqconv: the real kernels on the emulator
_emu = None


def _lib():
    global _emu
    if _emu is None:
        import checks

        _emu = checks.load()
    return _emu


def _p(t):
    if t is None:
        return None
    assert t.is_contiguous() and not t.is_cuda
    return ctypes.c_void_p(t.data_ptr())


def _ok(st, what):
    assert st == 0, f"{what} -> status {st}"


def seq_dwconv_fwd(x, w, stride=1):
    B, T, C = x.shape
    y = torch.empty((B, -(-T // stride), C), dtype=BF)
    _ok(_lib().tfasr_seq_dwconv_fwd(_p(x), _p(w), _p(y), B, T, C, w.shape[0], int(stride), None), "seq_dwconv_fwd")
    return y


def seq_dwconv_bwd(dy, x, w, dw, stride=1, need_dx=True):
    B, T, C = x.shape
    assert dy.shape == (B, -(-T // stride), C) and dw.shape == w.shape and dw.dtype == torch.float32
    dx = torch.empty_like(x) if need_dx else None
    _ok(_lib().tfasr_seq_dwconv_bwd(_p(dy), _p(x), _p(w), _p(dx), _p(dw), B, T, C, w.shape[0], int(stride), None), "seq_dwconv_bwd")
    return dx


def seq_masked_mean(y, length):
    B, T, C = y.shape
    assert length.dtype == torch.int32
    m = torch.empty((B, C), dtype=torch.float32)
    _ok(_lib().tfasr_seq_masked_mean(_p(y), _p(length), _p(m), B, T, C, None), "seq_masked_mean")
    return m


def se_fc_fwd(m, W1, b1, W2, b2, act):
    B, C = m.shape
    H = W1.shape[1]
    z1 = torch.empty((B, H), dtype=torch.float32)
    s = torch.empty((B, C), dtype=torch.float32)
    _ok(_lib().tfasr_se_fc_fwd(_p(m), _p(W1), _p(b1), _p(W2), _p(b2), _p(z1), _p(s), B, C, H, int(act), None), "se_fc_fwd")
    return z1, s


def se_fc_bwd(ds, s, z1, m, W1, W2, length, T, act, dW1, db1, dW2, db2):
    B, C = m.shape
    H = W1.shape[1]
    assert dW1.shape == (C, H) and dW2.shape == (H, C) and db1.shape == (H,) and db2.shape == (C,)
    dz2, dz1, dm = torch.empty_like(s), torch.empty_like(z1), torch.empty_like(s)
    _ok(_lib().tfasr_se_fc_bwd(_p(ds), _p(s), _p(z1), _p(m), _p(W1), _p(W2), _p(length), _p(dz2), _p(dz1), _p(dm), _p(dW1), _p(db1), _p(dW2),
                               _p(db2), B, int(T), C, H, int(act), None), "se_fc_bwd")
    return dm


def se_apply_fwd(y, s, res, act):
    B, T, C = y.shape
    out = torch.empty_like(y)
    _ok(_lib().tfasr_se_apply_fwd(_p(y), _p(s), _p(res), _p(out), B, T, C, int(act), None), "se_apply_fwd")
    return out


def se_apply_bwd(dout, y, s, res, act):
    B, T, C = y.shape
    dres = torch.empty_like(y)
    ds = torch.empty((B, C), dtype=torch.float32)
    _ok(_lib().tfasr_se_apply_bwd(_p(dout), _p(y), _p(s), _p(res), _p(dres), _p(ds), B, T, C, int(act), None), "se_apply_bwd")
    return dres, ds


def se_dy(dres, s, dm, length):
    B, T, C = dres.shape
    dy = torch.empty_like(dres)
    _ok(_lib().tfasr_se_dy(_p(dres), _p(s), _p(dm), _p(length), _p(dy), B, T, C, None), "se_dy")
    return dy


def conv_im2col(x, g):
    import checks

    col = torch.empty((g.B * g.To * g.Fo, g.Kp), dtype=BF)
    gg = checks.ConvGeom(*[getattr(g, n) for n, _ in checks.ConvGeom._fields_])
    _ok(_lib().tfasr_conv_im2col(_p(x), int(x.dtype == torch.float32), _p(col), ctypes.byref(gg), None), "conv_im2col")
    return col


def conv_col2im(dcol, g):
    import checks

    da = torch.empty((g.B, g.T, g.F, g.Cin), dtype=BF)
    gg = checks.ConvGeom(*[getattr(g, n) for n, _ in checks.ConvGeom._fields_])
    _ok(_lib().tfasr_conv_col2im(_p(dcol), _p(da), ctypes.byref(gg), None), "conv_col2im")
    return da


def greedy_decode(weights, e_proj, nframes, prev_tok, states, frame_idx, tok_idx, tokens, iters, cap_iter, blank, mode, max_iters=0, per_frame_cap=1,
                  threads=128):
    """kernels.greedy_decode on the CPU emulator (the real kernel source, host pointers)."""
    import checks

    B, T, J = e_proj.shape
    P = checks.GreedyParams()
    for n in ("emb", "w_ih", "w_hh", "b", "ln_g", "ln_b", "wp", "bp", "wv", "bv"):
        t = weights.get(n)
        assert t is None or t.dtype == torch.float32
        setattr(P, n, _p(t))
    for n, t in (("e_proj", e_proj), ("nframes", nframes), ("prev_tok", prev_tok), ("states", states), ("frame_idx", frame_idx), ("tok_idx", tok_idx),
                 ("tokens", tokens), ("iters", iters), ("cap_iter", cap_iter)):
        assert t.dtype == (torch.float32 if n in ("e_proj", "states") else torch.int32), n
        setattr(P, n, _p(t))
    V, E = weights["emb"].shape
    R = weights["w_hh"].shape[0]
    P.T, P.E, P.R, P.J, P.V, P.max_tokens, P.blank = T, E, R, J, V, tokens.shape[1], int(blank)
    P.mode, P.max_iters, P.per_frame_cap = int(mode), int(max_iters), int(per_frame_cap)
    _ok(_lib().tfasr_greedy_decode_f32(ctypes.byref(P), B, min(int(threads), 128), None), "greedy_decode")


# ------------------------------------------------------------------------------------------------ "ideal" mode: fp32 storage, no bf16 rounding
# Precision-agnostic torch implementations of the seqconv contracts (include/tfasr_b200.h) so that the fused route can be compared with
# the fp32 library route at ~1e-4: any wiring mistake in fused_seq.py is an O(1) difference there.  (That the real kernels honour the
# same contracts is tests/emu/checks.py's job.)
import torch.nn.functional as F


def _dw(x, w, stride):
    K, C = w.shape
    return F.conv1d(F.pad(x.transpose(1, 2), (K - 1, 0)), w.t().unsqueeze(1), None, stride=stride, groups=C).transpose(1, 2)


def ideal_seq_dwconv_fwd(x, w, stride=1):
    return _dw(x.float(), w.float(), stride).contiguous()


def ideal_seq_dwconv_bwd(dy, x, w, dw, stride=1, need_dx=True):
    with torch.enable_grad():
        xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
        gx, gw = torch.autograd.grad(_dw(xr, wr, stride), (xr, wr), dy.float())
    dw.add_(gw)
    return gx.contiguous() if need_dx else None


def _len(length, T):
    return length.to(torch.int64).clamp(0, T)


def ideal_seq_masked_mean(y, length):
    B, T, C = y.shape
    n = _len(length, T)
    mask = (torch.arange(T)[None, :] < n[:, None]).float()[..., None]
    return (y.float() * mask).sum(1) / n.clamp(min=1)[:, None].float()


def ideal_se_fc_fwd(m, W1, b1, W2, b2, act):
    z1 = m @ W1 + b1
    return z1, torch.sigmoid(_act(z1, act) @ W2 + b2)


def ideal_se_fc_bwd(ds, s, z1, m, W1, W2, length, T, act, dW1, db1, dW2, db2):
    dz2 = ds * s * (1 - s)
    dz1 = (dz2 @ W2.t()) * _act_grad(z1, act)
    dW1.add_(m.t() @ dz1)
    db1.add_(dz1.sum(0))
    dW2.add_(_act(z1, act).t() @ dz2)
    db2.add_(dz2.sum(0))
    return (dz1 @ W1.t()) / _len(length, T).clamp(min=1)[:, None].float()


def ideal_se_apply_fwd(y, s, res, act):
    return _act(y.float() * s[:, None, :] + (res.float() if res is not None else 0.0), act)


def ideal_se_apply_bwd(dout, y, s, res, act):
    z = y.float() * s[:, None, :] + (res.float() if res is not None else 0.0)
    dpre = dout.float() * _act_grad(z, act)
    return dpre, (dpre * y.float()).sum(1)


def ideal_se_dy(dres, s, dm, length):
    B, T, C = dres.shape
    valid = (torch.arange(T)[None, :] < length[:, None]).float()[..., None]
    return dres.float() * s[:, None, :] + valid * dm[:, None, :]


def _im2col(x, g):
    xp = F.pad(x.float(), (0, 0, g.pl, g.kw + g.sw * g.Fo, g.pt, g.kh + g.sh * g.To))  # generous zero border after the data
    cols = []
    for i in range(g.kh):
        for j in range(g.kw):
            cols.append(xp[:, i: i + g.sh * g.To: g.sh, j: j + g.sw * g.Fo: g.sw, :])
    col = torch.cat(cols, -1).reshape(g.B * g.To * g.Fo, g.kh * g.kw * g.Cin)
    return F.pad(col, (0, g.Kp - col.shape[1])).contiguous()


def ideal_conv_im2col(x, g):
    return _im2col(x, g)


def ideal_conv_col2im(dcol, g):
    with torch.enable_grad():
        xr = torch.zeros((g.B, g.T, g.F, g.Cin), requires_grad=True)
        (gx,) = torch.autograd.grad(_im2col(xr, g), xr, dcol.float())
    return gx


class _TorchProxy:
    """`torch` as seen by the fused modules in ideal mode: bfloat16 reads as float32, everything else is torch."""

    def __getattr__(self, name):
        return torch.float32 if name == "bfloat16" else getattr(torch, name)


PATCHED = ["gemm", "bn_stats", "bn_finalize", "bn_apply", "bn_bwd_stats", "bn_bwd_finalize", "bn_bwd_apply", "scale_mask_colsum", "colsum",
           "lstm_supported", "seq_dwconv_fwd", "seq_dwconv_bwd", "seq_masked_mean", "se_fc_fwd", "se_fc_bwd", "se_apply_fwd", "se_apply_bwd", "se_dy",
           "conv_im2col", "conv_col2im", "greedy_decode"]


def install(monkeypatch, ideal=False):
    """Route tensorflowasr_b200.kernels / PackPlan / the fused switches to the CPU stand-ins for the duration of one test.
    ideal=True: fp32 storage everywhere (no bf16 rounding) and torch implementations of the seqconv contracts."""
    from tensorflowasr_b200 import ctc_models, fused, fused_seq, kernels, ops

    me = sys.modules[__name__]
    for n in PATCHED:
        monkeypatch.setattr(kernels, n, getattr(me, ("ideal_" + n) if (ideal and hasattr(me, "ideal_" + n)) else n))
    if ideal:
        monkeypatch.setattr(me, "BF", torch.float32)
        for mod in (fused, fused_seq, ctc_models):
            monkeypatch.setattr(mod, "torch", _TorchProxy())
        monkeypatch.setattr(ops, "compute_dtype", lambda: torch.float32)  # the cuDNN-LSTM glue of the DeepSpeech2 route casts through it
    monkeypatch.setattr(fused.PackPlan, "run", pack_plan_run)
    monkeypatch.setattr(ops, "fused_active", lambda x: ops._FUSED and ops.precision() == "bf16")
