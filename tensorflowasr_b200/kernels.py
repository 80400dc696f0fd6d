"""Thin Python launchers over the C ABI (include/tfasr_b200.h) for the dense / fused encoder kernels."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_ACT = {None: 0, "none": 0, "swish": 1, "relu": 2, "tanh": 3}
_MUL = {None: 0, "swish_grad": 1, "tanh_grad": 2, "mul": 3}


def gemm(a, b, *, out=None, out_dtype=torch.bfloat16, bias=None, activation=None, aux=None, residual=None, residual_scale=1.0,
         mul=None, mul_mode=None, colsum=None, glu=False, alpha=1.0, dropout_p=0.0, dropout_seed=None, dropout_salt=0, mn_major=False,
         accumulate=False, k_splits=1, rowstats=None):
    """C = epilogue(alpha * A . B^T) on the tcgen05 GEMM.
    rowstats: optional f32 [M, 2 * ceil(N / 256), 2] receiving per-row, per-column-slice (max, sum exp(v - max)) of the output.

    default: a [M,K], b [N,K] (bf16, row-major).  mn_major=True: a [K,M], b [K,N] -> C [M,N] = a^T . b.
    """
    lib = _lib.load()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if mn_major:
        K, M = a.shape
        K2, N = b.shape
    else:
        M, K = a.shape
        N, K2 = b.shape
    assert K == K2, (a.shape, b.shape)
    n_out = N // 2 if glu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, n_out)
    e = _lib.GemmEpilogue()
    e.bias = _lib.ptr(bias) if bias is not None else None
    e.aux = aux.data_ptr() if aux is not None else None
    e.ld_aux = aux.stride(0) if aux is not None else 0
    e.residual = residual.data_ptr() if residual is not None else None
    e.ld_residual = residual.stride(0) if residual is not None else 0
    e.mul = mul.data_ptr() if mul is not None else None
    e.ld_mul = mul.stride(0) if mul is not None else 0
    e.colsum = colsum.data_ptr() if colsum is not None else None
    e.c_f32 = int(out.dtype == torch.float32)
    e.activation = _ACT[activation]
    e.glu = int(bool(glu))
    e.mul_mode = _MUL[mul_mode]
    e.residual_scale = float(residual_scale)
    e.alpha = float(alpha)
    e.dropout_p = float(dropout_p)
    e.dropout_seed_ptr = dropout_seed.data_ptr() if dropout_seed is not None else None  # int64 device scalar
    e.dropout_salt = int(dropout_salt) & 0xFFFFFFFFFFFFFFFF
    e.atomic_out = int(bool(accumulate))
    e.k_splits = int(k_splits)
    if rowstats is not None:
        assert rowstats.dtype == torch.float32 and rowstats.is_contiguous() and rowstats.shape == (M, 2 * ((N + 255) // 256), 2)
        e.rowstats = rowstats.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    _lib.check(
        lib.tfasr_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                            int(mn_major), int(mn_major), ctypes.byref(e), _lib.stream_ptr(a.device)),
        "tfasr_gemm_bf16",
    )
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-3):
    """x [M, D] bf16|f32 -> (y bf16 [M, D], mean f32 [M], rstd f32 [M])."""
    lib = _lib.load()
    M, D = x.shape
    y = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    mean = torch.empty(M, dtype=torch.float32, device=x.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    _lib.check(lib.tfasr_layernorm_fwd(_lib.ptr(x), int(x.dtype == torch.float32), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(y), _lib.ptr(mean),
                                       _lib.ptr(rstd), M, D, float(eps), _lib.stream_ptr(x.device)), "tfasr_layernorm_fwd")
    return y, mean, rstd


def ffn_supported(D: int, H: int) -> bool:
    return bool(_lib.load().tfasr_ffn_supported(int(D), int(H)))


def ffn_fwd(x, ln_g, ln_b, w1t, b1, w2t, b2, factor, drop_p=0.0, seed=None, salt1=0, salt2=0, eps=1e-3):
    """The whole FFModule forward in one launch (csrc/ffn_fused.cu).  x [M, D] bf16, w1t [H, D], w2t [D, H] bf16.
    -> (out [M, D], xn [M, D], z [M, H], h [M, H] bf16, mean [M], rstd [M] f32)."""
    lib = _lib.load()
    M, D = x.shape
    H = w1t.shape[0]
    dev = x.device
    out = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
    xn = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
    z = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
    h = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
    mean = torch.empty(M, dtype=torch.float32, device=dev)
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    _lib.check(lib.tfasr_ffn_fwd_bf16(_lib.ptr(x), _lib.ptr(ln_g), _lib.ptr(ln_b), _lib.ptr(w1t), _lib.ptr(b1), _lib.ptr(w2t), _lib.ptr(b2), _lib.ptr(out),
                                      _lib.ptr(xn), _lib.ptr(z), _lib.ptr(h), _lib.ptr(mean), _lib.ptr(rstd), M, D, H, float(eps), float(factor),
                                      float(drop_p), _lib.ptr(seed) if drop_p > 0 else None, int(salt1) & (2 ** 64 - 1), int(salt2) & (2 ** 64 - 1),
                                      _lib.stream_ptr(dev)), "tfasr_ffn_fwd_bf16")
    return out, xn, z, h, mean, rstd


def ffn_bwd(dout, z, w2, w1, factor, drop_p=0.0, seed=None, salt1=0, salt2=0):
    """FFModule's gradient chain in one launch.  dout [M, D], z [M, H] (saved), w2 = W2 [H, D], w1 = W1 [D, H] bf16.
    -> (dy [M, D] = factor * Drop2'(dout), dz [M, H], da [M, D] = gradient of LN(x)), all bf16."""
    lib = _lib.load()
    M, D = dout.shape
    H = z.shape[1]
    dy = torch.empty_like(dout)
    dz = torch.empty_like(z)
    da = torch.empty_like(dout)
    _lib.check(lib.tfasr_ffn_bwd_bf16(_lib.ptr(dout), _lib.ptr(z), _lib.ptr(w2), _lib.ptr(w1), _lib.ptr(dy), _lib.ptr(dz), _lib.ptr(da), M, D, H, float(factor),
                                      float(drop_p), _lib.ptr(seed) if drop_p > 0 else None, int(salt1) & (2 ** 64 - 1), int(salt2) & (2 ** 64 - 1),
                                      _lib.stream_ptr(dout.device)), "tfasr_ffn_bwd_bf16")
    return dy, dz, da


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None):
    """-> dx bf16 [M, D] (= LN backward of dy, + dres); dgamma / dbeta (f32 [D]) are accumulated in place."""
    lib = _lib.load()
    M, D = x.shape
    dx = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.tfasr_layernorm_bwd(_lib.ptr(dy), _lib.ptr(x), int(x.dtype == torch.float32), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                                       _lib.ptr(dres), _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta), M, D, _lib.stream_ptr(x.device)),
               "tfasr_layernorm_bwd")
    return dx


def scale_mask_colsum(x, scale=1.0, dropout_p=0.0, seed=None, salt=0, colsum=None, add=None, write_out=True):
    """out = scale * dropout_mask * x (+ add), bf16 [M, N]; colsum (f32 [N]) += column sums of out."""
    lib = _lib.load()
    M, N = x.shape
    out = torch.empty_like(x) if write_out else None
    _lib.check(lib.tfasr_scale_mask_colsum(_lib.ptr(x), _lib.ptr(add), _lib.ptr(out), _lib.ptr(colsum), M, N, float(scale), float(dropout_p),
                                           seed.data_ptr() if seed is not None else None, int(salt) & 0xFFFFFFFFFFFFFFFF,
                                           _lib.stream_ptr(x.device)), "tfasr_scale_mask_colsum")
    return out


def glu_bwd(aux, dg, colsum=None):
    """aux [M, 2D] interleaved pre-GLU (a, b) pairs, dg [M, D] -> daux [M, 2D]; colsum (f32 [2D]) += column sums."""
    lib = _lib.load()
    M, D = dg.shape
    daux = torch.empty_like(aux)
    _lib.check(lib.tfasr_glu_bwd(_lib.ptr(aux), _lib.ptr(dg), _lib.ptr(daux), _lib.ptr(colsum), M, D, _lib.stream_ptr(dg.device)), "tfasr_glu_bwd")
    return daux


def _joint_v2(U1, J):
    from . import ops

    return ops.joint_v2_active() and J % 8 == 0 and J <= 4096 and U1 <= 72


def joint_tanh_fwd(e, p):
    """e [B,T,J], p [B,U1,J] (bf16) -> h [B,T,U1,J] bf16 = tanh(e[:, :, None] + p[:, None])."""
    lib = _lib.load()
    B, T, J = e.shape
    U1 = p.shape[1]
    h = torch.empty((B, T, U1, J), dtype=torch.bfloat16, device=e.device)
    fn, name = (lib.tfasr_joint_tanh_fwd_v2, "tfasr_joint_tanh_fwd_v2") if _joint_v2(U1, J) else (lib.tfasr_joint_tanh_fwd, "tfasr_joint_tanh_fwd")
    _lib.check(fn(_lib.ptr(e), _lib.ptr(p), _lib.ptr(h), B, T, U1, J, _lib.stream_ptr(e.device)), name)
    return h


def joint_reduce_bwd(dz):
    """dz [B,T,U1,J] bf16 -> (de [B,T,J] f32 = sum over u, dp [B,U1,J] f32 = sum over t)."""
    lib = _lib.load()
    B, T, U1, J = dz.shape
    de = torch.empty((B, T, J), dtype=torch.float32, device=dz.device)
    dp = torch.zeros((B, U1, J), dtype=torch.float32, device=dz.device)
    from . import ops

    fn, name = (lib.tfasr_joint_reduce_bwd_v2, "tfasr_joint_reduce_bwd_v2") if (_joint_v2(U1, J) and ops.joint_reduce_v2_active()) else (lib.tfasr_joint_reduce_bwd, "tfasr_joint_reduce_bwd")
    _lib.check(fn(_lib.ptr(dz), _lib.ptr(de), _lib.ptr(dp), B, T, U1, J, _lib.stream_ptr(dz.device)), name)
    return de, dp


# ------------------------------------------------------------------------------------------------ conv subsampling / batch norm
def _st(t):
    return _lib.stream_ptr(t.device)


def _conv1_v2(C):
    from . import ops

    return ops.conv1_v2_active() and C % 8 == 0 and C <= 256


def conv1_fwd(x, w, bias):
    """x [B,T,F] f32, w [3,3,1,C] f32 -> y [B,T1,F1,C] bf16 (3x3, stride 2, causal)."""
    lib = _lib.load()
    B, T, F = x.shape
    C = w.shape[-1]
    y = torch.empty((B, (T + 1) // 2, (F + 1) // 2, C), dtype=torch.bfloat16, device=x.device)
    fn, name = (lib.tfasr_conv1_3x3s2_fwd_v2, "tfasr_conv1_3x3s2_fwd_v2") if _conv1_v2(C) else (lib.tfasr_conv1_3x3s2_fwd, "tfasr_conv1_3x3s2_fwd")
    _lib.check(fn(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, T, F, C, _st(x)), name)
    return y


def conv1_wgrad(x, dy, dw):
    lib = _lib.load()
    B, T, F = x.shape
    C = dy.shape[-1]
    fn, name = (lib.tfasr_conv1_3x3s2_wgrad_v2, "tfasr_conv1_3x3s2_wgrad_v2") if _conv1_v2(C) else (lib.tfasr_conv1_3x3s2_wgrad, "tfasr_conv1_3x3s2_wgrad")
    _lib.check(fn(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), B, T, F, C, _st(x)), name)


def _im2col_v2(C):
    from . import ops

    return ops.im2col_v2_active() and C % 8 == 0 and C <= 4096


def im2col(a):
    """a [B,T1,F1,C] bf16 -> col [B*T2*F2, 9C]."""
    lib = _lib.load()
    B, T1, F1, C = a.shape
    T2, F2 = (T1 + 1) // 2, (F1 + 1) // 2
    col = torch.empty((B * T2 * F2, 9 * C), dtype=torch.bfloat16, device=a.device)
    fn, name = (lib.tfasr_im2col_3x3s2_v2, "tfasr_im2col_3x3s2_v2") if _im2col_v2(C) else (lib.tfasr_im2col_3x3s2, "tfasr_im2col_3x3s2")
    _lib.check(fn(_lib.ptr(a), _lib.ptr(col), B, T1, F1, C, _st(a)), name)
    return col


def col2im(dcol, shape):
    lib = _lib.load()
    B, T1, F1, C = shape
    da = torch.empty(shape, dtype=torch.bfloat16, device=dcol.device)
    fn, name = (lib.tfasr_col2im_3x3s2_v2, "tfasr_col2im_3x3s2_v2") if _im2col_v2(C) else (lib.tfasr_col2im_3x3s2, "tfasr_col2im_3x3s2")
    _lib.check(fn(_lib.ptr(dcol), _lib.ptr(da), B, T1, F1, C, _st(dcol)), name)
    return da


BN_COUNT_ON_DEVICE = -2.0  # `count` sentinel of tfasr_bn_finalize / tfasr_bn_bwd_apply: the row count is the float at sums[2 * C]

_BN_TEMPLATES = {}


def _bn_sums(C, R, device, with_count):
    """Zeroed accumulator [2, C]; with_count: flat [2 C + 1] whose last slot holds this replica's row count R, so that the all-reduce of
    the statistics also delivers the global count (replicas may pad to different lengths).  One copy launch instead of one fill launch."""
    if not with_count:
        return torch.zeros((2, C), dtype=torch.float32, device=device)
    key = (C, int(R), str(device))
    t = _BN_TEMPLATES.get(key)
    if t is None:
        t = torch.zeros(2 * C + 1, dtype=torch.float32)
        t[2 * C] = float(R)
        t = _BN_TEMPLATES[key] = t.to(device)
    return t.clone()


def bn_stats(x2d, with_count=False):
    lib = _lib.load()
    R, C = x2d.shape
    sums = _bn_sums(C, R, x2d.device, with_count)
    _lib.check(lib.tfasr_bn_stats(_lib.ptr(x2d), _lib.ptr(sums), R, C, _st(x2d)), "tfasr_bn_stats")
    return sums


def bn_finalize(sums, gamma, beta, moving_mean, moving_var, count, eps, momentum):
    """-> out [4, C] f32 = (scale, shift, mean, rstd); moving statistics updated in place."""
    lib = _lib.load()
    C = gamma.numel()
    out = torch.empty((4, C), dtype=torch.float32, device=sums.device)
    _lib.check(lib.tfasr_bn_finalize(_lib.ptr(sums), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(moving_mean), _lib.ptr(moving_var), _lib.ptr(out), C,
                                     float(count), float(eps), float(momentum), _st(sums)), "tfasr_bn_finalize")
    return out


def bn_bwd_finalize(sums, dgamma, dbeta):
    lib = _lib.load()
    _lib.check(lib.tfasr_bn_bwd_finalize(_lib.ptr(sums), _lib.ptr(dgamma), _lib.ptr(dbeta), dgamma.numel(), _st(sums)), "tfasr_bn_bwd_finalize")


def bn_apply(x2d, scale, shift, act=1):
    lib = _lib.load()
    R, C = x2d.shape
    y = torch.empty_like(x2d)
    _lib.check(lib.tfasr_bn_apply(_lib.ptr(x2d), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(y), R, C, int(act), _st(x2d)), "tfasr_bn_apply")
    return y


def bn_bwd_stats(dy, x2d, scale, shift, mean, rstd, act=1, with_count=False):
    lib = _lib.load()
    R, C = x2d.shape
    sums = _bn_sums(C, R, x2d.device, with_count)
    _lib.check(lib.tfasr_bn_bwd_stats(_lib.ptr(dy), _lib.ptr(x2d), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(sums),
                                      R, C, int(act), _st(x2d)), "tfasr_bn_bwd_stats")
    return sums


def bn_bwd_apply(dy, x2d, scale, shift, mean, rstd, sums, count, act=1):
    lib = _lib.load()
    R, C = x2d.shape
    dx = torch.empty_like(x2d)
    _lib.check(lib.tfasr_bn_bwd_apply(_lib.ptr(dy), _lib.ptr(x2d), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(sums),
                                      float(count), _lib.ptr(dx), R, C, int(act), _st(x2d)), "tfasr_bn_bwd_apply")
    return dx


def colsum(x2d, out):
    """out (f32 [N]) += column sums of x2d [M, N] bf16."""
    scale_mask_colsum(x2d, colsum=out, write_out=False)


def rel_attention_fwd(qkv, pos, ubias, vbias, length, B, T, N, H, chunk=None, history=None):
    """-> (ctx [B*T, N*H] bf16, probs [B,N,T,T] bf16).  `pos` [B*R, N*H]: contiguous, or a column slice of the all-blocks projection
    (row stride = pos.stride(0))."""
    lib = _lib.load()
    ctx = torch.empty((B * T, N * H), dtype=torch.bfloat16, device=qkv.device)
    probs = torch.empty((B, N, T, T), dtype=torch.bfloat16, device=qkv.device)
    ch = int(chunk) if (chunk is not None and history is not None) else 0
    hi = int(history) if (chunk is not None and history is not None) else 0
    from . import ops

    if ops.attention_tc_active(T) and pos.dim() == 2 and pos.stride(1) == 1 and lib.tfasr_rel_attention_tc_supported(T, N, H, qkv.stride(0), pos.stride(0)):
        # forward on tcgen05 (csrc/attention_tc.cu): any T, contiguous pos or a column slice of the all-blocks projection
        _lib.check(lib.tfasr_rel_attention_fwd_tc(_lib.ptr(qkv), qkv.stride(0), _lib.ptr_rows(pos), pos.stride(0), _lib.ptr(ubias), _lib.ptr(vbias),
                                                  _lib.ptr(length), B, T, N, H, ch, hi, _lib.ptr(ctx), _lib.ptr(probs), _st(qkv)),
                   "tfasr_rel_attention_fwd_tc")
        return ctx, probs
    if pos.dim() == 2 and pos.stride(0) != N * H:
        assert pos.stride(1) == 1
        _lib.check(lib.tfasr_rel_attention_fwd_ld(_lib.ptr(qkv), qkv.stride(0), _lib.ptr_rows(pos), pos.stride(0), _lib.ptr(ubias), _lib.ptr(vbias),
                                                  _lib.ptr(length), B, T, N, H, ch, hi, _lib.ptr(ctx), _lib.ptr(probs), _st(qkv)),
                   "tfasr_rel_attention_fwd_ld")
        return ctx, probs
    _lib.check(lib.tfasr_rel_attention_fwd(_lib.ptr(qkv), qkv.stride(0), _lib.ptr(pos), _lib.ptr(ubias), _lib.ptr(vbias), _lib.ptr(length), B, T, N, H,
                                           ch, hi, _lib.ptr(ctx), _lib.ptr(probs), _st(qkv)), "tfasr_rel_attention_fwd")
    return ctx, probs


def rel_attention_bwd(qkv, pos, ubias, vbias, length, B, T, N, H, probs, dctx, du, dvb, chunk=None, history=None, ctx=None, dpos_out=None):
    """-> (dqkv like qkv, dpos like pos); du / dvb (f32 [N,H]) accumulated.  `ctx`: the forward's output (tensor-core path).
    `dpos_out`: where to write dpos ([B*R, N*H] view with unit column stride, e.g. this block's slice of the shared gradient)."""
    lib = _lib.load()
    dqkv = torch.empty_like(qkv)
    ds = torch.empty_like(probs)
    ch = int(chunk) if (chunk is not None and history is not None) else 0
    hi = int(history) if (chunk is not None and history is not None) else 0
    if dpos_out is not None or (pos.dim() == 2 and pos.stride(0) != N * H):
        dpos = dpos_out if dpos_out is not None else torch.empty((pos.shape[0], N * H), dtype=torch.bfloat16, device=pos.device)
        assert pos.stride(-1) == 1 and dpos.stride(-1) == 1
        _lib.check(lib.tfasr_rel_attention_bwd_ld(_lib.ptr(qkv), qkv.stride(0), _lib.ptr_rows(pos), pos.stride(0), _lib.ptr(ubias), _lib.ptr(vbias),
                                                  _lib.ptr(length), B, T, N, H, ch, hi, _lib.ptr(probs), _lib.ptr(dctx), _lib.ptr(ctx), _lib.ptr(ds),
                                                  _lib.ptr(dqkv), _lib.ptr_rows(dpos), dpos.stride(0), _lib.ptr(du), _lib.ptr(dvb), _st(qkv)),
                   "tfasr_rel_attention_bwd_ld")
        return dqkv, dpos
    dpos = torch.empty_like(pos)
    _lib.check(lib.tfasr_rel_attention_bwd(_lib.ptr(qkv), qkv.stride(0), _lib.ptr(pos), _lib.ptr(ubias), _lib.ptr(vbias), _lib.ptr(length), B, T, N, H,
                                           ch, hi, _lib.ptr(probs), _lib.ptr(dctx), _lib.ptr(ctx), _lib.ptr(ds), _lib.ptr(dqkv), _lib.ptr(dpos), _lib.ptr(du),
                                           _lib.ptr(dvb), _st(qkv)), "tfasr_rel_attention_bwd")
    return dqkv, dpos


def dwconv_fwd(x, w, bias):
    """x [B,T,C] bf16, w [K,C] f32 -> y [B,T,C] bf16 (causal depthwise conv)."""
    lib = _lib.load()
    B, T, C = x.shape
    y = torch.empty_like(x)
    _lib.check(lib.tfasr_dwconv_causal_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, T, C, w.shape[0], _st(x)), "tfasr_dwconv_causal_fwd")
    return y


def dwconv_bwd(dy, x, w, dw, dbias=None):
    """-> dx [B,T,C] bf16; dw (f32 [K,C]) and dbias (f32 [C] or None) accumulated."""
    lib = _lib.load()
    B, T, C = x.shape
    dx = torch.empty_like(x)
    from . import ops

    if ops.dwconv_wgrad_v2_active():
        _lib.check(lib.tfasr_dwconv_causal_bwd_dx(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), B, T, C, w.shape[0], _st(x)), "tfasr_dwconv_causal_bwd_dx")
        _lib.check(lib.tfasr_dwconv_causal_wgrad_v2(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(dbias), B, T, C, w.shape[0], _st(x)),
                   "tfasr_dwconv_causal_wgrad_v2")
        return dx
    _lib.check(lib.tfasr_dwconv_causal_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(w), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(dbias), B, T, C, w.shape[0],
                                           _st(x)), "tfasr_dwconv_causal_bwd")
    return dx


def lstm_supported(R: int) -> bool:
    return bool(_lib.load().tfasr_lstm_supported(int(R)))


def lstm_rec_fwd(gx, whh, length=None, h0=None, c0=None, want_state=False):
    """gx [B,U,4R] f32, whh [R,4R] f32 -> (y [B,U,R] bf16, saved = (gates [B,U,4R] bf16, c [B,U,R] f32), (hT, cT) or None)."""
    lib = _lib.load()
    B, U, G = gx.shape
    R = G // 4
    y = torch.empty((B, U, R), dtype=torch.bfloat16, device=gx.device)
    saved_g = torch.empty((B, U, 4 * R), dtype=torch.bfloat16, device=gx.device)
    saved_c = torch.empty((B, U, R), dtype=torch.float32, device=gx.device)
    hT = torch.empty((B, R), dtype=torch.float32, device=gx.device) if want_state else None
    cT = torch.empty((B, R), dtype=torch.float32, device=gx.device) if want_state else None
    _lib.check(lib.tfasr_lstm_rec_fwd(_lib.ptr(gx), _lib.ptr(whh), _lib.ptr(h0), _lib.ptr(c0), _lib.ptr(length), _lib.ptr(y), _lib.ptr(saved_g),
                                      _lib.ptr(saved_c), _lib.ptr(hT), _lib.ptr(cT), B, U, R, _st(gx)), "tfasr_lstm_rec_fwd")
    return y, (saved_g, saved_c), ((hT, cT) if want_state else None)


def lstm_rec_bwd(dy, saved, whh, length=None, c0=None):
    """dy [B,U,R] bf16 -> dgates [B,U,4R] bf16 (pre-activation gradients)."""
    lib = _lib.load()
    B, U, R = dy.shape
    dgates = torch.empty((B, U, 4 * R), dtype=torch.bfloat16, device=dy.device)
    _lib.check(lib.tfasr_lstm_rec_bwd(_lib.ptr(dy), _lib.ptr(saved[0]), _lib.ptr(saved[1]), _lib.ptr(c0), _lib.ptr(whh), _lib.ptr(length), _lib.ptr(dgates), None, None,
                                      B, U, R, _st(dy)), "tfasr_lstm_rec_bwd")
    return dgates


# ------------------------------------------------------------------------------------------------ ContextNet / DeepSpeech2 encoder bodies
_BN_ACT = {None: 0, "linear": 0, "none": 0, "swish": 1, "silu": 1, "relu": 2}


def act_code(name) -> int:
    return _BN_ACT[name]


def seq_dwconv_fwd(x, w, stride=1):
    """x [B,T,C] bf16, w [K,C] f32 -> y [B, ceil(T/stride), C] bf16 (causal depthwise conv, contextnet.py:53-63)."""
    lib = _lib.load()
    B, T, C = x.shape
    y = torch.empty((B, -(-T // stride), C), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.tfasr_seq_dwconv_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), B, T, C, w.shape[0], int(stride), _st(x)), "tfasr_seq_dwconv_fwd")
    return y


def seq_dwconv_bwd(dy, x, w, dw, stride=1, need_dx=True):
    """-> dx [B,T,C] bf16 (or None); dw (f32 [K,C]) accumulated."""
    lib = _lib.load()
    B, T, C = x.shape
    dx = torch.empty_like(x) if need_dx else None
    _lib.check(lib.tfasr_seq_dwconv_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(w), _lib.ptr(dx), _lib.ptr(dw), B, T, C, w.shape[0], int(stride), _st(x)),
               "tfasr_seq_dwconv_bwd")
    return dx


def seq_masked_mean(y, length):
    """y [B,T,C] bf16, length [B] int32 -> m [B,C] f32 (mean over the valid frames)."""
    lib = _lib.load()
    B, T, C = y.shape
    m = torch.empty((B, C), dtype=torch.float32, device=y.device)
    _lib.check(lib.tfasr_seq_masked_mean(_lib.ptr(y), _lib.ptr(length), _lib.ptr(m), B, T, C, _st(y)), "tfasr_seq_masked_mean")
    return m


def se_fc_fwd(m, W1, b1, W2, b2, act):
    """-> (z1 [B,H] f32, s [B,C] f32 = sigmoid(act(m W1 + b1) W2 + b2))."""
    lib = _lib.load()
    B, C = m.shape
    H = W1.shape[1]
    z1 = torch.empty((B, H), dtype=torch.float32, device=m.device)
    s = torch.empty((B, C), dtype=torch.float32, device=m.device)
    _lib.check(lib.tfasr_se_fc_fwd(_lib.ptr(m), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(z1), _lib.ptr(s), B, C, H, int(act),
                                   _st(m)), "tfasr_se_fc_fwd")
    return z1, s


def se_fc_bwd(ds, s, z1, m, W1, W2, length, T, act, dW1, db1, dW2, db2):
    """-> dm [B,C] f32 (already divided by the utterance length); parameter gradients accumulated."""
    lib = _lib.load()
    B, C = m.shape
    H = W1.shape[1]
    dz2 = torch.empty_like(s)
    dz1 = torch.empty_like(z1)
    dm = torch.empty_like(s)
    _lib.check(lib.tfasr_se_fc_bwd(_lib.ptr(ds), _lib.ptr(s), _lib.ptr(z1), _lib.ptr(m), _lib.ptr(W1), _lib.ptr(W2), _lib.ptr(length), _lib.ptr(dz2),
                                   _lib.ptr(dz1), _lib.ptr(dm), _lib.ptr(dW1), _lib.ptr(db1), _lib.ptr(dW2), _lib.ptr(db2), B, int(T), C, H, int(act),
                                   _st(m)), "tfasr_se_fc_bwd")
    return dm


def se_apply_fwd(y, s, res, act):
    lib = _lib.load()
    B, T, C = y.shape
    out = torch.empty_like(y)
    _lib.check(lib.tfasr_se_apply_fwd(_lib.ptr(y), _lib.ptr(s), _lib.ptr(res), _lib.ptr(out), B, T, C, int(act), _st(y)), "tfasr_se_apply_fwd")
    return out


def se_apply_bwd(dout, y, s, res, act):
    """-> (dres [B,T,C] bf16, ds [B,C] f32)."""
    lib = _lib.load()
    B, T, C = y.shape
    dres = torch.empty_like(y)
    ds = torch.empty((B, C), dtype=torch.float32, device=y.device)
    _lib.check(lib.tfasr_se_apply_bwd(_lib.ptr(dout), _lib.ptr(y), _lib.ptr(s), _lib.ptr(res), _lib.ptr(dres), _lib.ptr(ds), B, T, C, int(act), _st(y)),
               "tfasr_se_apply_bwd")
    return dres, ds


def se_dy(dres, s, dm, length):
    lib = _lib.load()
    B, T, C = dres.shape
    dy = torch.empty_like(dres)
    _lib.check(lib.tfasr_se_dy(_lib.ptr(dres), _lib.ptr(s), _lib.ptr(dm), _lib.ptr(length), _lib.ptr(dy), B, T, C, _st(dres)), "tfasr_se_dy")
    return dy


def conv_geom(B, T, F, Cin, kh, kw, sh, sw, padding):
    """Geometry of Conv2D(kernel (kh,kw), strides (sh,sw), padding) on [B,T,F,Cin]: keras "same" (total = max((ceil(n/s)-1) s + k - n, 0),
    total // 2 in front) or the reference's causal padding (k - 1 in front, then VALID; convolution.py:25-37)."""
    if padding == "same":
        To, Fo = -(-T // sh), -(-F // sw)
        pt = max((To - 1) * sh + kh - T, 0) // 2
        pl = max((Fo - 1) * sw + kw - F, 0) // 2
    elif padding == "causal":
        pt, pl = kh - 1, kw - 1
        To, Fo = (T + pt - kh) // sh + 1, (F + pl - kw) // sw + 1
    else:
        raise ValueError(padding)
    return _lib.ConvGeom(B, T, F, Cin, To, Fo, kh, kw, sh, sw, pt, pl, (kh * kw * Cin + 7) // 8 * 8)


def conv_im2col(x, g):
    """x [B,T,F,Cin] f32 | bf16 -> col [B*To*Fo, Kp] bf16."""
    lib = _lib.load()
    col = torch.empty((g.B * g.To * g.Fo, g.Kp), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.tfasr_conv_im2col(_lib.ptr(x), int(x.dtype == torch.float32), _lib.ptr(col), ctypes.byref(g), _st(x)), "tfasr_conv_im2col")
    return col


def conv_col2im(dcol, g):
    lib = _lib.load()
    da = torch.empty((g.B, g.T, g.F, g.Cin), dtype=torch.bfloat16, device=dcol.device)
    _lib.check(lib.tfasr_conv_col2im(_lib.ptr(dcol), _lib.ptr(da), ctypes.byref(g), _st(dcol)), "tfasr_conv_col2im")
    return da


# ------------------------------------------------------------------------------------------------ persistent greedy decode
def greedy_decode(weights, e_proj, nframes, prev_tok, states, frame_idx, tok_idx, tokens, iters, cap_iter, blank, mode, max_iters=0, per_frame_cap=1,
                  threads=512):
    """One launch of the persistent greedy transducer decode kernel (include/tfasr_b200.h: tfasr_greedy_decode_f32).  `weights` = dict of fp32
    contiguous tensors emb [V,E], w_ih [E,4R], w_hh [R,4R], b [4R], ln_g / ln_b [R] or None, wp [R,J], bp [J], wv [J,V], bv [V]; e_proj [B,T,J] f32;
    the remaining tensors are the in/out decoder state (int32 except states [B,2,R] f32)."""
    lib = _lib.load()
    B, T, J = e_proj.shape
    P = _lib.GreedyParams()
    for n in ("emb", "w_ih", "w_hh", "b", "ln_g", "ln_b", "wp", "bp", "wv", "bv"):
        t = weights.get(n)
        assert t is None or (t.dtype == torch.float32), n
        setattr(P, n, _lib.ptr(t))
    for n, t in (("e_proj", e_proj), ("nframes", nframes), ("prev_tok", prev_tok), ("states", states), ("frame_idx", frame_idx), ("tok_idx", tok_idx),
                 ("tokens", tokens), ("iters", iters), ("cap_iter", cap_iter)):
        assert t.dtype == (torch.float32 if n in ("e_proj", "states") else torch.int32), n
        setattr(P, n, _lib.ptr(t))
    V, E = weights["emb"].shape
    R = weights["w_hh"].shape[0]
    assert weights["w_ih"].shape == (E, 4 * R) and weights["wp"].shape == (R, J) and weights["wv"].shape == (J, V) and states.shape == (B, 2, R)
    P.T, P.E, P.R, P.J, P.V, P.max_tokens, P.blank = T, E, R, J, V, tokens.shape[1], int(blank)
    P.mode, P.max_iters, P.per_frame_cap = int(mode), int(max_iters), int(per_frame_cap)
    if _lib.EMULATED:
        threads = min(int(threads), 64)  # one OS thread per CUDA thread on the emulator: keep the CTA small
    _lib.check(lib.tfasr_greedy_decode_f32(ctypes.byref(P), B, int(threads), _st(e_proj)), "tfasr_greedy_decode_f32")
