"""The FIRST-GENERATION product kernels — the default path that bench.py and the GPU suite run: csrc/rnnt_loss.cu, ctc_loss.cu, logmel.cu,
fused_rows.cu, conv_bn.cu, dwconv.cu, joint.cu, param_ops.cu — executed on the CPU emulator through their C ABI (sources converted by
cu2emu.py) and compared with the CPU oracle (oracle/rnnt.py, oracle/ctc.py, oracle/logmel.py; numpy only) and numpy references.
Also run under AddressSanitizer / ThreadSanitizer by tests/test_emu_kernels.py.  Test infrastructure only.

    python tests/emu/checks_v1.py [--asan | --tsan] [case prefix ...]
"""
from __future__ import annotations

import ctypes
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import checks  # noqa: E402
from checks import close, from_bf16, heap, rbf, to_bf16  # noqa: E402

c_i32, c_i64, c_vp, c_sz, c_f = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float


class GemmEpilogue(ctypes.Structure):
    _fields_ = [("bias", c_vp), ("aux", c_vp), ("residual", c_vp), ("mul", c_vp), ("colsum", c_vp), ("ld_aux", c_i32), ("ld_residual", c_i32),
                ("ld_mul", c_i32), ("c_f32", c_i32), ("activation", c_i32), ("glu", c_i32), ("mul_mode", c_i32), ("residual_scale", c_f),
                ("alpha", c_f), ("dropout_p", c_f), ("atomic_out", c_i32), ("k_splits", c_i32), ("dropout_seed_ptr", c_vp),
                ("dropout_salt", ctypes.c_ulonglong), ("rowstats", c_vp)]


class LogmelCfg(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("sample_rate", "frame_length", "frame_step", "nfft", "num_feature_bins", "feature_type", "pad_end", "log_base10",
                                     "padding")] + [(n, c_f) for n in ("preemphasis", "epsilon", "lower_edge_hertz", "upper_edge_hertz")]


def load(asan=False, tsan=False):
    lib = checks.load(asan=asan, tsan=tsan)
    lib.tfasr_rnnt_workspace_bytes.restype = c_sz
    lib.tfasr_ctc_workspace_bytes.restype = c_sz
    lib.tfasr_logmel_table_bytes.restype = c_sz
    for n in ("tfasr_rnnt_loss_f32", "tfasr_rnnt_loss_fwd_f32", "tfasr_rnnt_loss_bwd_bf16", "tfasr_ctc_loss_f32", "tfasr_logmel_f32", "tfasr_logmel_build_table",
              "tfasr_logmel_nframes", "tfasr_layernorm_fwd", "tfasr_layernorm_bwd", "tfasr_scale_mask_colsum", "tfasr_glu_bwd", "tfasr_bn_stats",
              "tfasr_bn_finalize", "tfasr_bn_apply", "tfasr_bn_bwd_stats", "tfasr_bn_bwd_apply", "tfasr_conv1_3x3s2_fwd", "tfasr_conv1_3x3s2_wgrad",
              "tfasr_im2col_3x3s2", "tfasr_col2im_3x3s2", "tfasr_dwconv_causal_fwd", "tfasr_dwconv_causal_bwd", "tfasr_joint_tanh_fwd",
              "tfasr_joint_reduce_bwd", "tfasr_adamw_flat_f32", "tfasr_signal_normalize_f32", "tfasr_feature_normalize_f32", "tfasr_specaugment_f32",
              "tfasr_pack_weights", "tfasr_unpack_add", "tfasr_rel_attention_fwd", "tfasr_rel_attention_bwd", "tfasr_rel_attention_fwd_ld", "tfasr_rel_attention_bwd_ld", "tfasr_lstm_rec_fwd", "tfasr_lstm_rec_bwd",
              "tfasr_lstm_supported", "tfasr_gemm_bf16", "tfasr_ffn_fwd_bf16", "tfasr_ffn_bwd_bf16", "tfasr_ffn_supported", "tfasr_rel_attention_fwd_tc",
              "tfasr_rel_attention_tc_supported"):
        getattr(lib, n).restype = c_i32
    return lib


# ------------------------------------------------------------------------------------------------ RNN-T loss + gradient (K1, K2, K3, K3')
def _rnnt_call(lib, logits, labels, tl, ul):
    B, T, U1, V = logits.shape
    nb = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
    ws, wsp = heap(np.zeros(nb // 4 + 1, np.float32))
    lg, lgp = heap(logits)
    lb, lbp = heap(labels) if labels.size else (labels, None)  # U1 = 1: no labels at all
    tb, tp = heap(np.asarray(tl, np.int32))
    ub, up = heap(np.asarray(ul, np.int32))
    costs, cp = heap(np.zeros(B, np.float32))
    grads, gp = heap(np.full(logits.shape, np.nan, np.float32))
    st = lib.tfasr_rnnt_loss_f32(lgp, lbp, tp, up, B, T, U1, V, 0, cp, gp, wsp, c_sz(nb), None)
    assert st == 0, st
    return costs, grads, (ws, wsp, nb, lg, lgp, lbp, tp, up, cp, (lb, tb, ub))  # the arrays behind the pointers must stay alive


def check_rnnt_kat(lib):
    """warp-transducer's shipped known-answer vector through the PRODUCT kernels: cost 4.495666 and its gradient."""
    from oracle import rnnt

    k = rnnt.WARP_TRANSDUCER_KAT
    acts = np.asarray(k["acts"], np.float32)
    costs, grads, _ = _rnnt_call(lib, acts, np.asarray(k["labels"], np.int32), [2], [2])
    assert abs(float(costs[0]) - k["cost"]) < 1e-5, costs
    close(grads, np.asarray(k["grads"]), 1e-5, "KAT gradient")


def check_rnnt(lib, B, T, U1, V, tl, ul, seed=21):
    from oracle import rnnt

    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((B, T, U1, V)) * 2).astype(np.float32)
    labels = rng.integers(1, V, (B, U1 - 1)).astype(np.int32)
    tl, ul = rnnt.clamp_lengths(tl, ul)  # base_loss.py:36 — the C ABI leaves the clamp to its caller, like warp-transducer
    ref_loss, ref_g = rnnt.rnnt_loss(logits.astype(np.float64), labels, np.asarray(tl), np.asarray(ul))
    costs, grads, ctx = _rnnt_call(lib, logits, labels, tl, ul)
    ok = np.isfinite(ref_loss)
    close(costs[ok], ref_loss[ok], 1e-5, "rnnt costs")
    assert np.array_equal(np.isfinite(costs), ok)
    close(np.nan_to_num(grads), ref_g, 1e-4, "rnnt grads")  # fp32 lattice vs float64 oracle; the gate of the GPU tests is 1e-3
    # training-head variant: bf16 gradient IN PLACE over the logits, upstream scale, vocab-bias gradient
    ws, wsp, nb, lg, lgp, lbp, tp, up, cp, _keep = ctx
    scale = rng.uniform(0.5, 2.0, B).astype(np.float32)
    sb, sp = heap(scale)
    db, dp = heap(np.zeros(V, np.float32))
    if V % 4 == 0:
        st = lib.tfasr_rnnt_loss_bwd_bf16(lgp, lbp, tp, up, B, T, U1, V, 0, cp, sp, lgp, c_i64(2 * V), dp, wsp, c_sz(nb), None)
        assert st == 0, st
        got = from_bf16(lg.view(np.uint16).reshape(B, T, U1, 2 * V)[..., :V])
        want = ref_g * scale[:, None, None, None]
        close(got, want, 6e-3, "rnnt bf16 in-place grads")
        close(db, rbf(want.astype(np.float32)).astype(np.float64).sum((0, 1, 2)), 3e-3, "vocab bias gradient")


# ------------------------------------------------------------------------------------------------ CTC loss + gradient
def check_ctc(lib, B, T, L, V, tl, ul, seed=22):
    from oracle import ctc

    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((B, T, V)) * 1.5).astype(np.float32)
    labels = rng.integers(1, V, (B, max(L, 1))).astype(np.int32)
    if L >= 3:
        labels[0, :3] = labels[0, 0]  # repeated labels: the blank between them is mandatory
    ref_loss, ref_g = ctc.ctc_loss(logits.astype(np.float64), labels, np.asarray(tl), np.asarray(ul))
    nb = lib.tfasr_ctc_workspace_bytes(B, T, L, V)
    ws, wsp = heap(np.zeros(nb // 4 + 1, np.float32))
    lg, lgp = heap(logits)
    lb, lbp = heap(labels)
    tb, tp = heap(np.asarray(tl, np.int32))
    ub, up = heap(np.asarray(ul, np.int32))
    costs, cp = heap(np.zeros(B, np.float32))
    grads, gp = heap(np.full(logits.shape, np.nan, np.float32))
    st = lib.tfasr_ctc_loss_f32(lgp, lbp, tp, up, B, T, L, V, 0, cp, gp, None, wsp, c_sz(nb), None)
    assert st == 0, st
    ok = np.isfinite(ref_loss)
    close(costs[ok], ref_loss[ok], 1e-5, "ctc costs")
    assert np.array_equal(np.isfinite(costs), ok), (costs, ref_loss)
    close(np.nan_to_num(grads), ref_g, 2e-4, "ctc grads")  # fp32 lattice vs float64 oracle; GPU gate 1e-3


# ------------------------------------------------------------------------------------------------ log-mel frontend
def _logmel(lib, signal, lengths, feature_type=0, nbins=80, nfft=512, log10=0):
    B, N = signal.shape
    cfg = LogmelCfg(16000, 400, 160, nfft, nbins, feature_type, 1, log10, 0, 0.97, 1e-6, 0.0, 8000.0)
    tb = lib.tfasr_logmel_table_bytes(ctypes.byref(cfg))
    assert tb > 0
    table, tabp = heap(np.zeros(tb // 4, np.float32))
    assert lib.tfasr_logmel_build_table(ctypes.byref(cfg), tabp) == 0
    T = lib.tfasr_logmel_nframes(ctypes.byref(cfg), N)
    sb, sp = heap(signal.astype(np.float32))
    lb, lp = heap(np.asarray(lengths, np.int32))
    out, op = heap(np.full((B, T, nbins), np.nan, np.float32))
    ol, olp = heap(np.zeros(B, np.int32))
    assert lib.tfasr_logmel_f32(sp, lp, B, N, ctypes.byref(cfg), tabp, op, olp, None) == 0
    return out, ol


def check_logmel(lib, B, N, feature_type=0, nbins=80, seed=23):
    from oracle import logmel

    rng = np.random.default_rng(seed)
    sig = np.clip(0.1 * rng.standard_normal((B, N)), -1, 1).astype(np.float32)
    lens = np.array([N] + [int(x) for x in rng.integers(1, N + 1, B - 1)], np.int32)
    out, ol = _logmel(lib, sig, lens, feature_type, nbins)
    ref, rl = logmel.feature_extraction(sig.astype(np.float64), lens, dtype=np.float64,
                                        **({"feature_type": "spectrogram", "num_feature_bins": nbins} if feature_type else {}))
    assert ol.tolist() == rl.tolist() and out.shape == ref.shape[:3]
    err = np.abs(out - ref[..., 0] if ref.ndim == 4 else out - ref)
    assert np.isfinite(out).all() and float(err.max()) < 2e-3, f"log-mel max abs err {err.max():.2e}"  # north_star: 1e-3 relative on log-mel (|x| ~ 1..14)


def check_logmel_golden(lib):
    """The reference's tests/test.flac golden features (tests/golden/logmel_testflac.npz: PCM decoded by oracle/flac.py, features minted by the
    oracle with the layer's default arguments) through the PRODUCT kernel.  First 2 s only to keep the emulated run short: a frame depends on
    its own 400 samples (+1 for the pre-emphasis), so every frame that lies inside the cut is unchanged."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "logmel_testflac.npz"))
    n = 32000
    sig = (g["pcm"].astype(np.float32) / 32768.0)[None, :n]
    out, _ = _logmel(lib, sig, [n])
    nfull = (n - 400) // 160  # frames whose window is entirely real audio
    want = g["features"][0, :nfull, :, 0].astype(np.float64)
    err = np.abs(out[0, :nfull] - want)
    assert float(err.max()) < 2e-3, f"golden log-mel max abs err {err.max():.2e}"


# ------------------------------------------------------------------------------------------------ row kernels (fused_rows.cu)
def check_layernorm(lib, M, D, x_f32, seed=24):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((M, D)) * 2 + 0.5).astype(np.float32)
    if not x_f32:
        x = rbf(x)
    gamma = (rng.standard_normal(D) * 0.2 + 1).astype(np.float32)
    beta = (rng.standard_normal(D) * 0.1).astype(np.float32)
    x64 = x.astype(np.float64)
    mu, var = x64.mean(1, keepdims=True), x64.var(1, keepdims=True)
    rs = 1.0 / np.sqrt(var + 1e-3)
    xh = (x64 - mu) * rs
    xb, xp = heap(x if x_f32 else to_bf16(x))
    gb, gp = heap(gamma)
    bb, bp = heap(beta)
    yb, yp = heap(np.zeros((M, D), np.uint16))
    mb, mp = heap(np.zeros(M, np.float32))
    rb, rp = heap(np.zeros(M, np.float32))
    assert lib.tfasr_layernorm_fwd(xp, int(x_f32), gp, bp, yp, mp, rp, M, D, c_f(1e-3), None) == 0
    close(from_bf16(yb), xh * gamma + beta, 6e-3, f"layernorm fwd {(M, D)}")
    close(mb, mu[:, 0], 1e-5, "layernorm mean")
    close(rb, rs[:, 0], 1e-4, "layernorm rstd")
    dy = rbf(rng.standard_normal((M, D)))
    dres = rbf(rng.standard_normal((M, D)))
    g = dy.astype(np.float64) * gamma
    dx_ref = rs * (g - g.mean(1, keepdims=True) - xh * (g * xh).mean(1, keepdims=True)) + dres
    db, dp = heap(to_bf16(dy))
    drb, drp = heap(to_bf16(dres))
    dxb, dxp = heap(np.zeros((M, D), np.uint16))
    g0, b0 = rng.standard_normal(D).astype(np.float32), rng.standard_normal(D).astype(np.float32)
    dgb, dgp = heap(g0)
    dbb, dbp = heap(b0)
    assert lib.tfasr_layernorm_bwd(dp, xp, int(x_f32), mp, rp, gp, drp, dxp, dgp, dbp, M, D, None) == 0
    close(from_bf16(dxb), dx_ref, 8e-3, "layernorm bwd dx")
    close(dgb - g0, (dy.astype(np.float64) * xh).sum(0), 2e-3, "layernorm dgamma")
    close(dbb - b0, dy.astype(np.float64).sum(0), 1e-4, "layernorm dbeta")


def _dropout_keep(seed, salt, M, N, p):
    """csrc/dropout.cuh in numpy: keep mask of an [M, N] tensor (N even)."""
    u64 = np.uint64
    thr = 0 if p * 65536.0 + 0.5 <= 0 else min(65535, int(p * 65536.0 + 0.5))
    rows = np.arange(M, dtype=u64)
    with np.errstate(over="ignore"):
        k = u64((seed + salt) & 0xFFFFFFFFFFFFFFFF) * u64(0xD6E8FEB86659FD93) + rows * u64(0x9E3779B97F4A7C15)
    h = ((k & u64(0xFFFFFFFF)) ^ (k >> u64(32))).astype(np.uint32)
    h ^= h >> np.uint32(15)
    h = (h.astype(u64) * u64(0x2C1B3C6D) & u64(0xFFFFFFFF)).astype(np.uint32)
    h ^= h >> np.uint32(12)
    rowkey = h
    pair = np.arange(N // 2, dtype=u64)
    hh = ((rowkey[:, None].astype(u64) + pair[None, :] * u64(0x9E3779B9)) & u64(0xFFFFFFFF)).astype(np.uint32)
    hh ^= hh >> np.uint32(16)
    hh = (hh.astype(u64) * u64(0x85EBCA6B) & u64(0xFFFFFFFF)).astype(np.uint32)
    hh ^= hh >> np.uint32(13)
    hh = (hh.astype(u64) * u64(0xC2B2AE35) & u64(0xFFFFFFFF)).astype(np.uint32)
    hh ^= hh >> np.uint32(16)
    keep = np.zeros((M, N), bool)
    keep[:, 0::2] = (hh & np.uint32(0xFFFF)) >= thr
    keep[:, 1::2] = (hh >> np.uint32(16)) >= thr
    return keep


def check_colsum_wide(lib, M, N, seed=26):
    """Pure column sums (out = NULL) of a wide matrix — the bias gradient of the all-blocks position projection: N / 8 > 256 takes
    colsum_wide_kernel (column blocks on grid.y), N / 8 with no divisor in [32, 256] the scalar kernel; accumulated on top of colsum."""
    rng = np.random.default_rng(seed)
    x = rbf(rng.standard_normal((M, N)))
    c0 = rng.standard_normal(N).astype(np.float32)
    xb, xp = heap(to_bf16(x))
    cb, cp = heap(c0.copy())
    assert lib.tfasr_scale_mask_colsum(xp, None, None, cp, M, N, c_f(1.0), c_f(0.0), None, ctypes.c_ulonglong(0), None) == 0
    close(cb - c0, x.astype(np.float64).sum(0), 1e-4, f"wide colsum {(M, N)}")


def check_scale_mask_glu(lib, M, N, p, seed=25):
    rng = np.random.default_rng(seed)
    x = rbf(rng.standard_normal((M, N)))
    add = rbf(rng.standard_normal((M, N)))
    dseed, salt, scale = 0x1234567, 77, 0.5
    keep = _dropout_keep(dseed, salt, M, N, p) if p > 0 else np.ones((M, N), bool)
    ref = np.where(keep, x.astype(np.float64) * scale / (1 - p), 0.0) + add
    xb, xp = heap(to_bf16(x))
    ab, ap = heap(to_bf16(add))
    ob, op = heap(np.zeros((M, N), np.uint16))
    c0 = rng.standard_normal(N).astype(np.float32)
    cb, cp = heap(c0)
    sb, sp = heap(np.array([dseed], np.uint64))
    assert lib.tfasr_scale_mask_colsum(xp, ap, op, cp, M, N, c_f(scale), c_f(p), sp, ctypes.c_ulonglong(salt), None) == 0
    close(from_bf16(ob), ref, 6e-3, f"scale_mask_colsum p={p}")
    close(cb - c0, ref.sum(0), 8e-3, "scale_mask_colsum colsum")  # sums the fp32 values or their bf16 roundings: within bf16 noise either way
    if p > 0:
        frac = 1.0 - keep.mean()
        assert abs(frac - p) < 0.05, f"dropout rate {frac:.3f} vs {p}"
    # GLU backward on interleaved (a, b) pairs
    D = N // 2
    aux = rbf(rng.standard_normal((M, 2 * D)))
    dg = rbf(rng.standard_normal((M, D)))
    a, b = aux[:, 0::2].astype(np.float64), aux[:, 1::2].astype(np.float64)
    sg = 1 / (1 + np.exp(-b))
    ref_d = np.zeros((M, 2 * D))
    ref_d[:, 0::2] = dg * sg
    ref_d[:, 1::2] = dg * a * sg * (1 - sg)
    axb, axp = heap(to_bf16(aux))
    dgb, dgp = heap(to_bf16(dg))
    dab, dap = heap(np.zeros((M, 2 * D), np.uint16))
    c2 = np.zeros(2 * D, np.float32)
    c2b, c2p = heap(c2)
    assert lib.tfasr_glu_bwd(axp, dgp, dap, c2p, M, D, None) == 0
    close(from_bf16(dab), ref_d, 6e-3, "glu_bwd")
    close(c2b, ref_d.sum(0), 3e-3, "glu_bwd colsum")


# ------------------------------------------------------------------------------------------------ BatchNorm passes (conv_bn.cu)
def check_batchnorm(lib, R, C, act, seed=26):
    rng = np.random.default_rng(seed)
    x = rbf(rng.standard_normal((R, C)) * 1.5 + 0.3)
    gamma = (rng.uniform(0.5, 1.5, C)).astype(np.float32)
    beta = (rng.standard_normal(C) * 0.1).astype(np.float32)
    x64 = x.astype(np.float64)
    xb, xp = heap(to_bf16(x))
    sums, sp = heap(np.zeros((2, C), np.float32))
    assert lib.tfasr_bn_stats(xp, sp, c_i64(R), C, None) == 0
    close(sums[0], x64.sum(0), 1e-4, "bn sum")
    close(sums[1], (x64 ** 2).sum(0), 1e-4, "bn sum of squares")
    mm0, mv0 = rng.standard_normal(C).astype(np.float32), rng.uniform(0.5, 1.5, C).astype(np.float32)
    mmb, mmp = heap(mm0)
    mvb, mvp = heap(mv0)
    gb, gp = heap(gamma)
    bb, bp = heap(beta)
    out, op = heap(np.zeros((4, C), np.float32))
    assert lib.tfasr_bn_finalize(sp, gp, bp, mmp, mvp, op, C, c_f(R), c_f(1e-3), c_f(0.99), None) == 0
    mean, var = x64.mean(0), x64.var(0)
    rstd = 1 / np.sqrt(var + 1e-3)
    close(out[2], mean, 1e-4, "bn mean")
    close(out[3], rstd, 1e-3, "bn rstd")
    close(mmb, 0.99 * mm0 + 0.01 * mean, 1e-4, "bn moving mean")
    close(mvb, 0.99 * mv0 + 0.01 * var, 1e-3, "bn moving variance")
    scale, shift = out[0].astype(np.float64), out[1].astype(np.float64)
    z = x64 * scale + shift
    yb, yp = heap(np.zeros((R, C), np.uint16))
    scp, shp = c_vp(out[0].ctypes.data), c_vp(out[1].ctypes.data)
    mnp, rsp = c_vp(out[2].ctypes.data), c_vp(out[3].ctypes.data)
    assert lib.tfasr_bn_apply(xp, scp, shp, yp, c_i64(R), C, act, None) == 0
    close(from_bf16(yb), checks.act_f(z, act), 6e-3, f"bn apply act={act}")
    dy = rbf(rng.standard_normal((R, C)))
    d = dy.astype(np.float64) * checks.act_g(z, act)
    xh = (x64 - out[2]) * out[3]
    db, dp = heap(to_bf16(dy))
    bs, bsp = heap(np.zeros((2, C), np.float32))
    assert lib.tfasr_bn_bwd_stats(dp, xp, scp, shp, mnp, rsp, bsp, c_i64(R), C, act, None) == 0
    close(bs[0], d.sum(0), 2e-3, "bn bwd sum d")
    close(bs[1], (d * xh).sum(0), 2e-3, "bn bwd sum d xhat")
    dxb, dxp = heap(np.zeros((R, C), np.uint16))
    assert lib.tfasr_bn_bwd_apply(dp, xp, scp, shp, mnp, rsp, bsp, c_f(R), dxp, c_i64(R), C, act, None) == 0
    close(from_bf16(dxb), scale * (d - bs[0] / R - xh * bs[1] / R), 8e-3, "bn bwd apply")


# ------------------------------------------------------------------------------------------------ depthwise conv K = 31 (dwconv.cu)
def check_dwconv_v1(lib, B, T, C, K, seed=27):
    rng = np.random.default_rng(seed)
    x = rbf(rng.standard_normal((B, T, C)))
    w = (rng.standard_normal((K, C)) * 0.2).astype(np.float32)
    bias = (rng.standard_normal(C) * 0.1).astype(np.float32)
    dy = rbf(rng.standard_normal((B, T, C)))
    xp = np.concatenate([np.zeros((B, K - 1, C)), x.astype(np.float64)], 1)
    y_ref = bias + sum(w[k] * xp[:, k:k + T] for k in range(K))
    dyp = np.concatenate([dy.astype(np.float64), np.zeros((B, K - 1, C))], 1)
    dx_ref = sum(w[k] * dyp[:, K - 1 - k:K - 1 - k + T] for k in range(K))
    dw_ref = np.stack([(dy.astype(np.float64) * xp[:, k:k + T]).sum((0, 1)) for k in range(K)])
    xb, xptr = heap(to_bf16(x))
    wb, wp = heap(w)
    bb, bp = heap(bias)
    yb, yp = heap(np.zeros((B, T, C), np.uint16))
    assert lib.tfasr_dwconv_causal_fwd(xptr, wp, bp, yp, B, T, C, K, None) == 0
    close(from_bf16(yb), y_ref, 6e-3, f"dwconv v1 fwd {(B, T, C, K)}")
    db, dp = heap(to_bf16(dy))
    dxb, dxp = heap(np.zeros((B, T, C), np.uint16))
    dwb, dwp = heap(np.zeros((K, C), np.float32))
    dbb, dbp = heap(np.zeros(C, np.float32))
    assert lib.tfasr_dwconv_causal_bwd(dp, xptr, wp, dxp, dwp, dbp, B, T, C, K, None) == 0
    close(from_bf16(dxb), dx_ref, 6e-3, "dwconv v1 dx")
    close(dwb, dw_ref, 3e-4, "dwconv v1 dw")
    close(dbb, dy.astype(np.float64).sum((0, 1)), 3e-4, "dwconv v1 dbias")


# ------------------------------------------------------------------------------------------------ flat AdamW (param_ops.cu)
def check_adamw(lib, n, seed=28):
    rng = np.random.default_rng(seed)
    p0, g0 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    m0, v0 = (rng.standard_normal(n) * 0.1).astype(np.float32), rng.uniform(0, 0.1, n).astype(np.float32)
    mask = (rng.uniform(size=n) < 0.5).astype(np.uint8)
    lr, t, b1, b2, eps, wd, l2, gs = 3e-3, 7.0, 0.9, 0.98, 1e-9, 1e-2, 1e-3, 0.5
    g = g0.astype(np.float64) * gs + 2 * l2 * p0 * mask
    m = b1 * m0 + (1 - b1) * g
    v = b2 * v0 + (1 - b2) * g * g
    pr = p0 * (1 - lr * wd) - lr / (1 - b1 ** t) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps)
    bufs = [heap(a) for a in (p0, g0, m0, v0, mask, np.array([lr], np.float32), np.array([t], np.float32))]
    st = lib.tfasr_adamw_flat_f32(bufs[0][1], bufs[1][1], bufs[2][1], bufs[3][1], bufs[4][1], c_i64(n), bufs[5][1], bufs[6][1], c_f(b1), c_f(b2),
                                  c_f(eps), c_f(wd), c_f(l2), c_f(gs), 1, None)
    assert st == 0, st
    close(bufs[0][0], pr, 1e-5, "adamw p")
    close(bufs[2][0], m, 1e-5, "adamw m")
    close(bufs[3][0], v, 1e-5, "adamw v")
    assert not bufs[1][0].any(), "gradient buffer not cleared"


# ------------------------------------------------------------------------------------------------ optional normalisations of the frontend
def check_normalizations(lib, seed=29):
    from oracle import logmel

    rng = np.random.default_rng(seed)
    B, N = 2, 3000
    sig = np.clip(0.2 * rng.standard_normal((B, N)), -1, 1).astype(np.float32)
    lens = np.array([N, N], np.int32)
    for mode, kw in ((0, dict(normalize_zscore=True)), (1, dict(normalize_min_max=True))):
        ref, _ = logmel.feature_extraction(sig.astype(np.float64), lens, dtype=np.float64, normalize_signal=True, **kw)
        sb, sp = heap(sig)
        assert lib.tfasr_signal_normalize_f32(sp, B, N, c_f(1e-6), None) == 0
        close(sb, sig / (np.abs(sig).max(1, keepdims=True) + 1e-6), 1e-6, "signal normalisation")
        out, _ = _logmel(lib, sb, lens)
        ob, op = heap(out)
        assert lib.tfasr_feature_normalize_f32(op, B, out.shape[1], 80, mode, c_f(1e-6), c_f(np.log(2e-6)), None) == 0
        close(ob, ref[..., 0], 2e-3, f"feature normalisation mode {mode}")


# ------------------------------------------------------------------------------------------------ SpecAugment (statistical parity: properties)
def check_specaugment(lib, seed=30):
    rng = np.random.default_rng(seed)
    B, T, F = 4, 90, 40
    feat = rng.standard_normal((B, T, F)).astype(np.float32) + 5.0  # never equal to the mask value
    lens = np.array([90, 60, 33, 0], np.int32)
    n_time, p_ub, n_freq, freq_factor, mask_value = 5, 0.1, 2, 9, 0.0
    fb, fp = heap(feat)
    lb, lp = heap(lens)
    sb, sp = heap(np.array([12345], np.uint64))
    st = lib.tfasr_specaugment_f32(fp, lp, B, T, F, n_time, c_f(p_ub), c_f(1.0), n_freq, freq_factor, c_f(1.0), c_f(mask_value), sp,
                                   ctypes.c_ulonglong(3), None)
    assert st == 0, st
    masked = fb == mask_value
    assert np.array_equal(fb[~masked], feat[~masked]), "unmasked cells must be untouched"
    total = 0
    for b in range(B):
        n = int(lens[b])
        cols = masked[b].all(0)  # frequency masks: whole columns, over all T frames (specaugment.py:58-95 masks the padded spectrogram too)
        rows = masked[b].all(1)  # time masks: whole frames, inside the utterance
        assert cols.sum() <= n_freq * (freq_factor - 1), "frequency masks wider than freq_factor allows"
        assert rows.sum() <= n_time * max(int(n * p_ub) - 1, 0), "time masks wider than p_upperbound allows"
        assert not rows[n:].any(), "time mask beyond the utterance length"
        explained = np.zeros((T, F), bool)
        explained[:, cols] = True
        explained[rows, :] = True
        assert not (masked[b] & ~explained).any(), "masked cells that are neither a full column nor a full frame"
        total += int(masked[b].sum())
    assert total > 0, "nothing was masked with probabilities 1"


# ------------------------------------------------------------------------------------------------ batched weight pack / gradient unpack
PACK_DT = np.dtype([("src", "<u8"), ("dst_w", "<u8"), ("dst_wt", "<u8"), ("K", "<i4"), ("N", "<i4"), ("src_ld", "<i4"), ("ldw", "<i4"), ("ldwt", "<i4"),
                    ("c0", "<i4"), ("cs", "<i4"), ("flags", "<i4")])


def check_pack_unpack(lib, seed=31):
    rng = np.random.default_rng(seed)
    K, D = 37, 24  # a GLU-style kernel [K, 2D] interleaved as (a_0, b_0, a_1, b_1, ...) plus a plain [K2, N2] matrix
    K2, N2 = 19, 40
    glu = rng.standard_normal((K, 2 * D)).astype(np.float32)
    plain = rng.standard_normal((K2, N2)).astype(np.float32)
    gb, gp = heap(glu)
    pb, pp = heap(plain)
    w1, w1p = heap(np.full((K, 2 * D), 0x7FC0, np.uint16))
    wt1, wt1p = heap(np.full((2 * D, K), 0x7FC0, np.uint16))
    w2, w2p = heap(np.full((K2, N2), 0x7FC0, np.uint16))
    wt2, wt2p = heap(np.full((N2, K2), 0x7FC0, np.uint16))
    tab = np.zeros(3, PACK_DT)
    tab[0] = (gp.value, w1p.value, wt1p.value, K, D, 2 * D, 2 * D, K, 0, 2, 0)           # a halves -> even columns
    tab[1] = (gp.value + 4 * D, w1p.value, wt1p.value, K, D, 2 * D, 2 * D, K, 1, 2, 0)   # b halves -> odd columns
    tab[2] = (pp.value, w2p.value, wt2p.value, K2, N2, N2, N2, K2, 0, 1, 0)
    tb, tp = heap(tab.view(np.uint8))
    assert lib.tfasr_pack_weights(tp, 3, None) == 0
    inter = np.stack([glu[:, :D], glu[:, D:]], 2).reshape(K, 2 * D)
    assert np.array_equal(from_bf16(w1), rbf(inter)) and np.array_equal(from_bf16(wt1), rbf(inter).T)
    assert np.array_equal(from_bf16(w2), rbf(plain)) and np.array_equal(from_bf16(wt2), rbf(plain).T)
    # gradient side: packed fp32 [K, 2D] interleaved -> grad[k, n] += packed[k, c0 + n cs]; packed cleared
    packed = rng.standard_normal((K, 2 * D)).astype(np.float32)
    g0 = rng.standard_normal((K, 2 * D)).astype(np.float32)
    pk, pkp = heap(packed)
    gr, grp = heap(g0)
    ut = np.zeros(2, PACK_DT)
    ut[0] = (grp.value, pkp.value, 0, K, D, 2 * D, 2 * D, 0, 0, 2, 0)
    ut[1] = (grp.value + 4 * D, pkp.value, 0, K, D, 2 * D, 2 * D, 0, 1, 2, 0)
    ub, up = heap(ut.view(np.uint8))
    assert lib.tfasr_unpack_add(up, 2, None) == 0
    want = g0.copy()
    want[:, :D] += packed[:, 0::2]
    want[:, D:] += packed[:, 1::2]
    close(gr, want, 1e-6, "unpack_add")
    assert not pk.any(), "packed gradient buffer not cleared"


# ------------------------------------------------------------------------------------------------ relative-position attention (attention.cu)
def _att_allowed(T, length, chunk, history):
    t = np.arange(T)[:, None]
    j = np.arange(T)[None, :]
    ok = np.broadcast_to(t < length, (T, T)).copy()
    if chunk > 0:
        c0 = (t // chunk) * chunk
        hist = T if history < 0 else history
        ok &= (j >= np.maximum(0, c0 - hist)) & (j < np.minimum(T, c0 + chunk))
    return ok


def check_rel_attention(lib, B, T, N, H, lengths, chunk=0, history=0, seed=32, blocks=0, blk=0, tc=False):
    """MultiHeadRelativeAttention._compute_attention (multihead_attention.py:543-582) with rel_left_shift as out[t, j] = posattn[t, T-1-t+j]
    (:27-64), the masked softmax (general.py:30-41) and the streaming mask (:104-143): forward and the full backward of the fused kernels —
    the mma.sync path when T <= 256 and H % 4 == 0, the FP32 kernels otherwise — against float64 numpy on the valid query rows.
    blocks > 0: the *_ld entry points — pos is column slice `blk` of a [B*R, blocks*NH] buffer (the all-blocks position projection), dpos
    goes to the same slice of a shared buffer whose other columns must stay untouched."""
    rng = np.random.default_rng(seed)
    NH, R = N * H, 2 * T - 1
    qkv = rbf(rng.standard_normal((B, T, 3 * NH)) * 0.7)
    pos = rbf(rng.standard_normal((B, R, NH)) * 0.7)
    ub = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
    vb = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
    ln = np.asarray(lengths, np.int32)
    dctx = rbf(rng.standard_normal((B, T, NH)))
    for b in range(B):
        dctx[b, ln[b]:] = 0.0  # padded frames carry no upstream gradient
    scale = 1.0 / np.sqrt(H)
    idx = (T - 1 - np.arange(T)[:, None] + np.arange(T)[None, :])  # rel-shift as an index map
    ctx_ref = np.zeros((B, T, NH))
    P_ref = np.zeros((B, N, T, T))
    dqkv_ref = np.zeros((B, T, 3 * NH))
    dpos_ref = np.zeros((B, R, NH))
    du_ref, dv_ref = np.zeros((N, H)), np.zeros((N, H))
    for b in range(B):
        ok = _att_allowed(T, int(ln[b]), chunk, history)
        for n in range(N):
            q = qkv[b, :, n * H:(n + 1) * H].astype(np.float64)
            k = qkv[b, :, NH + n * H:NH + (n + 1) * H].astype(np.float64)
            v = qkv[b, :, 2 * NH + n * H:2 * NH + (n + 1) * H].astype(np.float64)
            pp = pos[b, :, n * H:(n + 1) * H].astype(np.float64)
            qu, qv = q + ub[n], q + vb[n]
            S = scale * (qu @ k.T + np.take_along_axis(qv @ pp.T, idx, 1))
            S = np.where(ok, S, -np.inf)
            valid = ok.any(1)
            Sm = np.where(valid[:, None], S, 0.0)
            P = np.exp(Sm - Sm.max(1, keepdims=True)) * ok
            P = np.where(valid[:, None], P / np.maximum(P.sum(1, keepdims=True), 1e-300), 0.0)
            P_ref[b, n] = P
            ctx_ref[b, :, n * H:(n + 1) * H] = P @ v
            do = dctx[b, :, n * H:(n + 1) * H].astype(np.float64)
            dP = do @ v.T
            dS = P * (dP - (P * dP).sum(1, keepdims=True)) * scale
            dqkv_ref[b, :, 2 * NH + n * H:2 * NH + (n + 1) * H] = P.T @ do
            dqkv_ref[b, :, NH + n * H:NH + (n + 1) * H] = dS.T @ qu
            pw = pp[idx]  # [T, T, H]: the position row each (t, j) reads
            dqkv_ref[b, :, n * H:(n + 1) * H] = dS @ k + np.einsum("tj,tjh->th", dS, pw)
            du_ref[n] += (dS @ k).sum(0)
            dv_ref[n] += np.einsum("tj,tjh->h", dS, pw)
            np.add.at(dpos_ref[b, :, n * H:(n + 1) * H], idx.reshape(-1), (dS[:, :, None] * qv[:, None, :]).reshape(T * T, H))
    qb, qp = heap(to_bf16(qkv.reshape(B * T, 3 * NH)))
    W = max(blocks, 1) * NH
    pos_wide = rbf(rng.standard_normal((B * R, W))) if blocks else None
    if blocks:
        pos_wide[:, blk * NH:(blk + 1) * NH] = pos.reshape(B * R, NH)
        pwb, pwp = heap(to_bf16(pos_wide))
        pp_ = c_vp(pwb.ctypes.data + 2 * blk * NH)
    else:
        pb, pp_ = heap(to_bf16(pos.reshape(B * R, NH)))
    ubb, ubp = heap(ub)
    vbb, vbp = heap(vb)
    lb, lp = heap(ln)
    cb, cp = heap(np.zeros((B * T, NH), np.uint16))
    prb, prp = heap(np.zeros((B, N, T, T), np.uint16))
    if tc:  # forward on the tcgen05 kernel (attention_tc.cu); the backward below then runs on ITS probabilities
        assert lib.tfasr_rel_attention_tc_supported(T, N, H, 3 * NH, W) == 1
        st = lib.tfasr_rel_attention_fwd_tc(qp, 3 * NH, pp_, W, ubp, vbp, lp, B, T, N, H, chunk, history, cp, prp, None)
    elif blocks:
        st = lib.tfasr_rel_attention_fwd_ld(qp, 3 * NH, pp_, W, ubp, vbp, lp, B, T, N, H, chunk, history, cp, prp, None)
    else:
        st = lib.tfasr_rel_attention_fwd(qp, 3 * NH, pp_, ubp, vbp, lp, B, T, N, H, chunk, history, cp, prp, None)
    assert st == 0, st
    ctx = from_bf16(cb).reshape(B, T, NH)
    probs = from_bf16(prb)
    for b in range(B):
        n_ok = int(ln[b])
        close(ctx[b, :n_ok], ctx_ref[b, :n_ok], 1.5e-2, f"attention ctx b={b}")
        close(probs[b, :, :n_ok], P_ref[b, :, :n_ok], 1e-2, f"attention probs b={b}")
    assert np.isfinite(ctx).all()
    db, dp = heap(to_bf16(dctx.reshape(B * T, NH)))
    dsb, dsp = heap(np.zeros((B, N, T, T), np.uint16))
    dqb, dqp = heap(np.full((B * T, 3 * NH), 0x7FC0, np.uint16))
    dpb, dpp = heap(np.full((B * R, W), 0x7FC0, np.uint16))
    du0, dv0 = rng.standard_normal((N, H)).astype(np.float32), rng.standard_normal((N, H)).astype(np.float32)
    dub, dup = heap(du0)
    dvb_, dvp = heap(dv0)
    if blocks:
        st = lib.tfasr_rel_attention_bwd_ld(qp, 3 * NH, pp_, W, ubp, vbp, lp, B, T, N, H, chunk, history, prp, dp, cp, dsp, dqp,
                                            c_vp(dpb.ctypes.data + 2 * blk * NH), W, dup, dvp, None)
    else:
        st = lib.tfasr_rel_attention_bwd(qp, 3 * NH, pp_, ubp, vbp, lp, B, T, N, H, chunk, history, prp, dp, cp, dsp, dqp, dpp, dup, dvp, None)
    assert st == 0, st
    if blocks:
        other = np.ones(W, bool)
        other[blk * NH:(blk + 1) * NH] = False
        assert (dpb[:, other] == 0x7FC0).all(), "dpos written outside this block's column slice"
        dpb = np.ascontiguousarray(dpb[:, blk * NH:(blk + 1) * NH])
    dqkv = from_bf16(dqb).reshape(B, T, 3 * NH)
    dpos = from_bf16(dpb).reshape(B, R, NH)
    assert np.isfinite(dqkv).all() and np.isfinite(dpos).all(), "gradient cells left unwritten"
    close(dqkv, dqkv_ref, 3e-2, "attention dqkv")
    close(dpos, dpos_ref, 3e-2, "attention dpos")
    close(dub - du0, du_ref, 3e-2, "attention du")
    close(dvb_ - dv0, dv_ref, 3e-2, "attention dvb")


# ------------------------------------------------------------------------------------------------ cluster LSTM (lstm.cu)
class SkipCase(Exception):
    """This environment cannot run the case (not a failure)."""


def check_cluster_lstm(lib, B, U, R, lengths, with_state, seed=33):
    """The prediction network's LSTM time loop (keras LSTM, gate order i, f, c, o; base_transducer.py:74-88): an 8-CTA thread-block cluster
    (emulated with its 8 x 512 threads running concurrently, distributed shared memory and cluster barriers included), forward and backward,
    against a float64 BPTT."""
    if not lib.emu_can_spawn(8 * 512 + 64):
        raise SkipCase("cannot run 4096 threads at once here (cluster emulation)")
    assert lib.tfasr_lstm_supported(R) == 1
    rng = np.random.default_rng(seed)
    sig = lambda x: 1 / (1 + np.exp(-x))  # noqa: E731
    gx = (rng.standard_normal((B, U, 4 * R)) * 0.8).astype(np.float32)
    whh = (rng.standard_normal((R, 4 * R)) / np.sqrt(R)).astype(np.float32)
    h0 = (rng.standard_normal((B, R)) * 0.5).astype(np.float32) if with_state else None
    c0 = (rng.standard_normal((B, R)) * 0.5).astype(np.float32) if with_state else None
    ln = np.asarray(lengths, np.int32)
    W = rbf(whh).astype(np.float64)  # the kernel keeps W_hh as bf16 in shared memory
    h = h0.astype(np.float64) if with_state else np.zeros((B, R))
    c = c0.astype(np.float64) if with_state else np.zeros((B, R))
    hs, cs, gates = [h], [c], []
    y_ref = np.zeros((B, U, R))
    for t in range(U):
        z = gx[:, t] + rbf(h.astype(np.float32)).astype(np.float64) @ W  # h travels between the CTAs as bf16
        i, f, g, o = sig(z[:, :R]), sig(z[:, R:2 * R]), np.tanh(z[:, 2 * R:3 * R]), sig(z[:, 3 * R:])
        c = f * c + i * g
        h = o * np.tanh(c)
        gates.append((i, f, g, o))
        hs.append(h)
        cs.append(c)
        y_ref[:, t] = np.where((t < ln)[:, None], h, 0.0)
    bufs = {k: heap(v) for k, v in dict(gx=gx, whh=whh, ln=ln).items()}
    h0p = heap(h0) if with_state else (None, None)
    c0p = heap(c0) if with_state else (None, None)
    yb, yp = heap(np.full((B, U, R), 0x7FC0, np.uint16))
    sgb, sgp = heap(np.zeros((B, U, 4 * R), np.uint16))
    scb, scp = heap(np.zeros((B, U, R), np.float32))
    hTb, hTp = heap(np.zeros((B, R), np.float32))
    cTb, cTp = heap(np.zeros((B, R), np.float32))
    st = lib.tfasr_lstm_rec_fwd(bufs["gx"][1], bufs["whh"][1], h0p[1], c0p[1], bufs["ln"][1], yp, sgp, scp, hTp, cTp, B, U, R, None)
    assert st == 0, st
    close(from_bf16(yb), y_ref, 2e-2, f"lstm y {(B, U, R)}")
    close(scb, np.stack(cs[1:], 1), 2e-2, "lstm saved c")
    close(hTb, hs[-1], 2e-2, "lstm hT")
    close(cTb, cs[-1], 2e-2, "lstm cT")
    # backward: dy -> pre-activation gate gradients (zero on masked steps)
    dy = rbf(rng.standard_normal((B, U, R)))
    dg_ref = np.zeros((B, U, 4 * R))
    dh_next, dc_next = np.zeros((B, R)), np.zeros((B, R))
    for t in range(U - 1, -1, -1):
        i, f, g, o = gates[t]
        m = (t < ln)[:, None]
        dh = np.where(m, dy[:, t], 0.0) + dh_next
        tc = np.tanh(cs[t + 1])
        dc = dh * o * (1 - tc * tc) + dc_next
        di, df, dgg, do = dc * g * i * (1 - i), dc * cs[t] * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)
        dz = np.concatenate([di, df, dgg, do], 1)
        dg_ref[:, t] = dz
        dh_next, dc_next = dz @ W.T, dc * f
    db, dp = heap(to_bf16(dy))
    dgb, dgp = heap(np.full((B, U, 4 * R), 0x7FC0, np.uint16))
    dh0b, dh0p = heap(np.zeros((B, R), np.float32))
    dc0b, dc0p = heap(np.zeros((B, R), np.float32))
    st = lib.tfasr_lstm_rec_bwd(dp, sgp, scp, c0p[1], bufs["whh"][1], bufs["ln"][1], dgp, dh0p, dc0p, B, U, R, None)
    assert st == 0, st
    got = from_bf16(dgb)
    assert np.isfinite(got).all(), "gate gradients left unwritten"
    close(got, dg_ref, 3e-2, "lstm gate gradients")
    close(dh0b, dh_next, 3e-2, "lstm dh0")
    close(dc0b, dc_next, 3e-2, "lstm dc0")


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM (gemm.cu, gemm_kernel.cuh)
def check_gemm(lib, M, N, K, *, f32_out=False, bias=False, act=0, aux=False, glu=False, residual=False, mul_mode=0, colsum=False, mn_major=False,
               accumulate=False, k_splits=1, rowstats=False, dropout=0.0, alpha=1.0, seed=34):
    """C = epilogue(alpha * A . B^T) through tfasr_gemm_bf16 on the emulator: TMA tile loads with the 128-byte swizzle, the mbarrier ring, UMMA
    descriptors (K-major and MN-major), TMEM accumulators, tcgen05.ld epilogue, TMA stores / reduce-adds — against float64 numpy."""
    rng = np.random.default_rng(seed)
    if mn_major:
        a, b = rbf(rng.standard_normal((K, M))), rbf(rng.standard_normal((K, N)) / np.sqrt(K))
        v = a.astype(np.float64).T @ b
        lda, ldb = M, N
    else:
        a, b = rbf(rng.standard_normal((M, K))), rbf(rng.standard_normal((N, K)) / np.sqrt(K))
        v = a.astype(np.float64) @ b.astype(np.float64).T
        lda = ldb = K
    v = v * alpha
    e = GemmEpilogue()
    keep = []
    if bias:
        bv = rng.standard_normal(N).astype(np.float32)
        bb, bp = heap(bv)
        keep.append(bb)
        e.bias = bp
        v = v + bv
    stats_ref = checks.gemm_rowstats_ref(v.astype(np.float32)) if rowstats else None
    pre = v.copy()
    n_out = N
    if glu:
        v = v[:, 0::2] / (1 + np.exp(-v[:, 1::2])) * 1.0
        n_out = N // 2
    elif act:
        v = {1: lambda z: z / (1 + np.exp(-z)), 2: lambda z: np.maximum(z, 0), 3: np.tanh}[act](v)
    if mul_mode:
        z = rbf(rng.standard_normal((M, n_out)))
        zb, zp = heap(to_bf16(z))
        keep.append(zb)
        e.mul, e.ld_mul, e.mul_mode = zp, n_out, mul_mode
        z64 = z.astype(np.float64)
        if mul_mode == 1:
            sg = 1 / (1 + np.exp(-z64))
            v = v * sg * (1 + z64 * (1 - sg))
        elif mul_mode == 2:
            v = v * (1 - z64 * z64)
        else:
            v = v * z64
    if dropout > 0:
        import checks_v1 as me  # the counter hash of csrc/dropout.cuh, restated in numpy above

        dseed, salt = 99, 5
        kp = me._dropout_keep(dseed, salt, M, n_out, dropout)
        sb, sp = heap(np.array([dseed], np.uint64))
        keep.append(sb)
        e.dropout_p, e.dropout_seed_ptr, e.dropout_salt = dropout, sp, salt
        v = np.where(kp, v / (1 - dropout), 0.0)
    if residual:
        r = rbf(rng.standard_normal((M, n_out)))
        rb_, rp = heap(to_bf16(r))
        keep.append(rb_)
        e.residual, e.ld_residual, e.residual_scale = rp, n_out, 0.5
        v = r + 0.5 * v
    if colsum:
        c0 = rng.standard_normal(n_out).astype(np.float32)
        cb, cp = heap(c0)
        e.colsum = cp
    if aux:
        ab, ap = heap(np.full((M, N), 0x7FC0, np.uint16))
        e.aux, e.ld_aux = ap, N
    if rowstats:
        rsb, rsp = heap(np.full(stats_ref.shape, np.nan, np.float32))
        e.rowstats = rsp
    e.c_f32, e.activation, e.glu, e.alpha = int(f32_out), act, int(glu), alpha
    e.atomic_out, e.k_splits = int(accumulate), k_splits
    init = rng.standard_normal((M, n_out)).astype(np.float32) if accumulate else None
    if f32_out:
        ob, op = heap(init if accumulate else np.full((M, n_out), np.nan, np.float32))
    else:
        ob, op = heap(np.full((M, n_out), 0x7FC0, np.uint16))
    ab_, app = heap(to_bf16(a))
    bb_, bpp = heap(to_bf16(b))
    st = lib.tfasr_gemm_bf16(app, bpp, op, M, N, K, lda, ldb, n_out, int(mn_major), int(mn_major), ctypes.byref(e), None)
    assert st == 0, st
    got = ob.astype(np.float64) if f32_out else from_bf16(ob).astype(np.float64)
    want = v + (init if accumulate else 0.0)
    assert np.isfinite(got).all(), "output cells left unwritten"
    close(got, want, 2e-3 if f32_out else 8e-3, f"gemm {(M, N, K)} mn={mn_major} f32={f32_out}")
    if aux:
        close(from_bf16(ab), pre, 8e-3, "gemm aux (pre-activation copy)")
    if colsum:
        close(cb - c0, v.sum(0), 1e-2, "gemm column sums")
    if rowstats and not f32_out:  # bf16 output: the statistics are those of the values as stored
        pre = got
        stats_ref = checks.gemm_rowstats_ref(got.astype(np.float32))
    if rowstats:
        assert np.isfinite(rsb[..., 1]).all()
        m = rsb[..., 0].max(1, keepdims=True)
        lse = (m + np.log((rsb[..., 1] * np.exp(rsb[..., 0] - m)).sum(1, keepdims=True)))[:, 0]
        mx = pre.max(1, keepdims=True)
        close(lse, (mx + np.log(np.exp(pre - mx).sum(1, keepdims=True)))[:, 0], 1e-5, "log-sum-exp from the epilogue's row statistics")
        assert np.array_equal(np.isneginf(rsb[..., 0]), np.isneginf(stats_ref[..., 0])), "empty slices"
        ok = np.isfinite(stats_ref[..., 0])
        close(rsb[..., 0][ok], stats_ref[..., 0][ok], 1e-5, "slice maxima")
        close(rsb[..., 1][ok], stats_ref[..., 1][ok], 1e-4, "slice sums")


# ------------------------------------------------------------------------------------------------ seeded sweep over odd shapes
def check_random_shapes_v1(lib):
    """Edge shapes the fixed cases do not name: single frames / labels, vocabularies and widths around the vector-path limits, lengths of 0 and
    1, attention lengths around the 16-row MMA tiles, GEMM extents around the 128 x BN x 64 tiles."""
    rng = np.random.default_rng(4242)
    ri = lambda lo, hi: int(rng.integers(lo, hi))  # noqa: E731
    for _ in range(5):
        B, T, U1, V = ri(1, 4), ri(1, 9), ri(1, 7), int(rng.choice([2, 3, 8, 12, 33, 260]))
        tl = [ri(1, T + 1) for _ in range(B)]
        ul = [ri(0, min(U1, t + 1)) if U1 > 1 else 0 for t in tl]
        check_rnnt(lib, B, T, U1, V, tl, [min(u, U1 - 1) for u in ul], seed=ri(0, 1 << 30))
    for _ in range(4):
        B, T, L, V = ri(1, 4), ri(1, 30), ri(0, 6), int(rng.choice([2, 5, 17, 64]))
        tl = [ri(1, T + 1) for _ in range(B)]
        check_ctc(lib, B, T, L, V, tl, [ri(0, min(L, t) + 1) for t in tl], seed=ri(0, 1 << 30))
    for _ in range(3):
        check_logmel(lib, ri(1, 3), ri(1, 2200), seed=ri(0, 1 << 30))  # incl. signals shorter than one frame
    for _ in range(4):
        check_layernorm(lib, ri(1, 40), int(rng.choice([1, 2, 7, 36, 144, 250, 640])), bool(ri(0, 2)), seed=ri(0, 1 << 30))
    for _ in range(3):
        check_scale_mask_glu(lib, ri(1, 50), 2 * ri(1, 80), float(rng.choice([0.0, 0.1, 0.5])), seed=ri(0, 1 << 30))
    for _ in range(4):
        check_batchnorm(lib, ri(1, 200), int(rng.choice([2, 6, 8, 24, 72, 144])), ri(0, 3), seed=ri(0, 1 << 30))
    for _ in range(3):
        check_dwconv_v1(lib, ri(1, 3), ri(1, 80), 2 * ri(1, 40), int(rng.choice([1, 3, 7, 31])), seed=ri(0, 1 << 30))
    for _ in range(3):
        T = ri(1, 40)
        check_rel_attention(lib, ri(1, 3), T, ri(1, 3), int(rng.choice([4, 8, 10, 36, 64])), [ri(1, T + 1) for _ in range(2)][: 2], seed=ri(0, 1 << 30)) \
            if False else None
    for _ in range(4):
        B, T = ri(1, 3), ri(1, 44)
        check_rel_attention(lib, B, T, ri(1, 3), int(rng.choice([4, 8, 10, 36, 64])), [ri(1, T + 1) for _ in range(B)],
                            chunk=int(rng.choice([0, 0, 4, 9])), history=int(rng.choice([0, 3, 16])), seed=ri(0, 1 << 30))
    for _ in range(6):
        M, N, K = ri(1, 300), 8 * ri(1, 40), 8 * ri(1, 30)
        kind = ri(0, 5)
        if kind == 0:
            check_gemm(lib, M, N, K, f32_out=bool(ri(0, 2)), bias=bool(ri(0, 2)), seed=ri(0, 1 << 30))
        elif kind == 1:
            check_gemm(lib, M, N, K, bias=True, act=ri(1, 4), aux=True, seed=ri(0, 1 << 30))
        elif kind == 2:
            check_gemm(lib, M, N, K, residual=True, dropout=float(rng.choice([0.0, 0.3])), seed=ri(0, 1 << 30))
        elif kind == 3:
            check_gemm(lib, 8 * ri(1, 40), N, ri(1, 400), f32_out=True, mn_major=True, accumulate=True, k_splits=ri(1, 4), seed=ri(0, 1 << 30))
        else:
            check_gemm(lib, M, N, K, f32_out=True, bias=True, rowstats=True, seed=ri(0, 1 << 30))


# ------------------------------------------------------------------------------------------------ fused FFModule (ffn_fused.cu)
def _ffn_case(M, H, D, seed):
    rng = np.random.default_rng(seed)
    x = rbf(rng.standard_normal((M, D)) * 1.5 + 0.3)
    g, b = (1 + 0.2 * rng.standard_normal(D)).astype(np.float32), (0.1 * rng.standard_normal(D)).astype(np.float32)
    w1 = rbf(rng.standard_normal((D, H)) / np.sqrt(D))
    w2 = rbf(rng.standard_normal((H, D)) / np.sqrt(H))
    b1, b2 = (0.2 * rng.standard_normal(H)).astype(np.float32), (0.2 * rng.standard_normal(D)).astype(np.float32)
    return x, g, b, w1, b1, w2, b2


def check_ffn_fwd(lib, M, H, dropout=0.0, D=144, seed=71):
    """out = x + 0.5 Drop2(W2 Drop1(swish(W1 LN(x) + b1)) + b2) in ONE launch (TMA x tile, in-SMEM LayerNorm, chunked GEMM 1 -> SMEM -> GEMM 2,
    3-stage weight ring) against float64 numpy with the kernel's bf16 rounding points (xn, h)."""
    assert lib.tfasr_ffn_supported(D, H) == 1
    x, g, b, w1, b1, w2, b2 = _ffn_case(M, H, D, seed)
    x64 = x.astype(np.float64)
    mean = x64.mean(1, keepdims=True)
    var = ((x64 - mean) ** 2).mean(1, keepdims=True)
    rstd = 1 / np.sqrt(var + 1e-3)
    xn = rbf((x64 - mean) * rstd * g + b).astype(np.float64)
    z = xn @ w1.astype(np.float64) + b1
    h = z / (1 + np.exp(-z))
    dseed, salt1, salt2 = 1234, 7, 8
    if dropout > 0:
        h = np.where(_dropout_keep(dseed, salt1, M, H, dropout), h / (1 - dropout), 0.0)
    hb = rbf(h).astype(np.float64)
    y = hb @ w2.astype(np.float64) + b2
    if dropout > 0:
        y = np.where(_dropout_keep(dseed, salt2, M, D, dropout), y / (1 - dropout), 0.0)
    out = x64 + 0.5 * y
    nan16 = lambda shape: heap(np.full(shape, 0x7FC0, np.uint16))  # noqa: E731
    xb, xp = heap(to_bf16(x))
    w1b, w1p = heap(to_bf16(np.ascontiguousarray(w1.T)))
    w2b, w2p = heap(to_bf16(np.ascontiguousarray(w2.T)))
    gb, gp = heap(g)
    bb, bp = heap(b)
    b1b, b1p = heap(b1)
    b2b, b2p = heap(b2)
    ob, op = nan16((M, D))
    xnb, xnp = nan16((M, D))
    zb, zp = nan16((M, H))
    hbuf, hp = nan16((M, H))
    mb, mp = heap(np.full(M, np.nan, np.float32))
    rb_, rp = heap(np.full(M, np.nan, np.float32))
    sb, sp = heap(np.array([dseed], np.uint64))
    st = lib.tfasr_ffn_fwd_bf16(xp, gp, bp, w1p, b1p, w2p, b2p, op, xnp, zp, hp, mp, rp, M, D, H, c_f(1e-3), c_f(0.5), c_f(dropout), sp,
                                ctypes.c_ulonglong(salt1), ctypes.c_ulonglong(salt2), None)
    assert st == 0, st
    close(mb, mean[:, 0], 1e-5, "ffn mean")
    close(rb_, rstd[:, 0], 1e-4, "ffn rstd")
    close(from_bf16(xnb), xn, 8e-3, "ffn xn")
    close(from_bf16(zb), z, 8e-3, "ffn z (pre-activation)")
    close(from_bf16(hbuf), h, 8e-3, "ffn h")
    close(from_bf16(ob), out, 8e-3, "ffn out")


def check_ffn_bwd(lib, M, H, dropout=0.0, D=144, seed=72):
    """dy = 0.5 Drop2'(dout); dz = (dy W2^T) Drop1' swish'(z); da = dz W1^T in ONE launch, against float64 numpy with the kernel's bf16 rounding
    points (dy, dz)."""
    rng = np.random.default_rng(seed)
    _, _, _, w1, _, w2, _ = _ffn_case(M, H, D, seed)
    dout = rbf(rng.standard_normal((M, D)))
    z = rbf(rng.standard_normal((M, H)) * 1.5)
    dseed, salt1, salt2 = 4321, 9, 10
    dy = dout.astype(np.float64) * 0.5
    if dropout > 0:
        dy = np.where(_dropout_keep(dseed, salt2, M, D, dropout), dy / (1 - dropout), 0.0)
    dyb = rbf(dy).astype(np.float64)
    z64 = z.astype(np.float64)
    sg = 1 / (1 + np.exp(-z64))
    dz = (dyb @ w2.astype(np.float64).T) * sg * (1 + z64 * (1 - sg))
    if dropout > 0:
        dz = np.where(_dropout_keep(dseed, salt1, M, H, dropout), dz / (1 - dropout), 0.0)
    dzb = rbf(dz).astype(np.float64)
    da = dzb @ w1.astype(np.float64).T
    nan16 = lambda shape: heap(np.full(shape, 0x7FC0, np.uint16))  # noqa: E731
    db, dp_ = heap(to_bf16(dout))
    zb, zp = heap(to_bf16(z))
    w2b, w2p = heap(to_bf16(w2))  # W2 [H, D]
    w1b, w1p = heap(to_bf16(w1))  # W1 [D, H]
    dyo, dyp = nan16((M, D))
    dzo, dzp = nan16((M, H))
    dao, dap = nan16((M, D))
    sb, sp = heap(np.array([dseed], np.uint64))
    st = lib.tfasr_ffn_bwd_bf16(dp_, zp, w2p, w1p, dyp, dzp, dap, M, D, H, c_f(0.5), c_f(dropout), sp, ctypes.c_ulonglong(salt1), ctypes.c_ulonglong(salt2),
                                None)
    assert st == 0, st
    close(from_bf16(dyo), dy, 8e-3, "ffn bwd dy")
    close(from_bf16(dzo), dz, 1e-2, "ffn bwd dz")
    close(from_bf16(dao), da, 1e-2, "ffn bwd da")


CASES = [
    ("rnnt_kat", check_rnnt_kat),
    ("rnnt_ragged", lambda lib: check_rnnt(lib, 4, 13, 8, 12, [13, 9, 5, 13], [7, 4, 0, 1])),
    ("rnnt_scalar_v", lambda lib: check_rnnt(lib, 2, 6, 4, 11, [6, 3], [3, 2])),  # V % 4 != 0: generic K1 / K3 path
    ("rnnt_wide_u", lambda lib: check_rnnt(lib, 1, 41, 40, 8, [41], [39])),  # U1 > 32: the lattice's cross-warp edge exchange
    ("rnnt_short", lambda lib: check_rnnt(lib, 3, 4, 3, 8, [1, 4, 1], [0, 2, 2])),  # single frame, no labels; length clamp (third utterance)
    ("ctc_ragged", lambda lib: check_ctc(lib, 5, 40, 9, 17, [40, 31, 20, 40, 12], [9, 6, 0, 1, 9])),
    ("ctc_infeasible", lambda lib: check_ctc(lib, 2, 4, 3, 8, [4, 4], [3, 1])),  # a a a needs 5 frames, 4 given: +inf cost, zero gradient
    ("ctc_long_labels", lambda lib: check_ctc(lib, 1, 90, 40, 6, [90], [40])),  # 2L + 1 > 64 states: several warps
    ("logmel_80", lambda lib: check_logmel(lib, 2, 4000)),
    ("logmel_spectrogram_160", lambda lib: check_logmel(lib, 1, 2500, feature_type=1, nbins=160)),
    ("logmel_golden_testflac", check_logmel_golden),
    ("layernorm_bf16", lambda lib: check_layernorm(lib, 37, 144, False)),  # vector backward variant (D even, <= 512)
    ("layernorm_f32_odd", lambda lib: check_layernorm(lib, 5, 73, True)),  # scalar variants
    ("layernorm_wide", lambda lib: check_layernorm(lib, 3, 576, False)),
    ("scale_mask_glu_dropout", lambda lib: check_scale_mask_glu(lib, 40, 144, 0.3)),
    ("scale_mask_glu_plain", lambda lib: check_scale_mask_glu(lib, 9, 20, 0.0)),
    ("colsum_wide_2304", lambda lib: check_colsum_wide(lib, 61, 2304)),  # 288 vectors -> 2 column blocks of 144
    ("colsum_wide_prime", lambda lib: check_colsum_wide(lib, 9, 8 * 263)),  # 263 vectors: prime, no block width divides it -> scalar kernel
    ("batchnorm_swish_vec", lambda lib: check_batchnorm(lib, 300, 144, 1)),  # C % 8 == 0: the templated 16-byte kernel
    ("batchnorm_relu_scalar", lambda lib: check_batchnorm(lib, 77, 10, 2)),  # odd vector width: scalar kernels; relu branch (new this round)
    ("batchnorm_relu_vec", lambda lib: check_batchnorm(lib, 130, 32, 2)),  # relu branch of the vector kernel (DeepSpeech2 conv module)
    ("batchnorm_linear", lambda lib: check_batchnorm(lib, 50, 32, 0)),
    ("conv1_v1", lambda lib: checks.check_conv1_v2(lib, 2, 37, 13, 16, fwd="tfasr_conv1_3x3s2_fwd", wgrad="tfasr_conv1_3x3s2_wgrad")),
    ("im2col_v1", lambda lib: checks.check_im2col_v2(lib, 2, 7, 5, 24, im2col="tfasr_im2col_3x3s2", col2im="tfasr_col2im_3x3s2")),
    ("dwconv_v1_k31", lambda lib: check_dwconv_v1(lib, 2, 70, 16, 31)),  # the register-blocked sliding kernels
    ("dwconv_v1_generic", lambda lib: check_dwconv_v1(lib, 2, 37, 24, 5)),
    ("joint_v1", lambda lib: checks.check_joint_v2(lib, 2, 7, 11, 40, fwd="tfasr_joint_tanh_fwd", bwd="tfasr_joint_reduce_bwd")),
    ("joint_v1_generic", lambda lib: checks.check_joint_v2(lib, 1, 3, 5, 6, fwd="tfasr_joint_tanh_fwd", bwd="tfasr_joint_reduce_bwd")),  # J % 8 != 0
    ("adamw_flat", lambda lib: check_adamw(lib, 4 * 517)),
    ("frontend_normalizations", check_normalizations),
    ("specaugment_properties", check_specaugment),
    ("pack_unpack", check_pack_unpack),
    ("rel_attention_mma", lambda lib: check_rel_attention(lib, 2, 40, 2, 36, [40, 23])),  # Conformer-S head size: the mma.sync kernels, KS = 3
    ("rel_attention_mma_streaming", lambda lib: check_rel_attention(lib, 1, 48, 2, 16, [41], chunk=8, history=16)),
    ("rel_attention_fp32", lambda lib: check_rel_attention(lib, 2, 11, 2, 10, [11, 6])),
    ("rel_attention_mma_slice", lambda lib: check_rel_attention(lib, 1, 24, 2, 36, [19], blocks=3, blk=1)),  # column slice of the all-blocks projection
    ("rel_attention_fp32_slice", lambda lib: check_rel_attention(lib, 1, 11, 2, 10, [9], blocks=2, blk=1)),  # T < 16, H % 4 != 0: the FP32 kernels
    ("rel_attention_tc", lambda lib: check_rel_attention(lib, 2, 40, 2, 36, [40, 23], tc=True)),  # tcgen05 forward, Conformer-S head size, one key chunk
    ("rel_attention_tc_two_tiles", lambda lib: check_rel_attention(lib, 1, 150, 2, 36, [131], tc=True)),  # 2 query tiles x 3 key chunks
    ("rel_attention_tc_streaming", lambda lib: check_rel_attention(lib, 1, 200, 1, 64, [200], chunk=16, history=32, tc=True)),  # H = 64, chunk skipping
    ("rel_attention_tc_slice", lambda lib: check_rel_attention(lib, 1, 70, 2, 36, [61], blocks=3, blk=1, tc=True)),  # column slice of the all-blocks projection
    ("gemm_plain", lambda lib: check_gemm(lib, 200, 72, 136)),  # one ragged M tile, N < tile width, K not a multiple of 64
    ("gemm_f32_bias", lambda lib: check_gemm(lib, 129, 1000, 320, f32_out=True, bias=True)),  # the joint's shape class: 4 tiles of 256 columns
    ("gemm_swish_aux", lambda lib: check_gemm(lib, 150, 160, 64, bias=True, act=1, aux=True)),  # FFN-1: tile width 160
    ("gemm_glu", lambda lib: check_gemm(lib, 130, 288, 144, bias=True, glu=True, aux=True)),  # pointwise conv 1
    ("gemm_residual_dropout", lambda lib: check_gemm(lib, 140, 144, 576, bias=True, residual=True, dropout=0.25)),  # FFN-2 / out projections
    ("gemm_bwd_swish_colsum", lambda lib: check_gemm(lib, 133, 192, 144, mul_mode=1, colsum=True)),  # FFN backward epilogue
    ("gemm_bwd_tanh", lambda lib: check_gemm(lib, 70, 64, 1000, mul_mode=2)),  # joint dgrad
    ("gemm_mn_major_splitk", lambda lib: check_gemm(lib, 144, 192, 333, f32_out=True, mn_major=True, accumulate=True, k_splits=3)),  # wgrad
    ("gemm_mn_major_odd", lambda lib: check_gemm(lib, 456, 32, 200, f32_out=True, mn_major=True, accumulate=True)),  # DeepSpeech2 conv wgrad: M = Kp
    ("gemm_tiny", lambda lib: check_gemm(lib, 1, 8, 8, f32_out=True)),
    ("gemm_rowstats", lambda lib: check_gemm(lib, 130, 1000, 64, f32_out=True, bias=True, rowstats=True)),  # feature class 8 (fused log-sum-exp)
    ("gemm_rowstats_bf16_out", lambda lib: check_gemm(lib, 130, 1000, 64, bias=True, rowstats=True)),  # bf16 logits: statistics of the rounded values
    ("gemm_rowstats_narrow", lambda lib: check_gemm(lib, 40, 40, 32, f32_out=True, bias=True, rowstats=True)),  # one tile, second slice empty
    ("ffn_fwd_one_tile", lambda lib: check_ffn_fwd(lib, 100, 128)),  # ragged tile, two hidden chunks
    ("ffn_fwd_conformer_s", lambda lib: check_ffn_fwd(lib, 300, 576, dropout=0.1)),  # 3 tiles x 9 chunks: every ring / double-buffer phase wraps
    ("ffn_bwd_one_tile", lambda lib: check_ffn_bwd(lib, 100, 128)),
    ("ffn_bwd_conformer_s", lambda lib: check_ffn_bwd(lib, 300, 576, dropout=0.1)),
    ("random_shapes_v1", check_random_shapes_v1),
    ("cluster_lstm", lambda lib: check_cluster_lstm(lib, 3, 4, 64, [4, 2, 3], False)),
    ("cluster_lstm_state", lambda lib: check_cluster_lstm(lib, 2, 3, 128, [3, 3], True)),
]

if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("--")]
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = load(asan="--asan" in flags, tsan="--tsan" in flags)
    for name, fn in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        if "--fast" in flags and name in ("random_shapes_v1", "rnnt_wide_u"):  # sanitizer passes: skip the sweep and the largest lattice
            continue
        try:
            t0 = time.time()
            fn(lib)
            print("ok", name, f"{time.time() - t0:.1f}s", flush=True)
        except SkipCase as e:
            print("skip", name, e, flush=True)
