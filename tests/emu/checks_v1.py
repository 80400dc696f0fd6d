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

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import checks  # noqa: E402
from checks import close, from_bf16, heap, rbf, to_bf16  # noqa: E402

c_i32, c_i64, c_vp, c_sz, c_f = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float


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
              "tfasr_joint_reduce_bwd", "tfasr_adamw_flat_f32"):
        getattr(lib, n).restype = c_i32
    return lib


# ------------------------------------------------------------------------------------------------ RNN-T loss + gradient (K1, K2, K3, K3')
def _rnnt_call(lib, logits, labels, tl, ul):
    B, T, U1, V = logits.shape
    nb = lib.tfasr_rnnt_workspace_bytes(B, T, U1, V)
    ws, wsp = heap(np.zeros(nb // 4 + 1, np.float32))
    lg, lgp = heap(logits)
    lb, lbp = heap(labels)
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
    labels = rng.integers(1, V, (B, max(U1 - 1, 1))).astype(np.int32)
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
]

if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("--")]
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = load(asan="--asan" in flags, tsan="--tsan" in flags)
    for name, fn in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        fn(lib)
        print("ok", name, flush=True)
