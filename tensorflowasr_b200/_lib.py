"""ctypes binding of libtfasr_b200.so (the C ABI declared in include/tfasr_b200.h).

The product path has NO CPU fallback: if the library is missing it is built with nvcc; if that is impossible, or a
call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_size_t, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFASR_LIB_PATH") or os.path.join(_PKG, "libtfasr_b200.so")  # the override is for A/B runs of two builds (tools/)


class LogmelCfg(ctypes.Structure):
    """Mirror of `tfasr_logmel_cfg`."""

    _fields_ = [
        ("sample_rate", c_int32), ("frame_length", c_int32), ("frame_step", c_int32), ("nfft", c_int32),
        ("num_feature_bins", c_int32), ("feature_type", c_int32), ("pad_end", c_int32), ("log_base10", c_int32),
        ("padding", c_int32), ("preemphasis", c_float), ("epsilon", c_float), ("lower_edge_hertz", c_float),
        ("upper_edge_hertz", c_float),
    ]


class ConvGeom(ctypes.Structure):
    """Mirror of `tfasr_conv_geom`."""

    _fields_ = [(n, c_int32) for n in ("B", "T", "F", "Cin", "To", "Fo", "kh", "kw", "sh", "sw", "pt", "pl", "Kp")]


class GreedyParams(ctypes.Structure):
    """Mirror of `tfasr_greedy_params`."""

    _fields_ = ([(n, c_void_p) for n in ("e_proj", "nframes", "emb", "w_ih", "w_hh", "b", "ln_g", "ln_b", "wp", "bp", "wv", "bv", "prev_tok", "states",
                                         "frame_idx", "tok_idx", "tokens", "iters", "cap_iter")]
                + [(n, c_int32) for n in ("T", "E", "R", "J", "V", "max_tokens", "blank", "mode", "max_iters", "per_frame_cap")])


class PeerComm(ctypes.Structure):
    """Mirror of `tfasr_peer_comm`."""

    _fields_ = [("windows", c_void_p), ("state", c_void_p), ("rank", c_int32), ("world", c_int32), ("capacity", c_int32), ("timeout_ms", c_int32)]


class GemmEpilogue(ctypes.Structure):
    """Mirror of `tfasr_gemm_epilogue`."""

    _fields_ = [
        ("bias", c_void_p), ("aux", c_void_p), ("residual", c_void_p), ("mul", c_void_p), ("colsum", c_void_p),
        ("ld_aux", c_int32), ("ld_residual", c_int32), ("ld_mul", c_int32), ("c_f32", c_int32), ("activation", c_int32),
        ("glu", c_int32), ("mul_mode", c_int32), ("residual_scale", c_float), ("alpha", c_float), ("dropout_p", c_float),
        ("atomic_out", c_int32), ("k_splits", c_int32), ("dropout_seed_ptr", c_void_p), ("dropout_salt", ctypes.c_ulonglong),
        ("rowstats", c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/tfasr_b200.h declares
SIGNATURES = {
    "tfasr_version": (c_int32, []),
    "tfasr_status_string": (c_char_p, [c_int32]),
    "tfasr_launch_count": (c_int32, []),
    "tfasr_launch_count_reset": (None, []),
    "tfasr_logmel_table_bytes": (c_size_t, [POINTER(LogmelCfg)]),
    "tfasr_logmel_build_table": (c_int32, [POINTER(LogmelCfg), c_void_p]),
    "tfasr_logmel_nframes": (c_int32, [POINTER(LogmelCfg), c_int32]),
    "tfasr_logmel_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, POINTER(LogmelCfg), c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfasr_signal_normalize_f32": (c_int32, [c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "tfasr_feature_normalize_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_void_p]),
    "tfasr_rnnt_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "tfasr_rnnt_loss_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_loss_fwd_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_lse_from_rowstats_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfasr_rnnt_loss_fwd_rowstats_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                   c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_loss_bwd_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_loss_bwd_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                           c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_loss_bwd_bf16_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                              c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_lse_from_rowstats_bf16": (c_int32, [c_void_p, ctypes.c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfasr_rnnt_loss_fwd_rowstats_bf16": (c_int32, [c_void_p, ctypes.c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                    c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_loss_bwd_bf16_io": (c_int32, [c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                              c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tfasr_rnnt_lattice_smem_supported": (c_int32, [c_int32, c_int32]),
    "tfasr_rnnt_lattice_smem_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                              c_void_p]),
    "tfasr_joint_tanh_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_joint_reduce_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_joint_tanh_fwd_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_joint_reduce_bwd_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_conv1_3x3s2_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_conv1_3x3s2_wgrad": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_conv1_3x3s2_fwd_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_conv1_3x3s2_wgrad_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_im2col_3x3s2_v2": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_col2im_3x3s2_v2": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_im2col_3x3s2": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_col2im_3x3s2": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_bn_stats": (c_int32, [c_void_p, c_void_p, ctypes.c_int64, c_int32, c_void_p]),
    "tfasr_bn_apply": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int32, c_int32, c_void_p]),
    "tfasr_bn_bwd_stats": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int32, c_int32,
                                     c_void_p]),
    "tfasr_bn_bwd_apply": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                     ctypes.c_int64, c_int32, c_int32, c_void_p]),
    "tfasr_rel_attention_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, c_void_p, c_void_p, c_void_p]),
    "tfasr_rel_attention_bwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfasr_rel_attention_tc_supported": (c_int32, [c_int32] * 5),
    "tfasr_rel_attention_fwd_tc": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                             c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "tfasr_rel_attention_fwd_ld": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                             c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "tfasr_rel_attention_bwd_ld": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                             c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                             c_void_p]),
    "tfasr_lstm_supported": (c_int32, [c_int32]),
    "tfasr_lstm_rec_fwd": (c_int32, [c_void_p] * 10 + [c_int32] * 3 + [c_void_p]),
    "tfasr_lstm_rec_bwd": (c_int32, [c_void_p] * 9 + [c_int32] * 3 + [c_void_p]),
    "tfasr_adamw_flat_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_float, c_float,
                                       c_float, c_float, c_float, c_float, c_int32, c_void_p]),
    "tfasr_pack_weights": (c_int32, [c_void_p, c_int32, c_void_p]),
    "tfasr_unpack_add": (c_int32, [c_void_p, c_int32, c_void_p]),
    "tfasr_bn_finalize": (c_int32, [c_void_p] * 6 + [c_int32, c_float, c_float, c_float, c_void_p]),
    "tfasr_bn_bwd_finalize": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "tfasr_dwconv_causal_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_dwconv_causal_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_dwconv_causal_bwd_dx": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_dwconv_causal_wgrad_v2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_specaugment_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_int32, c_int32, c_float,
                                        c_float, c_void_p, ctypes.c_ulonglong, c_void_p]),
    "tfasr_gemm_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  POINTER(GemmEpilogue), c_void_p]),
    "tfasr_ffn_supported": (c_int32, [c_int32, c_int32]),
    "tfasr_ffn_fwd_bf16": (c_int32, [c_void_p] * 13 + [c_int32, c_int32, c_int32, c_float, c_float, c_float, c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                     c_void_p]),
    "tfasr_ffn_bwd_bf16": (c_int32, [c_void_p] * 7 + [c_int32, c_int32, c_int32, c_float, c_float, c_void_p, ctypes.c_uint64, ctypes.c_uint64, c_void_p]),
    "tfasr_layernorm_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "tfasr_layernorm_bwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int32, c_int32, c_void_p]),
    "tfasr_scale_mask_colsum": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_void_p,
                                          ctypes.c_ulonglong, c_void_p]),
    "tfasr_glu_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "tfasr_seq_dwconv_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_seq_dwconv_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_seq_masked_mean": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "tfasr_se_fc_fwd": (c_int32, [c_void_p] * 7 + [c_int32] * 4 + [c_void_p]),
    "tfasr_se_fc_bwd": (c_int32, [c_void_p] * 14 + [c_int32] * 5 + [c_void_p]),
    "tfasr_se_apply_fwd": (c_int32, [c_void_p] * 4 + [c_int32] * 4 + [c_void_p]),
    "tfasr_se_apply_bwd": (c_int32, [c_void_p] * 6 + [c_int32] * 4 + [c_void_p]),
    "tfasr_se_dy": (c_int32, [c_void_p] * 5 + [c_int32] * 3 + [c_void_p]),
    "tfasr_conv_im2col": (c_int32, [c_void_p, c_int32, c_void_p, POINTER(ConvGeom), c_void_p]),
    "tfasr_conv_col2im": (c_int32, [c_void_p, c_void_p, POINTER(ConvGeom), c_void_p]),
    "tfasr_greedy_decode_f32": (c_int32, [POINTER(GreedyParams), c_int32, c_int32, c_void_p]),
    "tfasr_peer_window_bytes": (ctypes.c_size_t, [c_int32, c_int32]),
    "tfasr_peer_window_alloc": (c_int32, [ctypes.c_size_t, POINTER(c_void_p), c_void_p]),
    "tfasr_peer_window_open": (c_int32, [c_void_p, POINTER(c_void_p)]),
    "tfasr_peer_window_close": (c_int32, [c_void_p, c_int32]),
    "tfasr_peer_allreduce_f32": (c_int32, [POINTER(PeerComm), c_void_p, c_int32, c_void_p]),
    "tfasr_ctc_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "tfasr_ctc_loss_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None
EMULATED = False  # set by tests/emu_backend.py only: the kernels run as their own source on the CPU emulator, tensors live on the host


def load() -> ctypes.CDLL:
    """Load (building first if needed) libtfasr_b200.so and bind every declared symbol. Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import _build

        _build.build_library()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status == 0:
        return
    msg = load().tfasr_status_string(status).decode()
    if status == 1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libtfasr_b200 has no CPU path: tensor must live on a CUDA device")
    if not t.is_contiguous():
        raise RuntimeError("libtfasr_b200 expects contiguous row-major buffers")
    return t.data_ptr()


def ptr_rows(t) -> int | None:
    """Device pointer of a 2-D CUDA tensor whose rows are contiguous (unit column stride) but may be strided — a column slice of a wider
    buffer; the row stride travels separately (t.stride(0))."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libtfasr_b200 has no CPU path: tensor must live on a CUDA device")
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError("libtfasr_b200 expects unit column stride")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the B200 hot path has no CPU fallback; got a {t.device} tensor")


class _Timers:
    """CUDA-event timers around named kernel groups, on the launching stream (enabled by bench.py only)."""

    def __init__(self):
        self.on = False
        self.events = {}

    def enable(self, flag: bool):
        self.on = bool(flag)
        if flag:
            self.events = {}

    class _Region:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            if self.owner.on:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *exc):
            if self.owner.on:
                self.e1.record()
                self.owner.events.setdefault(self.name, []).append((self.e0, self.e1))
            return False

    def region(self, name):
        return self._Region(self, name)

    def report(self, steps: int) -> dict:
        """-> {name: ms per step}; synchronises."""
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) / max(steps, 1) for k, v in self.events.items()}


TIMERS = _Timers()
