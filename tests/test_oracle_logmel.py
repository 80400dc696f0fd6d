"""CPU: pins the log-mel oracle against independent implementations and the committed golden; checks the host-side
table builder of the C ABI against the oracle's filterbank."""
import ctypes
import os

import numpy as np
import torch

from oracle import logmel


def _pcm(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_testflac.npz"))
    return g, (g["pcm"].astype(np.float64) / 32768.0)[None, :]


def test_oracle_matches_golden(golden_dir):
    g, sig = _pcm(golden_dir)
    feat, flen = logmel.feature_extraction(sig, np.array([sig.shape[1]]), dtype=np.float64)
    assert feat.shape == (1, 894, 80, 1) and flen.tolist() == [894]
    np.testing.assert_allclose(feat.astype(np.float32), g["features"], rtol=0, atol=1e-6)
    # f32 restatement (the dtype TF computes in) stays within the parity tolerance of the f64 truth
    feat32, _ = logmel.feature_extraction(sig.astype(np.float32), np.array([sig.shape[1]]), dtype=np.float32)
    np.testing.assert_allclose(feat32, g["features"], rtol=1e-3, atol=1e-3)


def test_oracle_vs_torch_stft_and_loop_mel(golden_dir):
    """Independent path: torch framing + torch.fft.rfft for the power spectrum; a scalar-loop float64 evaluation of the
    published TF mel formula (triangles linear in HTK-mel, NOT in Hz as torchaudio/librosa do) for the filterbank."""
    g, sig = _pcm(golden_dir)
    x = torch.tensor(sig)
    y = torch.cat([x[:, :1], x[:, 1:] - 0.97 * x[:, :-1]], dim=1)
    T = -(-y.shape[1] // 160)
    need = (T - 1) * 160 + 400
    y = torch.nn.functional.pad(y, (0, need - y.shape[1]))
    win = torch.hann_window(400, periodic=True, dtype=torch.float64)
    frames = y.unfold(1, 400, 160) * win
    spec = (torch.fft.rfft(frames, n=512).abs() ** 2).numpy()
    np.testing.assert_allclose(spec, logmel.stft_power(logmel.preemphasis(sig, 0.97), dtype=np.float64), rtol=1e-9, atol=1e-12)
    import math

    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    edges = [mel(0.0) + i * (mel(8000.0) - mel(0.0)) / 81 for i in range(82)]
    W = np.zeros((257, 80))
    for k in range(1, 257):
        mk = mel(k * 8000.0 / 256)
        for m in range(80):
            lo, ce, hi = edges[m], edges[m + 1], edges[m + 2]
            W[k, m] = max(0.0, min((mk - lo) / (ce - lo), (hi - mk) / (hi - ce)))
    np.testing.assert_allclose(W, logmel.linear_to_mel_weight_matrix(dtype=np.float64), atol=1e-12)
    np.testing.assert_allclose(W, logmel.linear_to_mel_weight_matrix(dtype=np.float32), atol=2e-5)
    ref = np.log(spec @ W + 1e-6)[..., None]
    np.testing.assert_allclose(ref, g["features"].astype(np.float64), rtol=1e-5, atol=1e-5)


def test_oracle_vs_direct_dft():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 1000))
    S = logmel.stft_power(logmel.preemphasis(x, 0.97), dtype=np.float64)
    t = 3
    fr = np.zeros(512)
    y = logmel.preemphasis(x, 0.97)[0]
    seg = y[t * 160 : t * 160 + 400]
    fr[: len(seg)] = seg * logmel.hann_window_periodic(400, np.float64)[: len(seg)]
    k = np.arange(257)[:, None]
    n = np.arange(512)[None, :]
    dft = (fr[None, :] * np.exp(-2j * np.pi * k * n / 512)).sum(1)
    np.testing.assert_allclose(S[0, t], np.abs(dft) ** 2, rtol=1e-9, atol=1e-9)


def test_lengths_and_mask():
    assert logmel.get_nframes(np.array([160000, 159999, 1, 0, 143040])).tolist() == [1000, 1000, 1, 0, 894]
    assert int(logmel.get_nframes(160000, pad_end=False)) == 998
    m = logmel.compute_mask((2, 1600), np.array([1600, 700]))
    assert m.shape == (2, 10) and m.sum(1).tolist() == [10, 5]
    assert logmel.get_signal_chunk_size_and_step(16) == (2800, 2560)


def test_c_abi_table_matches_oracle_filterbank():
    """tfasr_logmel_build_table is host code: its banded mel weights must equal the oracle's dense matrix."""
    from tensorflowasr_b200 import _lib

    lib = _lib.load()
    for F, nfft, sr, lo, hi in [(80, 512, 16000, 0.0, 8000.0), (40, 256, 8000, 125.0, 3800.0), (128, 1024, 16000, 20.0, 7600.0)]:
        fl = 400 if nfft >= 512 else 200
        cfg = _lib.LogmelCfg(sample_rate=sr, frame_length=fl, frame_step=160, nfft=nfft, num_feature_bins=F, feature_type=0,
                             pad_end=1, log_base10=0, padding=0, preemphasis=0.97, epsilon=1e-6, lower_edge_hertz=lo, upper_edge_hertz=hi)
        nbytes = lib.tfasr_logmel_table_bytes(ctypes.byref(cfg))
        assert nbytes > 0
        buf = np.zeros(nbytes // 4, np.float32)
        assert lib.tfasr_logmel_build_table(ctypes.byref(cfg), buf.ctypes.data) == 0
        ibuf = buf.view(np.int32)
        tw_off = (fl + 3) // 4 * 4
        np.testing.assert_allclose(buf[:fl], logmel.hann_window_periodic(fl), atol=1e-7)
        tw = buf[tw_off : tw_off + 2 * nfft].reshape(nfft, 2)
        ang = 2 * np.pi * np.arange(nfft) / nfft
        np.testing.assert_allclose(tw[:, 0], np.cos(ang), atol=1e-7)
        np.testing.assert_allclose(tw[:, 1], -np.sin(ang), atol=1e-7)
        off = tw_off + 2 * nfft
        start, count, woff = ibuf[off : off + F], ibuf[off + F : off + 2 * F], ibuf[off + 2 * F : off + 3 * F]
        w = buf[off + 3 * F :]
        dense = np.zeros((nfft // 2 + 1, F), np.float32)
        for m in range(F):
            dense[start[m] : start[m] + count[m], m] = w[woff[m] : woff[m] + count[m]]
        ref = logmel.linear_to_mel_weight_matrix(F, nfft // 2 + 1, sr, lo, hi, np.float32)
        np.testing.assert_allclose(dense, ref, rtol=0, atol=1e-4)  # fp32 rounding of narrow triangles
    bad = _lib.LogmelCfg(sample_rate=16000, frame_length=400, frame_step=160, nfft=384, num_feature_bins=80)
    assert lib.tfasr_logmel_table_bytes(ctypes.byref(bad)) == 0
