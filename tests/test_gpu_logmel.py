"""GPU parity: fused log-mel kernel (through the C ABI) vs the CPU oracle and the committed golden.
Tolerance (north_star): log-mel within 1e-3 relative in fp32 -> assert_allclose(rtol=1e-3, atol=1e-3) on values in [-14, 12]."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-3


def _run(sig, lens, **kw):
    from tensorflowasr_b200 import FeatureExtraction

    fe = FeatureExtraction(**kw)
    f, l = fe((torch.tensor(sig, dtype=torch.float32, device="cuda"), torch.tensor(lens, dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()
    return f.cpu().numpy(), l.cpu().numpy()


def _oracle_kw(kw):
    m = dict(kw)
    if "preemphasis" in m:
        m["preemphasis_coef"] = m.pop("preemphasis")
    return m


def test_golden_test_flac(golden_dir):
    """BASELINE config 1: SpeechFeaturizer 80-dim log-mel on the reference's tests/test.flac."""
    g = np.load(os.path.join(golden_dir, "logmel_testflac.npz"))
    sig = (g["pcm"].astype(np.float32) / 32768.0)[None, :]
    f, l = _run(sig, [sig.shape[1]])
    assert f.shape == (1, 894, 80, 1) and l.tolist() == [894]
    np.testing.assert_allclose(f, g["features"], rtol=RTOL, atol=ATOL)
    err = np.abs(f - g["features"]).max()
    print("max abs err vs f64 oracle golden:", err)
    assert err < 2e-4  # well inside the gate: fp32 FFT noise only
    # tests/test_layers.py:37 variant: 16000 zero samples appended, length unchanged
    sig_pad = np.concatenate([sig, np.zeros((1, 16000), np.float32)], axis=1)
    f2, l2 = _run(sig_pad, [sig.shape[1]])
    assert f2.shape == (1, 994, 80, 1) and l2.tolist() == [894]
    np.testing.assert_allclose(f2, g["features_padded"], rtol=RTOL, atol=ATOL)


def test_spectrogram_feature_type(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_testflac.npz"))
    sig = (g["pcm"].astype(np.float32) / 32768.0)[None, :]
    f, _ = _run(sig, [sig.shape[1]], feature_type="spectrogram", num_feature_bins=160)
    np.testing.assert_allclose(f, g["spectrogram160"], rtol=RTOL, atol=2e-3)


@pytest.mark.parametrize(
    "kw",
    [
        dict(),
        dict(preemphasis=0.0),
        dict(log_base="10"),
        dict(nfft=1024, frame_ms=50, num_feature_bins=64),
        dict(nfft=256, frame_ms=16, stride_ms=8, num_feature_bins=40, sample_rate=8000, upper_edge_hertz=4000.0),
        dict(pad_end=False),
        dict(padding=480),
        dict(lower_edge_hertz=125.0, upper_edge_hertz=7600.0, num_feature_bins=23),
        dict(normalize_signal=True),
        dict(normalize_zscore=True),
        dict(normalize_min_max=True),
    ],
)
def test_random_batches_vs_oracle(kw):
    rng = np.random.default_rng(1234)
    for B, N in [(3, 16000), (2, 4801), (1, 401), (5, 2560)]:
        sig = (0.1 * rng.standard_normal((B, N))).clip(-1, 1).astype(np.float32)
        lens = rng.integers(N // 2, N + 1, B).astype(np.int32)
        lens[0] = N
        f, l = _run(sig, lens, **kw)
        ref, rl = logmel.feature_extraction(sig.astype(np.float64), lens, dtype=np.float64, **_oracle_kw(kw))
        assert f.shape == ref.shape, (f.shape, ref.shape)
        assert l.tolist() == rl.tolist()
        tol = 5e-3 if (kw.get("normalize_zscore") or kw.get("normalize_min_max")) else ATOL
        np.testing.assert_allclose(f, ref, rtol=RTOL, atol=tol)


def test_edge_cases():
    # silence -> log(eps) everywhere; a single sample; unaligned N (scalar load path); frame count boundaries
    f, l = _run(np.zeros((2, 3200), np.float32), [3200, 0])
    assert np.allclose(f, np.log(1e-6), atol=1e-5) and l.tolist() == [20, 0]
    for N in (1, 159, 160, 161, 399, 400, 2799, 2800, 2801, 5003):
        rng = np.random.default_rng(N)
        sig = rng.standard_normal((2, N)).astype(np.float32) * 0.3
        f, l = _run(sig, [N, max(N - 1, 0)])
        ref, rl = logmel.feature_extraction(sig.astype(np.float64), np.array([N, max(N - 1, 0)]), dtype=np.float64)
        assert f.shape == ref.shape and l.tolist() == rl.tolist()
        np.testing.assert_allclose(f, ref, rtol=RTOL, atol=ATOL)


def test_full_size_properties():
    """BASELINE size (B=32 x 10 s): chunk-invariance property (feature_extraction.py:141-162: features of a chunk equal
    the same frames of the whole signal) + bf16 input path + length outputs."""
    from tensorflowasr_b200 import FeatureExtraction

    torch.manual_seed(1234)
    fe = FeatureExtraction()
    sig = (0.1 * torch.randn(32, 160000, device="cuda")).clamp(-1, 1)
    lens = torch.full((32,), 160000, dtype=torch.int32, device="cuda")
    f, l = fe((sig, lens))
    assert f.shape == (32, 1000, 80, 1) and l.tolist() == [1000] * 32
    assert torch.isfinite(f).all()
    size, step = fe.get_signal_chunk_size_and_step(16)
    # chunk starting at frame 32 (sample 32*160); pre-emphasis needs the previous sample -> compare frames >= 1 of the chunk
    c0 = 32 * 160
    fc, _ = fe((sig[:, c0 - 160 : c0 - 160 + size + 160].contiguous(), lens))
    torch.testing.assert_close(fc[:, 1:17], f[:, 32:48], rtol=1e-4, atol=1e-4)
    # spot-check 2 utterances against the oracle
    ref, _ = logmel.feature_extraction(sig[:2].double().cpu().numpy(), np.array([160000, 160000]), dtype=np.float64)
    np.testing.assert_allclose(f[:2].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    fb, _ = fe((sig.to(torch.bfloat16), lens))
    assert fb.dtype == torch.bfloat16 and fb.shape == f.shape
