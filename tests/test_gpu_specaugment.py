"""GPU: SpecAugment kernel — structural properties of the reference's masking (specaugment.py:58-137); the draws themselves come from a
different random stream than TensorFlow's, so parity is statistical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_specaugment_masks_have_reference_structure():
    from tensorflowasr_b200 import FeatureExtraction, fused

    cfg = {"feature_augment": {"time_masking": {"prob": 1.0, "num_masks": 10, "mask_factor": -1, "p_upperbound": 0.05, "mask_value": 0},
                               "freq_masking": {"prob": 1.0, "num_masks": 1, "mask_factor": 27, "mask_value": 0}}}
    fe = FeatureExtraction(augmentation_config=cfg)
    torch.manual_seed(0)
    B, N = 16, 160000
    sig = (0.1 * torch.randn(B, N, device="cuda")).clamp(-1, 1)
    lens = torch.full((B,), N, dtype=torch.int32, device="cuda")
    lens[1] = N // 2
    clean, L = fe((sig, lens), training=False)
    fused.STATE.seed_tensor(sig.device).fill_(7)
    aug, L2 = fe((sig, lens), training=True)
    assert torch.equal(L, L2)
    changed = (aug != clean)[..., 0]  # [B,T,F]
    assert torch.all(aug[changed.unsqueeze(-1)] == 0)
    for b in range(B):
        ch = changed[b].cpu().numpy()
        full_cols = ch.all(0)   # frequency bins masked for every frame
        full_rows = ch.all(1)   # frames masked for every bin
        # every changed cell belongs to a fully masked column or a fully masked row
        assert np.all(ch <= (full_cols[None, :] | full_rows[:, None]))
        assert full_cols.sum() < 27  # one band narrower than mask_factor
        cols = np.flatnonzero(full_cols)
        assert len(cols) == 0 or cols.max() - cols.min() + 1 == len(cols)  # one contiguous band
        nfr = int(L[b])
        rows = np.flatnonzero(full_rows)
        assert len(rows) <= 10 * int(np.floor(nfr * 0.05)) and (len(rows) == 0 or rows.max() < nfr)  # <= 10 masks of < 5 % of the utterance
    # a different step counter gives different masks; the same counter reproduces them
    fused.STATE.seed_tensor(sig.device).fill_(8)
    aug2, _ = fe((sig, lens), training=True)
    fused.STATE.seed_tensor(sig.device).fill_(7)
    aug3, _ = fe((sig, lens), training=True)
    assert not torch.equal(aug2, aug) and torch.equal(aug3, aug)
    # mean time-mask coverage ~ 10 * E[U{0..49}] / 1000 = 24.5 % before overlaps
    frac = changed[:, :, 40:].all(-1).float().mean().item()
    assert 0.08 < frac < 0.30
