"""CPU: pins the RNN-T / CTC oracles (warp-transducer known-answer vector, torchaudio, torch) and the mask oracle
against the reference's own golden matrices."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ctc, masks, rnnt


def test_rnnt_known_answer_vector():
    k = rnnt.WARP_TRANSDUCER_KAT
    loss, grads = rnnt.rnnt_loss(k["acts"], k["labels"], [2], [2])
    assert abs(loss[0] - k["cost"]) < 1e-5
    np.testing.assert_allclose(grads, k["grads"], atol=1e-6)


@pytest.mark.parametrize("seed", [0, 1])
def test_rnnt_vs_torchaudio(seed):
    import torchaudio

    rng = np.random.default_rng(seed)
    B, T, U, V = 4, 13, 7, 11
    lg = rng.standard_normal((B, T, U + 1, V)) * 2
    lab = rng.integers(1, V, (B, U)).astype(np.int32)
    tl = np.array([13, 9, 5, 13], np.int32)
    ul = np.array([7, 4, 0, 1], np.int32)
    loss, g = rnnt.rnnt_loss(lg, lab, tl, ul)
    x = torch.tensor(lg, dtype=torch.float32, requires_grad=True)
    l2 = torchaudio.functional.rnnt_loss(x, torch.tensor(lab), torch.tensor(tl), torch.tensor(ul), blank=0, reduction="none")
    l2.sum().backward()
    np.testing.assert_allclose(loss, l2.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(g, x.grad.numpy(), atol=2e-5)
    # rows outside the lattice carry no gradient; every row's gradient sums to zero (softmax Jacobian)
    assert np.abs(g[1, 9:]).max() == 0 and np.abs(g[1, :, 5:]).max() == 0
    assert np.abs(g.sum(-1)).max() < 1e-9


def test_rnnt_length_clamp():
    """base_loss.py:36: logit_length < label_length is raised to label_length."""
    tl, ul = rnnt.clamp_lengths([3, 10], [5, 4])
    assert tl.tolist() == [5, 10] and ul.tolist() == [5, 4]


def test_ctc_vs_torch():
    rng = np.random.default_rng(5)
    B, T, V, L = 5, 40, 17, 9
    lg = rng.standard_normal((B, T, V)) * 1.5
    lab = rng.integers(1, V, (B, L)).astype(np.int32)
    lab[0, :5] = [4, 4, 4, 7, 7]
    tl = np.array([40, 31, 20, 40, 12], np.int32)
    ul = np.array([9, 6, 0, 1, 9], np.int32)
    loss, g = ctc.ctc_loss(lg, lab, tl, ul)
    x = torch.tensor(lg, requires_grad=True)
    l2 = torch.nn.functional.ctc_loss(torch.log_softmax(x, -1).transpose(0, 1), torch.tensor(lab, dtype=torch.long),
                                      torch.tensor(tl, dtype=torch.long), torch.tensor(ul, dtype=torch.long), blank=0, reduction="none")
    l2.sum().backward()
    np.testing.assert_allclose(loss, l2.detach().numpy(), rtol=1e-9)
    np.testing.assert_allclose(g, x.grad.numpy(), atol=1e-9)


def test_ctc_infeasible():
    lg = np.zeros((1, 3, 5))
    loss, g = ctc.ctc_loss_and_grad(lg, np.array([[1, 1, 1, 1]]), [3], [4])
    assert np.isinf(loss[0]) and np.abs(g).max() == 0


def test_streaming_mask_reference_golden(golden_dir):
    for case in json.load(open(os.path.join(golden_dir, "streaming_mask.json"))):
        t = case["query_shape"][1]
        got = masks.compute_streaming_mask(case["chunk_size"], case["history_size"], t)
        assert got.tolist() == case["mask"]


def test_rel_shift_docstring_and_closed_form(golden_dir):
    d = json.load(open(os.path.join(golden_dir, "relshift_docstring.json")))
    x = np.array(d["input"])[None, None]
    got = masks.rel_left_shift(x, causal=True)[0, 0]
    # The docstring shows the shift with the (meaningless) upper triangle blanked; the code as written leaves the
    # classic Transformer-XL wrap-around values there ([3, 0, 4] in row 0) which the causal mask later removes.
    # The pinned part is the lower triangle.
    assert np.tril(got).tolist() == np.tril(np.array(d["causal_output"])).tolist()
    rng = np.random.default_rng(0)
    for T, M in [(5, 0), (7, 3), (1, 0)]:
        R = 2 * T + M - 1
        x = rng.standard_normal((2, 3, T, R))
        lit = masks.rel_left_shift(x)
        S = T + M
        assert lit.shape[-1] == S
        np.testing.assert_array_equal(lit[..., -S:], masks.rel_shift_closed_form(x, S))
