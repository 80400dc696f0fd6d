"""CPU: Transducer.recognize_batch / recognize_single through the persistent decode path (two-pass host logic of
transducer._recognize_batch_persistent + the REAL kernel source on the CPU emulator) against the host-driven loops — the verbatim mirror of
base_transducer.py:496-712 — on the same random model.  Tokens, next tokens and next decoder states must agree."""
import shutil

import pytest
import torch

import mock_kernels

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available (emulator build)")


def _model(V=16, seed=0):
    import tensorflowasr_b200 as tfasr

    cfg = tfasr.conformer_small_config(vocab_size=V)
    cfg["speech_config"].pop("augmentation_config", None)
    cfg.update(encoder_num_blocks=1, encoder_dmodel=16, encoder_head_size=4, encoder_num_heads=4, prediction_embed_dim=12, prediction_rnn_units=16,
               joint_dim=20)
    cfg["encoder_subsampling"]["config"]["filters"] = [16, 16]
    return tfasr.Conformer(**cfg, seed=seed)


# (seed, blank bias) draws whose host loop terminates; (2, 2.0) ends through the token budget (every utterance emits on every iteration):
# there the persistent path must notice and hand over to the host loop
@pytest.mark.parametrize("seed,blank_bias,min_emitted", [(5, 3.0, 2), (7, 4.0, 3), (4, 4.0, 1), (2, 2.0, 40)])
def test_persistent_decode_host_logic_matches_host_loop(monkeypatch, seed, blank_bias, min_emitted):
    from tensorflowasr_b200 import ops, schemas

    ops.set_precision("fp32")
    mock_kernels.install(monkeypatch)
    model = _model(seed=seed)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "joint" in n or "prediction" in n:
                p.add_(0.3 * torch.randn(p.shape, generator=g))
        model.joint_net["vocab.bias"][0] += blank_bias
    B, T, D = 3, 9, 16
    enc = torch.randn(B, T, D, generator=g)
    enc_len = torch.tensor([9, 5, 1])

    def fake_encode(inputs):
        b = inputs.inputs.shape[0]
        return enc[:b], enc_len[:b].to(torch.int64), [], model.get_initial_tokens(b), model.get_initial_decoder_states(b)

    monkeypatch.setattr(model, "_encode_for_decode", fake_encode)
    calls, orig_call_next = [0], model.call_next

    def guarded_call_next(*a):  # a random model may never emit blank: fail instead of looping forever
        calls[0] += 1
        assert calls[0] < 2000, "host greedy loop does not terminate on this draw"
        return orig_call_next(*a)

    monkeypatch.setattr(model, "call_next", guarded_call_next)
    dummy = schemas.PredictInput(torch.zeros(B, 1), torch.ones(B, dtype=torch.int32))
    dummy1 = schemas.PredictInput(torch.zeros(1, 1), torch.ones(1, dtype=torch.int32))
    out = {}
    try:
        for flag in (False, True):
            ops.set_persistent_decode(flag)
            out[flag] = (model.recognize(dummy), model.recognize(dummy1, max_tokens_per_frame=2))
    finally:
        ops.set_persistent_decode(False)
    emitted = int((out[False][0].tokens[:, 2:] != 0).sum())
    for k in (0, 1):
        a, b = out[False][k], out[True][k]
        assert a.tokens.shape == b.tokens.shape and a.tokens.tolist() == b.tokens.tolist(), (a.tokens, b.tokens)
        assert a.next_tokens.flatten().tolist() == b.next_tokens.flatten().tolist()
        assert a.next_decoder_states.shape == b.next_decoder_states.shape
        torch.testing.assert_close(a.next_decoder_states.float(), b.next_decoder_states.float(), rtol=1e-4, atol=1e-5)
    assert emitted >= min_emitted, f"degenerate draw: {emitted} symbols emitted"
