"""GPU parity of the persistent greedy decode kernel (csrc/decode.cu): tokens bit-exact against the host-driven loops
(the verbatim mirror of base_transducer.py:496-712) and against the float64 oracle, fp32 mode.

First run on a B200 in round 2: green, and the persistent kernel became the default decode route (54.5 vs 750.7 us per iteration at batch 32).
"""
import numpy as np
import pytest
import torch

from oracle import conformer as oc
from oracle import logmel as ologmel

pytestmark = pytest.mark.gpu


def test_persistent_greedy_decode_matches_host_loop_and_oracle():
    from tensorflowasr_b200 import ops, schemas

    from test_gpu_model import _batch, _setup

    model, ocfg, params = _setup(V=32, seed=5, blank_bias=-0.6)
    data, host = _batch(V=32, B=3, N=16000, ragged=True)
    sig, sig_len = host[0], host[1]
    model.eval()
    res = {}
    try:
        for flag in (False, True):
            ops.set_persistent_decode(flag)
            res[flag] = (model.recognize(schemas.PredictInput(sig.cuda(), sig_len.cuda())),
                         model.recognize(schemas.PredictInput(sig[:1].cuda(), sig_len[:1].cuda())))
    finally:
        ops.set_persistent_decode(False)
    for k in (0, 1):
        a, b = res[False][k], res[True][k]
        assert a.tokens.shape == b.tokens.shape
        assert a.tokens.cpu().tolist() == b.tokens.cpu().tolist()
        assert a.next_tokens.cpu().flatten().tolist() == b.next_tokens.cpu().flatten().tolist()
        torch.testing.assert_close(a.next_decoder_states.float(), b.next_decoder_states.float(), rtol=1e-4, atol=1e-5)
    feats, flen = ologmel.feature_extraction(sig.numpy().astype(np.float64), sig_len.numpy(), dtype=np.float64)
    p = {k: v.double() for k, v in params.items()}
    enc, elen = oc.conformer_encoder(torch.from_numpy(feats), torch.from_numpy(flen).long(), p, ocfg, training=False)
    ref, margin = oc.greedy_batch(enc, elen, p, ocfg)
    if margin > 1e-3:
        assert res[True][0].tokens.cpu().long().tolist() == ref.tolist()


def test_persistent_decode_at_conformer_s_sizes():
    """E = R = J = 320, V = 1000 (small.yml.j2): shared memory opt-in, 1024-thread CTA, split-K shapes of the real model."""
    from tensorflowasr_b200 import kernels

    torch.manual_seed(3)
    B, T, E, R, J, V = 4, 40, 320, 320, 320, 1000
    dev = "cuda"
    w = dict(emb=torch.randn(V, E) * 0.5, w_ih=torch.randn(E, 4 * R) / E ** 0.5, w_hh=torch.randn(R, 4 * R) / R ** 0.5, b=torch.randn(4 * R) * 0.1,
             ln_g=1 + 0.1 * torch.randn(R), ln_b=0.1 * torch.randn(R), wp=torch.randn(R, J) / R ** 0.5, bp=torch.randn(J) * 0.1,
             wv=torch.randn(J, V) * (2.0 / J ** 0.5), bv=torch.randn(V) * 0.1)
    w["bv"][0] += 4.0
    e_proj = torch.randn(B, T, J)
    nframes = torch.tensor([40, 33, 7, 1], dtype=torch.int32)

    def step(b, t, tok, h, c):
        z = w["emb"][tok].double() @ w["w_ih"].double() + h @ w["w_hh"].double() + w["b"].double()
        c2 = torch.sigmoid(z[R:2 * R]) * c + torch.sigmoid(z[:R]) * torch.tanh(z[2 * R:3 * R])
        h2 = torch.sigmoid(z[3 * R:]) * torch.tanh(c2)
        y = (h2 - h2.mean()) / torch.sqrt(h2.var(unbiased=False) + 1e-3) * w["ln_g"].double() + w["ln_b"].double()
        logits = torch.tanh(e_proj[b, t].double() + y @ w["wp"].double() + w["bp"].double()) @ w["wv"].double() + w["bv"].double()
        top = logits.topk(2).values
        return int(logits.argmax()), h2, c2, float(top[0] - top[1])

    # mode 2 (recognize_single semantics) per utterance against the float64 walk
    cap = 3
    for b in range(B):
        n = int(nframes[b])
        toks_ref, prev, h, c, frame, per, margin = [], 0, torch.zeros(R, dtype=torch.float64), torch.zeros(R, dtype=torch.float64), 0, 0, 1e9
        while frame < n:
            cur, h2, c2, m = step(b, frame, prev, h, c)
            margin = min(margin, m)
            if cur == 0:
                frame, per = frame + 1, 0
            else:
                per += 1
                toks_ref.append(cur)
                prev, h, c = cur, h2, c2
                if per >= cap:
                    frame, per = frame + 1, 0
        wd = {k: v.to(dev).contiguous() for k, v in w.items()}
        toks = torch.zeros(1, max(n * cap, 1), dtype=torch.int32, device=dev)
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        st = torch.zeros(1, 2, R, device=dev)
        pt = z.clone()
        kernels.greedy_decode(wd, e_proj[b:b + 1].to(dev).contiguous(), nframes[b:b + 1].to(dev), pt, st, z.clone(), torch.full_like(z, -1), toks, z.clone(),
                              z.clone(), 0, mode=2, per_frame_cap=cap, threads=1024)
        if margin > 1e-3:
            got = toks[0].cpu().tolist()
            assert got[: len(toks_ref)] == toks_ref and all(t == 0 for t in got[len(toks_ref):])
            torch.testing.assert_close(st[0, 0].cpu().double(), h, rtol=1e-3, atol=1e-4)
