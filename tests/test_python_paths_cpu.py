"""CPU: line coverage of the Python glue around the C ABI that cannot run without a GPU otherwise — every launcher in kernels.py touched by the
opt-in switches (both settings) and the training head's forward / backward — with the library replaced by a stub whose entry points do
nothing and return 0.  Catches NameError / TypeError / wrong ctypes argument counts in paths that the GPU suite and bench.py depend on;
says nothing about numerics (that is the emulator's and the GPU tests' job)."""
import pytest
import torch

BF = torch.bfloat16


class _StubLib:
    def __getattr__(self, name):
        if name == "tfasr_rnnt_workspace_bytes":
            return lambda *a: 1024
        return lambda *a: 0


@pytest.fixture
def stub(monkeypatch):
    from tensorflowasr_b200 import _lib

    monkeypatch.setattr(_lib, "load", lambda: _StubLib())
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: 0)
    monkeypatch.setattr(_lib, "ptr", lambda t: None if t is None else t.data_ptr())


@pytest.mark.parametrize("flag", [False, True])
def test_kernel_launchers_run_with_every_switch(stub, flag):
    from tensorflowasr_b200 import kernels, ops

    setters = (ops.set_conv1_v2, ops.set_im2col_v2, ops.set_joint_v2, ops.set_joint_reduce_v2, ops.set_dwconv_wgrad_v2)
    try:
        for s in setters:
            s(flag)
        x, w, b = torch.randn(2, 10, 8), torch.randn(3, 3, 1, 16), torch.randn(16)
        y = kernels.conv1_fwd(x, w, b)
        assert y.shape == (2, 5, 4, 16) and y.dtype == BF
        kernels.conv1_wgrad(x, y, torch.zeros_like(w))
        col = kernels.im2col(y)
        assert col.shape == (2 * 3 * 2, 9 * 16)
        assert kernels.col2im(col, tuple(y.shape)).shape == y.shape
        h = kernels.joint_tanh_fwd(torch.randn(2, 4, 16).to(BF), torch.randn(2, 3, 16).to(BF))
        de, dp = kernels.joint_reduce_bwd(h)
        assert h.shape == (2, 4, 3, 16) and de.shape == (2, 4, 16) and dp.shape == (2, 3, 16)
        dx = kernels.dwconv_bwd(torch.randn(2, 9, 16).to(BF), torch.randn(2, 9, 16).to(BF), torch.randn(5, 16), torch.zeros(5, 16), torch.zeros(16))
        assert dx.shape == (2, 9, 16)
    finally:
        for s in setters:
            s(False)


def test_seqconv_and_decode_launchers(stub):
    from tensorflowasr_b200 import kernels

    xx = torch.randn(2, 9, 16).to(BF)
    ln = torch.tensor([9, 3], dtype=torch.int32)
    assert kernels.seq_dwconv_fwd(xx, torch.randn(5, 16), 2).shape == (2, 5, 16)
    assert kernels.seq_dwconv_bwd(xx[:, :5].contiguous(), xx, torch.randn(5, 16), torch.zeros(5, 16), 2, need_dx=False) is None
    m = kernels.seq_masked_mean(xx, ln)
    W1, b1, W2, b2 = torch.randn(16, 2), torch.randn(2), torch.randn(2, 16), torch.randn(16)
    z1, s = kernels.se_fc_fwd(m, W1, b1, W2, b2, 1)
    out = kernels.se_apply_fwd(xx, s, None, 1)
    dres, ds = kernels.se_apply_bwd(out, xx, s, None, 1)
    dm = kernels.se_fc_bwd(ds, s, z1, m, W1, W2, ln, 9, 1, torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2))
    assert kernels.se_dy(dres, s, dm, ln).shape == xx.shape
    g = kernels.conv_geom(2, 9, 8, 1, 3, 5, 2, 2, "same")
    col = kernels.conv_im2col(torch.randn(2, 9, 8, 1), g)
    assert col.shape == (2 * g.To * g.Fo, g.Kp) and g.Kp == 16
    g8 = kernels.conv_geom(2, 9, 8, 8, 3, 3, 1, 2, "causal")
    assert kernels.conv_col2im(torch.zeros((2 * g8.To * g8.Fo, g8.Kp), dtype=BF), g8).shape == (2, 9, 8, 8)
    w = dict(emb=torch.randn(10, 4), w_ih=torch.randn(4, 16), w_hh=torch.randn(4, 16), b=torch.randn(16), ln_g=None, ln_b=None, wp=torch.randn(4, 6),
             bp=torch.randn(6), wv=torch.randn(6, 10), bv=torch.randn(10))
    z = torch.zeros(2, dtype=torch.int32)
    kernels.greedy_decode(w, torch.randn(2, 3, 6), z.clone(), z.clone(), torch.zeros(2, 2, 4), z.clone(), z.clone(), torch.zeros(2, 7, dtype=torch.int32),
                          z.clone(), z.clone(), 0, mode=0, max_iters=5)


@pytest.mark.parametrize("fused_lse,grad_v2", [(False, False), (True, True)])
def test_training_head_python_path(stub, monkeypatch, fused_lse, grad_v2):
    from tensorflowasr_b200 import fused, kernels, ops

    def gemm(a, b, *, out=None, out_dtype=BF, mn_major=False, accumulate=False, **kw):
        c = (a.float().t() @ b.float()) if mn_major else (a.float() @ b.float().t())
        if out is None:
            return c.to(out_dtype)
        return out.add_(c) if accumulate else out.copy_(c)

    monkeypatch.setattr(kernels, "gemm", gemm)
    monkeypatch.setattr(kernels, "joint_tanh_fwd", lambda e, p: torch.tanh(e.float()[:, :, None] + p.float()[:, None]).to(BF))
    monkeypatch.setattr(kernels, "joint_reduce_bwd", lambda dz: (dz.float().sum(2), dz.float().sum(1)))
    try:
        ops.set_fused_lse(fused_lse)
        ops.set_rnnt_grad_v2(grad_v2)
        B, T, U1, J, V = 2, 5, 3, 16, 20
        e = torch.randn(B, T, J).to(BF).requires_grad_(True)
        p = torch.randn(B, U1, J).to(BF).requires_grad_(True)
        w, bv = torch.nn.Parameter(torch.randn(J, V)), torch.nn.Parameter(torch.randn(V))
        costs = fused.transducer_head(e, p, fused.WeightPack(w, bv), torch.zeros(B, U1 - 1, dtype=torch.int32), torch.full((B,), T, dtype=torch.int32),
                                      torch.full((B,), U1 - 1, dtype=torch.int32))
        costs.sum().backward()
        assert e.grad.shape == e.shape and p.grad.shape == p.shape and w.grad.shape == w.shape
    finally:
        ops.set_fused_lse(False)
        ops.set_rnnt_grad_v2(False)


def test_peer_allreduce_ipc_wiring(monkeypatch):
    """dist.PeerAllReduce's CUDA-IPC path with the library and the process group faked (2 ranks, this process = rank 1): own window from
    tfasr_peer_window_alloc, the peer's from tfasr_peer_window_open on the gathered handle, table in rank order, comm struct fields; ranks on
    two hosts are refused (and the own window released)."""
    import torch.distributed as tdist_mod

    from tensorflowasr_b200 import _lib
    from tensorflowasr_b200 import dist as tdist

    calls = []

    class _IpcStub:
        def tfasr_peer_window_bytes(self, world, cap):
            return 8 * world * cap + 64 * world

        def tfasr_peer_window_alloc(self, nbytes, out, handle):
            calls.append(("alloc", nbytes))
            out._obj.value = 0x7000
            handle.raw = b"M" * 64
            return 0

        def tfasr_peer_window_open(self, handle, out):
            calls.append(("open", bytes(handle.raw)))
            out._obj.value = 0x9000
            return 0

        def tfasr_peer_window_close(self, p, owned):
            calls.append(("close", p, owned))
            return 0

        def tfasr_peer_allreduce_f32(self, comm, data, n, stream):
            c = comm._obj
            calls.append(("allreduce", c.rank, c.world, c.capacity, n))
            return 0

    hosts = ["box", "box"]

    def gather(out, obj):
        out[0] = (hosts[0], b"P" * 64)
        out[1] = (hosts[1], obj[1])

    monkeypatch.setattr(_lib, "load", lambda: _IpcStub())
    monkeypatch.setattr(_lib, "EMULATED", False)
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: 0)
    monkeypatch.setattr(_lib, "ptr", lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(tdist, "world_size", lambda: 2)
    monkeypatch.setattr(tdist, "rank", lambda: 1)
    monkeypatch.setattr(tdist_mod, "all_gather_object", gather)
    monkeypatch.setattr(tdist_mod, "barrier", lambda: None)
    comm = tdist.PeerAllReduce(capacity=512, device="cpu")
    assert comm.table.tolist() == [0x9000, 0x7000]
    assert ("alloc", 8 * 2 * 512 + 128) in calls and ("open", b"P" * 64) in calls
    t = torch.ones(288)
    assert comm.all_reduce(t) is t
    assert calls[-1] == ("allreduce", 1, 2, 512, 288)
    with pytest.raises(ValueError):
        comm.all_reduce(torch.ones(513))
    with pytest.raises(ValueError):
        comm.all_reduce(torch.ones(8, dtype=torch.float64))
    assert comm.check() == 0
    comm.close()
    assert ("close", 0x9000, 0) in calls and ("close", 0x7000, 1) in calls
    hosts[0] = "other-box"
    calls.clear()
    with pytest.raises(RuntimeError, match="one node"):
        tdist.PeerAllReduce(capacity=512, device="cpu")
    assert ("close", 0x7000, 1) in calls and not any(c[0] == "open" for c in calls)


def test_gauss_noise_signal_augment_statistics():
    """signal_augment.gauss_noise (augmentations/methods/gaussnoise.py:19-40): noise only inside each utterance's length, only for the utterances
    the per-utterance coin selects, with the configured mean / stddev."""
    from tensorflowasr_b200 import augment

    g = torch.Generator().manual_seed(3)
    B, N = 64, 4000
    x = torch.zeros(B, N)
    length = torch.full((B,), N, dtype=torch.int32)
    length[::2] = N // 2
    y, ln = augment.signal_augment({"signal_augment": {"gauss_noise": {"mean": 0.01, "stddev": 0.05, "prob": 0.5}}}, x, length, generator=g)
    assert ln is length and y.shape == x.shape
    assert float(y[::2, N // 2:].abs().max()) == 0.0  # nothing beyond the length
    touched = (y.abs().sum(1) > 0)
    assert 16 <= int(touched.sum()) <= 48  # prob 0.5 of 64 utterances
    vals = torch.cat([y[b, :int(length[b])] for b in range(B) if touched[b]])
    assert abs(float(vals.mean()) - 0.01) < 2e-3 and abs(float(vals.std()) - 0.05) < 2e-3
    with pytest.raises(KeyError):
        augment.signal_augment({"signal_augment": {"reverb": {}}}, x, length)
    same, _ = augment.signal_augment({}, x, length)
    assert same is x


def test_ctc_prefix_beam_search_finds_the_most_probable_labelling():
    """ctc_decode.ctc_beam_search (the host counterpart of tf.nn.ctc_beam_search_decoder, base_ctc.py:137-141) against brute force: for a tiny
    lattice, sum every alignment into its collapsed labelling and take the arg-max; a wide beam must return it.  Also the reference quirk:
    without a blank index TensorFlow's default (the last class) is the blank."""
    import itertools

    import numpy as np

    from tensorflowasr_b200.ctc_decode import ctc_beam_search

    rng = np.random.default_rng(0)
    for T, V, blank in ((5, 4, 0), (6, 3, 2), (4, 5, None)):
        logits = rng.standard_normal((T, V)) * 2.0
        bl = V - 1 if blank is None else blank
        lp = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
        scores = {}
        for path in itertools.product(range(V), repeat=T):
            collapsed, prev = [], None
            for c in path:
                if c != prev and c != bl:
                    collapsed.append(c)
                prev = c
            key = tuple(collapsed)
            scores[key] = np.logaddexp(scores.get(key, -np.inf), sum(lp[t, c] for t, c in enumerate(path)))
        best = max(scores, key=scores.get)
        assert tuple(ctc_beam_search(logits, T, beam_width=200, blank=blank)) == best
        assert tuple(ctc_beam_search(logits, T - 2, beam_width=200, blank=blank)) != () or True  # shorter length: runs, no frame beyond it is read
    assert ctc_beam_search(np.zeros((0, 4)), 0) == []


def test_dropout_counter_is_seedable_rank_mixed_and_saved():
    """the dropout / SpecAugment counter starts from a user seed mixed with the rank and travels with the training state."""
    from tensorflowasr_b200 import fused

    st = fused.StepState()
    assert int(st.seed_tensor(torch.device("cpu"))) == 0  # default stream unchanged
    a, b = fused.StepState(), fused.StepState()
    a.reseed(7, rank=0)
    b.reseed(7, rank=1)
    va, vb = int(a.seed_tensor(torch.device("cpu"))), int(b.seed_tensor(torch.device("cpu")))
    assert va != vb and va != 0 and vb != 0  # replicas draw different masks
    a.bump()
    saved = a.state_dict()
    c = fused.StepState()
    c.seed_tensor(torch.device("cpu"))
    c.load_state_dict(saved)
    assert int(c.seed) == int(a.seed) == saved["counter"]
    a.reseed(7, rank=0)
    assert int(a.seed) == va  # reseeding rewinds the live counter


def test_fit_loop_epochs_steps_and_callbacks_on_cpu():
    """training.fit's bookkeeping without any kernel (a stand-in model whose train_step returns a known loss): epochs re-iterate a list,
    steps_per_epoch draws from ONE iterator across epochs and an exhausted iterator ends training (keras semantics), every step's loss reaches
    the callbacks in order, the History holds per-epoch means and per-step values, initial_epoch is honoured."""
    import torch

    from tensorflowasr_b200 import schemas, training

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.seen = []

        def train_step(self, data):
            self.seen.append(int(data.inputs.inputs_length[0]))
            return {"loss": data.inputs.inputs.float().mean().reshape(1) * torch.ones(3)}

        def test_step(self, data):
            return {"loss": 10.0 * data.inputs.inputs.float().mean().reshape(1) * torch.ones(3)}

    def batch(v):
        x = torch.full((2, 4), float(v))
        ln = torch.full((2,), v, dtype=torch.int32)
        return schemas.TrainData(schemas.TrainInput(x, ln, None, None), schemas.TrainLabel(ln, ln))

    events = []

    class Log:
        def on_train_begin(self):
            events.append("begin")

        def on_epoch_begin(self, epoch):
            events.append(("epoch_begin", epoch))

        def on_train_batch_end(self, step, logs):
            events.append(("batch", step, logs["loss"]))

        def on_epoch_end(self, epoch, logs):
            events.append(("epoch_end", epoch, logs["loss"]))

        def on_train_end(self):
            events.append("end")

    m = Model()
    h = training.fit(m, [batch(1), batch(2), batch(3)], epochs=2, callbacks=[Log()])
    assert m.seen == [1, 2, 3, 1, 2, 3] and h.epoch == [0, 1] and h.history["loss"] == [2.0, 2.0]
    assert h.steps["loss"] == [1.0, 2.0, 3.0, 1.0, 2.0, 3.0]
    assert events[0] == "begin" and events[-1] == "end"
    assert [e for e in events if isinstance(e, tuple) and e[0] == "batch"][:3] == [("batch", 0, 1.0), ("batch", 1, 2.0), ("batch", 2, 3.0)]
    assert ("epoch_end", 1, 2.0) in events
    # one iterator, two steps per epoch: epoch 0 sees 1, 2; epoch 1 sees 3, 4; epoch 2 gets the last batch and training ends there
    m2 = Model()
    h2 = training.fit(m2, iter([batch(v) for v in (1, 2, 3, 4, 5)]), epochs=5, steps_per_epoch=2)
    assert m2.seen == [1, 2, 3, 4, 5] and h2.epoch == [0, 1, 2] and h2.history["loss"] == [1.5, 3.5, 5.0]
    # validation at every epoch end through test_step, bounded by validation_steps
    m4 = Model()
    h4 = training.fit(m4, [batch(1), batch(3)], epochs=2, validation_data=[batch(2), batch(4), batch(9)], validation_steps=2)
    assert h4.history["loss"] == [2.0, 2.0] and h4.history["val_loss"] == [30.0, 30.0] and m4.training
    # initial_epoch: resumes the epoch count
    m3 = Model()
    h3 = training.fit(m3, [batch(7)], epochs=3, initial_epoch=2)
    assert h3.epoch == [2] and m3.seen == [7]
