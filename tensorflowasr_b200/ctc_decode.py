"""CTC prefix beam search on the host — the counterpart of `tf.nn.ctc_beam_search_decoder`, which the reference calls for
`CtcModel.recognize_beam` (tensorflow_asr/models/ctc/base_ctc.py:128-160) and which TensorFlow itself runs as a CPU op
(core/util/ctc/ctc_beam_search.h, tensorflow~=2.18, absent here: restated from its published algorithm — log-softmax of the inputs per
frame, prefix beams carrying (log p ending in blank, log p ending in a label), top `beam_width` prefixes kept per frame, best prefix
returned).  Decode path, not the training hot path: plain numpy, float64.

Reference quirk kept on purpose: base_ctc.py:137-141 passes no blank index, so TensorFlow uses its default — the LAST class — although the
models train with blank = 0 (the greedy path passes blank_index=self.blank).  `blank=None` reproduces that call; pass the model's blank to
decode with the index the model was trained with."""
from __future__ import annotations

import numpy as np

NEG = -np.inf


def ctc_beam_search(logits: np.ndarray, length: int, beam_width: int = 10, blank: int | None = None) -> list[int]:
    """logits [T, V] (raw, any scale) of ONE utterance, `length` valid frames -> the most probable label sequence (blank-free, repeats merged
    by the CTC collapse)."""
    T, V = logits.shape
    blank = V - 1 if blank is None else int(blank)
    x = np.asarray(logits[: max(0, min(int(length), T))], np.float64)
    lp = x - x.max(-1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(-1, keepdims=True))
    beams = {(): (0.0, NEG)}  # prefix -> (log p of alignments ending in blank, ... ending in the prefix's last label)
    labels = [c for c in range(V) if c != blank]
    for t in range(lp.shape[0]):
        row = lp[t]
        nxt: dict = {}

        def add(prefix, pb, pnb):
            o = nxt.get(prefix)
            nxt[prefix] = (pb, pnb) if o is None else (np.logaddexp(o[0], pb), np.logaddexp(o[1], pnb))

        for prefix, (pb, pnb) in beams.items():
            total = np.logaddexp(pb, pnb)
            add(prefix, total + row[blank], NEG)  # stay on the prefix through a blank
            if prefix:
                add(prefix, NEG, pnb + row[prefix[-1]])  # ... or by repeating its last label
            for c in labels:
                # a repeated label only extends the prefix across a blank
                add(prefix + (c,), NEG, (pb if (prefix and c == prefix[-1]) else total) + row[c])
        beams = dict(sorted(nxt.items(), key=lambda kv: -np.logaddexp(kv[1][0], kv[1][1]))[:beam_width])
    best = max(beams.items(), key=lambda kv: np.logaddexp(kv[1][0], kv[1][1]))[0]
    return list(best)
