"""Mints the committed golden fixtures.  Runs ONLY in the build container (it reads
/root/reference, which does not exist on the GPU box); the outputs are committed.

    python tests/golden/make_golden.py

Fixtures:
  * logmel_testflac.npz   — PCM of the reference's tests/test.flac (decoded by oracle/flac.py, MD5-verified
                            against STREAMINFO) and the oracle log-mel of it (float64 arithmetic, stored f32),
                            with the default FeatureExtraction arguments (feature_extraction.py:36-53), plus
                            the tests/test_layers.py:37 variant (16 000 zero samples appended, length unchanged).
  * streaming_mask.json   — the two golden matrices asserted in the reference's tests/test_mask.py:6-55,
                            extracted from that file's AST (the test itself needs TensorFlow).
  * relshift_docstring.json — the causal example in multihead_attention.py:31-41.
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference"

from oracle import flac, logmel  # noqa: E402


def main():
    pcm, sr, bps = flac.decode_flac(os.path.join(REF, "tests", "test.flac"))
    assert sr == 16000 and bps == 16 and pcm.shape[1] == 1
    pcm16 = pcm[:, 0].astype(np.int16)
    sig = (pcm16.astype(np.float64) / 32768.0)[None, :]  # data_util.read_raw_audio: 16-bit PCM / 32768
    n = np.array([sig.shape[1]], np.int32)
    feat, flen = logmel.feature_extraction(sig, n, dtype=np.float64)
    sig_pad = np.concatenate([sig, np.zeros((1, 16000))], axis=1)
    feat_pad, flen_pad = logmel.feature_extraction(sig_pad, n, dtype=np.float64)
    spec, _ = logmel.feature_extraction(sig, n, feature_type="spectrogram", num_feature_bins=160, dtype=np.float64)
    np.savez_compressed(
        os.path.join(HERE, "logmel_testflac.npz"),
        pcm=pcm16, features=feat.astype(np.float32), features_length=flen,
        features_padded=feat_pad.astype(np.float32), features_padded_length=flen_pad,
        spectrogram160=spec.astype(np.float32),
    )
    print("logmel golden", feat.shape, flen, feat_pad.shape, flen_pad)

    tree = ast.parse(open(os.path.join(REF, "tests", "test_mask.py")).read())
    calls, consts = [], []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "compute_streaming_mask":
            calls.append([ast.literal_eval(node.args[0]), ast.literal_eval(node.args[1]),
                          ast.literal_eval(node.args[2].args[0])])
        if isinstance(node, ast.Call) and getattr(node.func, "attr", None) == "constant":
            consts.append(ast.literal_eval(node.args[0]))
    calls.sort(key=lambda c: c[2][1])
    consts.sort(key=lambda m: len(m[0]))
    out = [dict(chunk_size=c[0], history_size=c[1], query_shape=c[2], mask=m) for c, m in zip(calls, consts)]
    json.dump(out, open(os.path.join(HERE, "streaming_mask.json"), "w"))
    print("streaming masks", [(o["chunk_size"], o["history_size"], o["query_shape"]) for o in out])

    json.dump(dict(input=[[1, 2, 3], [4, 5, 6], [7, 8, 9]], causal_output=[[3, 0, 0], [5, 6, 0], [7, 8, 9]]),
              open(os.path.join(HERE, "relshift_docstring.json"), "w"))


if __name__ == "__main__":
    main()
