"""Plain-SIMT kernels executed on the CPU emulator (tests/emu) through their C ABI: (1) the kernels written against csrc/simt.cuh
(ContextNet / DeepSpeech2 encoder bodies, persistent decode, second-generation hot-path kernels) against numpy references; (2) the
first-generation PRODUCT kernels (RNN-T, CTC, log-mel; sources converted by tests/emu/cu2emu.py) against the oracle, the warp-transducer
known-answer vector and the test.flac golden features; then everything once more under AddressSanitizer and ThreadSanitizer — the CPU
stand-ins for compute-sanitizer memcheck / racecheck.  GPU parity tests of the same entry points: tests/test_gpu_*.py."""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def emu_lib():
    import checks

    return checks.load()


def _case_names():
    import checks

    return [n for n, _ in checks.CASES]


@pytest.mark.parametrize("name", _case_names())
def test_emulated_kernel(emu_lib, name):
    import checks

    dict(checks.CASES)[name](emu_lib)


def _v1_case_names():
    import checks_v1

    return [n for n, _ in checks_v1.CASES]


@pytest.fixture(scope="module")
def emu_lib_v1():
    import checks_v1

    return checks_v1.load()


@pytest.mark.parametrize("name", _v1_case_names())
def test_emulated_product_kernel(emu_lib_v1, name):
    """The first-generation (default-path, GPU-verified) kernels — csrc/rnnt_loss.cu, ctc_loss.cu, logmel.cu, ... converted by
    tests/emu/cu2emu.py — on the CPU emulator against the oracle: warp-transducer's known-answer vector, the test.flac golden features,
    oracle/rnnt.py / ctc.py / logmel.py on ragged batches."""
    import checks_v1

    try:
        dict(checks_v1.CASES)[name](emu_lib_v1)
    except checks_v1.SkipCase as e:
        pytest.skip(str(e))


def test_emulated_kernels_under_sanitizers():
    """AddressSanitizer = the stand-in for compute-sanitizer memcheck: every buffer handed to the emulated kernels is an exact-size heap block,
    so an out-of-bounds global or shared access is a hard error (an injected off-by-one is caught with the kernel's source line).
    ThreadSanitizer = the stand-in for racecheck: every CUDA thread is an OS thread, __syncthreads / __syncwarp / shuffles / cluster.sync /
    mbarriers are real synchronisation and atomics real atomics, so TSan reports exactly the accesses they do NOT order (a deleted
    __syncthreads is reported; on its first run it found an unordered entry-state read in the decode kernel).
    Both sanitizers, both check scripts, side by side, overlapping the rest of the suite (tests/emu/san_runner.py)."""
    import san_runner

    results = san_runner.finish()  # started by tests/conftest.py right after collection (or here, when this test runs on its own)
    if not results:
        pytest.skip("neither libasan nor libtsan available")
    for kind, script, rc, out, err in results:
        assert rc == 0, f"{kind} {script}\n" + out[-2000:] + err[-4000:]
        assert "AddressSanitizer" not in err and "WARNING: ThreadSanitizer" not in err, f"{kind} {script}\n" + err[-6000:]
