"""The plain-SIMT kernels of csrc/seqconv_kernels.cuh (ContextNet / DeepSpeech2 encoder bodies) executed on the
CPU emulator (tests/emu) through their C ABI and checked against numpy references; then once more under AddressSanitizer, the CPU
stand-in for compute-sanitizer memcheck.  The GPU parity tests of the same entry points are in tests/test_gpu_seqconv.py."""
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


def test_emulated_kernels_under_address_sanitizer():
    import build_emu

    rt = build_emu.asan_runtime()
    if rt is None:
        pytest.skip("libasan not available")
    build_emu.build(asan=True)
    env = dict(os.environ, LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "checks.py"), "--asan"], capture_output=True, text=True, env=env, timeout=1500)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-4000:]


def test_emulated_kernels_under_thread_sanitizer():
    """Every CUDA thread is an OS thread in the emulator, __syncthreads / __syncwarp / shuffles are real barriers and atomics real atomics:
    ThreadSanitizer then reports exactly the accesses that are NOT ordered by them — the CPU stand-in for compute-sanitizer racecheck
    (a deleted __syncthreads is reported; it found an unordered entry-state read in the decode kernel when first run)."""
    import build_emu

    rt = build_emu.tsan_runtime()
    if rt is None:
        pytest.skip("libtsan not available")
    build_emu.build(tsan=True)
    env = dict(os.environ, LD_PRELOAD=rt, TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 halt_on_error=0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "checks.py"), "--tsan"], capture_output=True, text=True, env=env, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-6000:]

