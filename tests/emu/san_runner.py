"""Sanitizer passes of the emulator checks as background subprocesses: AddressSanitizer (memcheck stand-in) and ThreadSanitizer (racecheck
stand-in) builds of the emulator library, each running checks.py and checks_v1.py with --fast.  tests/conftest.py starts them as soon as the
collection shows that tests/test_emu_kernels.py::test_emulated_kernels_under_sanitizers will run, so they overlap with the rest of the CPU
suite; the test itself only waits for them and reads their verdicts.  Output goes to temporary files (a full pipe would block the child)."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
_STATE = {"jobs": None}


def start():
    """-> list of (kind, script, Popen, out_path, err_path); [] when neither sanitizer runtime is installed.  Idempotent."""
    if _STATE["jobs"] is not None:
        return _STATE["jobs"]
    sys.path.insert(0, HERE)
    import build_emu

    specs = []
    asan, tsan = build_emu.asan_runtime(), build_emu.tsan_runtime()
    if asan is not None:
        build_emu.build(asan=True)
        specs += [("asan", script, dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")) for script in ("checks.py", "checks_v1.py")]
    if tsan is not None:
        build_emu.build(tsan=True)
        # numpy's reference matmuls stay on their thread: OpenBLAS's own (uninstrumented) worker threads would show up as races
        env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 halt_on_error=0", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        specs += [("tsan", script, env) for script in ("checks.py", "checks_v1.py")]
    jobs = []
    for kind, script, env in specs:
        out = tempfile.NamedTemporaryFile("w+", prefix=f"tfasr_{kind}_", suffix=".out", delete=False)
        err = tempfile.NamedTemporaryFile("w+", prefix=f"tfasr_{kind}_", suffix=".err", delete=False)
        p = subprocess.Popen([sys.executable, os.path.join(HERE, script), "--" + kind, "--fast"], stdout=out, stderr=err, env=env)
        jobs.append((kind, script, p, out.name, err.name))
        out.close()
        err.close()
    _STATE["jobs"] = jobs
    return jobs


def finish(timeout=2400):
    """Wait for the passes; -> list of (kind, script, returncode, stdout, stderr)."""
    res = []
    for kind, script, p, out, err in start():
        try:
            p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            p.wait()
        res.append((kind, script, p.returncode, open(out).read(), open(err).read()))
        for f in (out, err):
            try:
                os.unlink(f)
            except OSError:
                pass
    _STATE["jobs"] = None
    return res


def abort():
    """Session ended before the test collected the results (a failure under -x, Ctrl-C): stop exactly the children started here."""
    jobs = _STATE["jobs"] or []
    for _, _, p, out, err in jobs:
        if p.poll() is None:
            p.kill()
            p.wait()
        for f in (out, err):
            try:
                os.unlink(f)
            except OSError:
                pass
    _STATE["jobs"] = None
