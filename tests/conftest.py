import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_collection_finish(session):
    """The sanitizer passes of the emulator checks take ~2 minutes of mostly idle waiting on four subprocesses: start them as soon as it is
    known that their test will run (never on a `-m gpu` run), so that they overlap with the rest of the CPU suite."""
    if hasattr(session.config, "workerinput"):  # pytest-xdist worker: the test starts its own passes when it runs
        return
    if any(item.nodeid.endswith("test_emulated_kernels_under_sanitizers") and not any(m.name == "skip" for m in item.iter_markers())
           for item in session.items):
        import shutil

        if shutil.which("g++") is not None:
            sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
            import san_runner

            san_runner.start()


def pytest_sessionfinish(session, exitstatus):
    if "san_runner" in sys.modules:
        sys.modules["san_runner"].abort()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _kernel_route_switches_restored():
    """Tests flip the kernel-route switches of tensorflowasr_b200.ops (and the precision); every test starts from, and leaves behind, the
    defaults — so each test runs the route a user gets unless it says otherwise."""
    from tensorflowasr_b200 import ops

    saved, prec = ops.switch_state(), ops.precision()
    yield
    ops.restore_switches(saved)
    if ops.precision() != prec:
        ops.set_precision(prec)
