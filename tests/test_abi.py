"""CPU: the C-ABI library loads and exports every symbol include/tfasr_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from tensorflowasr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tfasr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfasr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tfasr_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= set(names)


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.tfasr_version() >= 100
    assert lib.tfasr_status_string(0) == b"ok"
    assert b"workspace" in lib.tfasr_status_string(3)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any launch."""
    lib = _lib.load()
    assert lib.tfasr_rnnt_workspace_bytes(2, 10, 5, 16) >= 5 * 4 * 2 * 10 * 5
    # blank != 0 -> invalid value (base_loss.py:24); NULL pointers -> invalid value
    assert lib.tfasr_rnnt_loss_f32(None, None, None, None, 1, 4, 3, 8, 0, None, None, None, 0, None) == 1
    assert lib.tfasr_ctc_loss_f32(None, None, None, None, 1, 4, 3, 8, 1, None, None, None, None, 0, None) == 1
    with pytest.raises(ValueError):
        _lib.check(1, "x")
    with pytest.raises(RuntimeError):
        _lib.check(2, "x")


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensorflowasr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_no_cpu_fallback():
    import torch

    from tensorflowasr_b200 import FeatureExtraction, rnnt_loss

    with pytest.raises(RuntimeError):
        FeatureExtraction()((torch.zeros(1, 1600), torch.tensor([1600])))
    with pytest.raises(RuntimeError):
        rnnt_loss(torch.zeros(1, 2, 3, 5), torch.tensor([2]), torch.zeros(1, 2, dtype=torch.int32), torch.tensor([2]))


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/tfasr_b200.h as gcc lays it out == the ctypes mirrors in _lib.py."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"tfasr_logmel_cfg": _lib.LogmelCfg, "tfasr_gemm_epilogue": _lib.GemmEpilogue, "tfasr_conv_geom": _lib.ConvGeom,
               "tfasr_greedy_params": _lib.GreedyParams, "tfasr_peer_comm": _lib.PeerComm}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tfasr_b200.h"', "int main(void) {"]
    for cname, mirror in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, mirror in structs.items():
        assert int(got[cname]) == ctypes.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(mirror, fname).offset, f"{cname}.{fname}"
