"""Boundary hygiene: the product never touches the oracle, and the C-ABI library exports nbx.h."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "numbotics_b200")


def _py_files(top):
    for d, _, files in os.walk(top):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(d, f)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/\\]_build|liboracle|nbx_oracle", re.M)
    offenders = [p for p in _py_files(PKG) if pat.search(open(p, encoding="utf-8").read())]
    assert offenders == []


def test_product_has_no_cpu_compute_fallback():
    """The Python layer may marshal arrays but must not restate the arithmetic: no sin/cos chains there."""
    for rel in ("robots/arm.py", "robots/helpers.py", "planning/sampling_based/connectors.py"):
        src = open(os.path.join(PKG, rel), encoding="utf-8").read()
        assert "np.sin" not in src and "np.cos" not in src and "torch.sin" not in src, rel


def test_nothing_reads_the_reference_at_runtime():
    """/root/reference does not exist on the GPU box: only the oracle's loader and the golden generator name it."""
    allowed = {os.path.join(ROOT, "oracle", "ref_loader.py"), os.path.join(ROOT, "tests", "golden", "make_golden.py"),
               os.path.join(ROOT, "tests", "test_oracle_pinned.py"), os.path.join(ROOT, "tests", "test_layout.py")}
    bad = []
    for top in (PKG, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        for p in _py_files(top):
            if p in allowed:
                continue
            if "/root/reference" in open(p, encoding="utf-8").read():
                bad.append(p)
    for p in (os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")):
        if os.path.exists(p) and "/root/reference" in open(p, encoding="utf-8").read():
            bad.append(p)
    assert bad == []


def test_cabi_library_exports_every_declared_symbol():
    from numbotics_b200 import _cabi, build
    build.build()
    header = open(os.path.join(ROOT, "include", "nbx.h"), encoding="utf-8").read()
    declared = set(re.findall(r"NBX_API[^;(]*?\b(nbx_\w+)\s*\(", header))
    assert declared, "no NBX_API declarations found in include/nbx.h"
    assert declared == set(_cabi.SIGNATURES), declared.symmetric_difference(set(_cabi.SIGNATURES))
    lib = _cabi.load()          # resolves every symbol; no compute call is made without a GPU
    for name in declared:
        assert hasattr(lib, name)
    assert lib.nbx_version() == 1
    assert _cabi.launch_count() >= 0


def test_cabi_missing_library_fails_loudly(monkeypatch, tmp_path):
    from numbotics_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.NbxError, match="no CPU fallback"):
        _cabi.load()


def test_compute_without_gpu_raises():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from numbotics_b200 import workloads
    arm = workloads.kinova_arm()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        arm.forward_kinematics(np.zeros(7), "tool_frame")
