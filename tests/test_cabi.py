"""The C-ABI shared library loads and exports every symbol include/buglab_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "buglab_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from buglab_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_library_exports_every_declared_symbol():
    from buglab_b200 import _lib

    lib = _lib.load()  # raises if the .so is missing or a declared symbol is not exported
    raw = ctypes.CDLL(_lib.library_path())
    for name in _declared_symbols():
        assert getattr(raw, name) is not None
    assert lib.bl_version() >= 100
    assert b"workspace" in lib.bl_error_string(-3)
    assert lib.bl_plan_workspace_bytes(1000, 100, 3) > 0  # pure host arithmetic, no GPU needed


def test_no_cpu_fallback():
    import pytest
    import torch

    from buglab_b200 import _lib, ops

    with pytest.raises(_lib.BuglabB200Error):
        ops.layer_norm(torch.randn(4, 8), torch.ones(8), torch.zeros(8))
    with pytest.raises(_lib.BuglabB200Error):
        z = torch.zeros(3, dtype=torch.int64)
        ops.build_edge_plan([(z, z)], 4)
