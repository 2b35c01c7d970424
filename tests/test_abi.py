"""CPU: the C-ABI library loads and exports every symbol include/nefii_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nefii_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nefii_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nefii_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert "nefii_sg_render_fwd" in names and len(names) >= 5
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_python_signature_table_covers_header():
    from nefii_b200 import _lib
    declared = set(_declared_symbols()) - {"nefii_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_version_and_error_text():
    from nefii_b200 import _lib
    assert _lib.raw().nefii_abi_version() >= 1
    assert isinstance(_lib.raw().nefii_last_error(), bytes)


def test_no_cpu_fallback():
    import pytest
    import torch
    from nefii_b200 import _lib
    from nefii_b200.model.sg_render import render_with_sg
    with pytest.raises(_lib.NefiiError):
        render_with_sg(torch.zeros(4, 7), torch.zeros(1, 3), torch.ones(1, 1), torch.zeros(5, 3), torch.zeros(5, 3),
                       torch.zeros(5, 3))
