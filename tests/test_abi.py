"""The C-ABI library loads and exports every symbol include/bigkrls_b200.h declares (CPU)."""
import ctypes
import os
import re

import pytest

from bigkrls_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bigkrls_b200.h")).read()
    return sorted(set(re.findall(r"BK_API [\w \*]+?\b(bk_\w+)\(", src)))


def test_header_symbols_exported():
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_bindings_cover_header():
    assert sorted(_lib._SIGNATURES) == declared_symbols()


def test_version_and_loud_failure_without_gpu():
    lib = _lib.load()
    assert lib.bk_version() == 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BKError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)
