"""The C-ABI library loads and exports every symbol include/fabolas_b200.h declares (no compute calls:
this tier has no GPU), and the package refuses to run without the CUDA extension / a GPU."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "fabolas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fab_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from fabolas_b200.build import build
    from fabolas_b200 import _capi
    lib = ctypes.CDLL(build())
    names = _declared()
    assert "fab_gp_loglik_batch" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature in fabolas_b200/_capi.py"
    assert sorted(_capi.SIGNATURES) == names
    assert lib.fab_version() >= 100


def test_no_cpu_fallback():
    import torch
    from fabolas_b200 import Engine, FabolasError, load_library
    with pytest.raises(FabolasError):
        load_library("/nonexistent/libfabolas_b200.so")
    if not torch.cuda.is_available():
        with pytest.raises(FabolasError):
            Engine()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "fabolas_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "cusim" not in text or f in ("prims.cuh", "_capi.py") or "FAB_CUSIM" in text, f
