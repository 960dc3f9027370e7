"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/caelo.h
declares (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from caelo_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "caelo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(caelo_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from caelo_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "no ctypes prototype for " + n
    assert set(_lib.SIGNATURES) == set(names)


def test_error_strings_and_version(lib):
    assert lib.caelo_version() >= 100
    assert lib.caelo_error_string(0) == b"ok"
    assert b"0 or 1" in lib.caelo_error_string(-4)
    assert b"no CPU fallback" in lib.caelo_error_string(-6)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under caelo_b200/ may reference it."""
    pkg = os.path.join(ROOT, "caelo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libcaelo_oracle" not in text, f


def test_create_fails_loudly_without_gpu(lib):
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert lib.caelo_create(0, ctypes.byref(h)) == -6
    from caelo_b200 import _lib, api
    with pytest.raises(_lib.CaeloError):
        api.Context(0)
