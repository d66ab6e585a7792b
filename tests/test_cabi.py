"""CPU-side checks of the boundary: the C-ABI library loads and exports every
symbol include/fenapack_cuda.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from fenapack_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fenapack_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fnp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_and_library_exports_every_symbol():
    names = declared_symbols()
    assert len(names) >= 30
    lib = capi.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(capi.SIGNATURES) == names


def test_version_and_error_string():
    lib = capi.load()
    assert b"sm_100a" in lib.fnp_version()
    assert isinstance(lib.fnp_last_error(), bytes)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.FenapackCudaError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)
