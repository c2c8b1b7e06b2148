"""CPU checks of the drop-in boundary: the library loads and exports every symbol the header declares."""

import ctypes
import os
import re

from qibochem_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "qibochem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native_lib):
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in the header but not exported"


def test_binding_table_matches_header(native_lib):
    assert sorted(_native.SIGNATURES) == _header_symbols()


def test_version_and_error_channel(native_lib):
    assert native_lib.qc_version() >= 100
    # argument validation happens before any CUDA call: usable without a GPU
    rc = native_lib.qc_create(0, 0, None, None)
    assert rc != 0
    assert b"qc_create" in native_lib.qc_last_error()


def test_no_cpu_fallback_without_gpu(native_lib):
    """Without a device the product path must fail loudly, never compute on the CPU."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qibochem_b200.engine import DeviceState

    with pytest.raises(RuntimeError):
        DeviceState(4)
