"""The C-ABI shared library exports every symbol include/rsigpu.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from rsicnv_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsigpu.h")).read()
    return sorted(set(re.findall(r"\b(rsigpu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_agree():
    assert set(declared_symbols()) == set(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "rsicnv_b200", "librsigpu.so")
    if not os.path.exists(path):
        pytest.skip("librsigpu.so not built here (make lib)")
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_no_device_is_an_error_not_a_fallback():
    """without a CUDA device the product refuses to run (RSIGPU_E_NODEVICE = 5)"""
    path = os.path.join(ROOT, "rsicnv_b200", "librsigpu.so")
    if not os.path.exists(path):
        pytest.skip("librsigpu.so not built here (make lib)")
    lib = api.load_library(path)
    if lib.rsigpu_num_devices() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.RsiGpuError) as e:
        api.Context(lib=path)
    assert e.value.code == 5


def test_struct_layouts():
    assert ctypes.sizeof(api.Cnv) == 128
    assert ctypes.sizeof(api.Params) == 64
    assert ctypes.sizeof(api.ChrStats) == 72
