"""CPU: libpicca_b200.so loads without a GPU and exports every symbol include/picca_b200.h
declares; struct layouts agree between the header and the ctypes mirror.  No compute calls."""
import ctypes
import os
import re

from picca_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "picca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_all_symbols():
    handle = _lib.lib()
    for name in declared_symbols():
        assert hasattr(handle, name), name
    assert handle.pb2_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match():
    handle = _lib.lib()
    assert handle.pb2_sizeof_params() == ctypes.sizeof(_lib.Params)
    assert handle.pb2_sizeof_catalog() == ctypes.sizeof(_lib.Catalog)
    assert handle.pb2_sizeof_pairs() == ctypes.sizeof(_lib.Pairs)


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpicca_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
