"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""

import ctypes
import os
import re

import pytest

from seaduck_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "seaduck_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/seaduck_b200.h but not exported"


def test_abi_version_and_error_path_without_gpu():
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "seaduck_b200.h")).read()
    assert L.sdk_abi_version() == int(re.search(r"#define SDK_ABI_VERSION (\d+)", header).group(1))
    out = ctypes.c_void_p()
    rc = L.sdk_grid_create(None, ctypes.byref(out))
    assert rc == -1  # SDK_EINVAL, no CUDA call was needed to find that out
    assert b"null" in L.sdk_last_error()
    assert L.sdk_particles_bytes(1000) > 1000 * 200


def test_struct_sizes_match_header():
    """The ctypes mirrors have exactly the size the C compiler gave the structures."""
    from seaduck_b200 import interp

    L = _lib.lib()
    L.sdk_struct_size.restype = ctypes.c_int64
    mirrors = [
        _lib.GridDesc, _lib.Velocity, _lib.Particles, _lib.AdvectParams, _lib.Log, _lib.Located, _lib.TrackIO,
        interp.KernelDesc, interp.Field, interp.Points, _lib.BudgetFields, _lib.WallFields, None, _lib.Exchange,
        interp.InterpJob, interp.InterpIO,
    ]
    for which, cls in enumerate(mirrors):
        if cls is not None:
            assert ctypes.sizeof(cls) == L.sdk_struct_size(which), cls.__name__
    import numpy as np

    assert np.dtype(_lib.SNAPSHOT_RECORD).itemsize == L.sdk_struct_size(12) == 32
    assert np.dtype(_lib.LOG_RECORD).itemsize == 128
    assert L.sdk_struct_size(99) == -1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libseaduck_b200.so")
    with pytest.raises(_lib.SdkError):
        _lib.lib()


def test_every_entry_point_is_typed():
    """ctypes passes untyped integer arguments as 32-bit ints: a 64-bit device pointer would be cut in half.  Every
    function the header declares must therefore have its argument types set by ``_lib.lib()`` itself."""
    L = _lib.lib()
    untyped = [n for n in declared_functions() if getattr(L, n).argtypes is None and n not in ("sdk_abi_version", "sdk_last_error", "sdk_sm_count")]
    assert not untyped, untyped
