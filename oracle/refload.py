"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (``/root/reference/seaduck``) in the
build container so that golden vectors can be generated from it.

The reference needs ``xarray`` and ``dask`` which are not installed here.  We
register tiny stand-in modules (``seaduck_b200.minixr`` posing as ``xarray``; an empty ``dask.array``
that is only touched for dask-backed data, reference ``seaduck/smart_read.py:73``) and then import
the reference package from where it lies.  Nothing is copied.  ``/root/reference`` does not exist on
the GPU box, so only ``oracle/make_golden.py`` and container-only tests call this.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SEADUCK_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "seaduck"))


def load_reference(numba=True):
    """Return the reference ``seaduck`` package (imported once per process)."""
    if "seaduck" in sys.modules:
        return sys.modules["seaduck"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from seaduck_b200 import minixr

    if "xarray" not in sys.modules:
        xr = types.ModuleType("xarray")
        for name in ("DataArray", "Dataset", "zeros_like", "concat", "merge"):
            setattr(xr, name, getattr(minixr, name))
        xr.__stand_in__ = True
        sys.modules["xarray"] = xr
    if "dask" not in sys.modules:
        dask = types.ModuleType("dask")
        dask_array = types.ModuleType("dask.array")
        dask.array = dask_array
        sys.modules["dask"] = dask
        sys.modules["dask.array"] = dask_array
    if not numba:
        sys.modules["numba"] = None  # makes `from numba import njit` raise ImportError
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("seaduck")
