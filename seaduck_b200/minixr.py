"""A very small labelled-array container (``Dataset`` / ``DataArray``).

seaduck wraps an ``xarray.Dataset`` (reference ``seaduck/ocedata.py:164``).  xarray is an
optional dependency here: :class:`seaduck_b200.OceData` accepts a real ``xarray.Dataset`` when
xarray is importable, and otherwise this stand-in, which implements only the part of the xarray
surface that the hot path touches (named dims, ``transpose``, ``isel``, ``concat``/``merge`` along
``Zl``, label based ``.loc`` assignment, named-dim broadcasting for ``Vol = drF * rA * hFacC``).

It is host-side plumbing: nothing in here is performance relevant.
"""

from __future__ import annotations

import numpy as np

__all__ = ["DataArray", "Dataset", "zeros_like", "concat", "merge"]


def _as_tuple(dims):
    if isinstance(dims, str):
        return (dims,)
    return tuple(dims)


class _Loc:
    """Label based indexer: ``da.loc[{"Zl": 0}] = 0`` (reference ``lagrangian.py:155``)."""

    def __init__(self, da):
        self._da = da

    def _index(self, key):
        index = [slice(None)] * self._da.ndim
        for dim, label in key.items():
            axis = self._da.dims.index(dim)
            coord = self._da.coords.get(dim)
            if coord is None:
                pos = int(label)
            else:
                coord = np.asarray(coord)
                hit = np.where(coord == label)[0]
                if len(hit) == 0:
                    raise KeyError(f"{label} not found in coordinate {dim}")
                pos = int(hit[0])
            index[axis] = pos
        return tuple(index)

    def __getitem__(self, key):
        return self._da.data[self._index(key)]

    def __setitem__(self, key, value):
        self._da.data[self._index(key)] = value


class DataArray:
    """N-d array with named dimensions."""

    __array_priority__ = 100

    def __init__(self, data, dims=None, coords=None, name=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            coords = data.coords if coords is None else coords
            data = data.data
        self.data = data if hasattr(data, "shape") and hasattr(data, "dtype") else np.asarray(data)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(self.data.ndim))
        self.dims = _as_tuple(dims)
        if len(self.dims) != self.data.ndim:
            raise ValueError(f"dims {self.dims} do not match data of shape {self.data.shape}")
        self.coords = dict(coords) if coords else {}
        self.name = name

    # ---- numpy-ish surface -------------------------------------------------
    @property
    def values(self):
        return np.asarray(self.data)

    @values.setter
    def values(self, new):
        new = np.asarray(new)
        if new.shape != self.data.shape:
            raise ValueError("replacement data must match the existing shape")
        self.data = new

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def nbytes(self):
        return self.data.nbytes

    @property
    def size(self):
        return self.data.size

    @property
    def sizes(self):
        return dict(zip(self.dims, self.data.shape))

    @property
    def loc(self):
        return _Loc(self)

    def __array__(self, dtype=None, copy=None):
        arr = np.asarray(self.data)
        if dtype is not None:
            arr = arr.astype(dtype, copy=False)
        return arr

    def __len__(self):
        if self.data.ndim == 0:
            raise TypeError("len() of unsized object")
        return self.data.shape[0]

    def __float__(self):
        return float(self.data)

    def __int__(self):
        return int(self.data)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        # expand Ellipsis
        if any(k is Ellipsis for k in key):
            pos = [i for i, k in enumerate(key) if k is Ellipsis][0]
            fill = self.ndim - (len(key) - 1)
            key = key[:pos] + (slice(None),) * fill + key[pos + 1 :]
        key = key + (slice(None),) * (self.ndim - len(key))
        new_dims = []
        new_coords = {}
        for dim, k in zip(self.dims, key):
            if isinstance(k, (int, np.integer)):
                continue
            new_dims.append(dim)
            if dim in self.coords:
                new_coords[dim] = np.asarray(self.coords[dim])[k]
        return DataArray(self.data[key], dims=new_dims, coords=new_coords, name=self.name)

    def __setitem__(self, key, value):
        self.data[key] = np.asarray(value)

    def __repr__(self):
        return f"<minixr.DataArray {self.name!r} {dict(zip(self.dims, self.shape))} {self.dtype}>"

    # ---- reductions / elementwise -----------------------------------------
    def min(self):
        return np.asarray(self.data).min()

    def max(self):
        return np.asarray(self.data).max()

    def fillna(self, value):
        arr = np.array(self.data, copy=True)
        if np.issubdtype(arr.dtype, np.floating):
            arr[np.isnan(arr)] = value
        return DataArray(arr, self.dims, self.coords, self.name)

    def astype(self, dtype):
        return DataArray(np.asarray(self.data).astype(dtype), self.dims, self.coords, self.name)

    def copy(self, deep=True):
        data = np.array(self.data, copy=True) if deep else self.data
        return DataArray(data, self.dims, dict(self.coords), self.name)

    def transpose(self, *dims, missing_dims="raise"):
        order = _resolve_order(self.dims, dims, missing_dims)
        if order == self.dims:
            return self
        axes = [self.dims.index(d) for d in order]
        return DataArray(np.transpose(self.data, axes), order, self.coords, self.name)

    def isel(self, **indexers):
        key = [slice(None)] * self.ndim
        for dim, k in indexers.items():
            if dim in self.dims:
                key[self.dims.index(dim)] = k
        return self[tuple(key)]

    # named-dim broadcasting ---------------------------------------------------
    def _binary(self, other, op):
        if isinstance(other, DataArray):
            # xarray semantics: result dims = dims of self, then the new dims of other
            dims = list(self.dims) + [d for d in other.dims if d not in self.dims]
            a = _expand(self, dims)
            b = _expand(other, dims)
            coords = dict(other.coords)
            coords.update(self.coords)
            return DataArray(op(a, b), dims, coords, self.name)
        return DataArray(op(np.asarray(self.data), other), self.dims, self.coords, self.name)

    def __add__(self, other):
        return self._binary(other, np.add)

    def __radd__(self, other):
        return self._binary(other, lambda a, b: np.add(b, a))

    def __sub__(self, other):
        return self._binary(other, np.subtract)

    def __rsub__(self, other):
        return self._binary(other, lambda a, b: np.subtract(b, a))

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    def __rmul__(self, other):
        return self._binary(other, lambda a, b: np.multiply(b, a))

    def __truediv__(self, other):
        return self._binary(other, np.true_divide)

    def __neg__(self):
        return DataArray(-np.asarray(self.data), self.dims, self.coords, self.name)


def _expand(da, dims):
    """View of ``da.data`` broadcastable against an array laid out as ``dims``."""
    present = [d for d in dims if d in da.dims]
    arr = np.transpose(np.asarray(da.data), [da.dims.index(d) for d in present])
    shape = [arr.shape[present.index(d)] if d in present else 1 for d in dims]
    return arr.reshape(shape)


def _resolve_order(have, want, missing_dims):
    want = list(want)
    if Ellipsis in want:
        pos = want.index(Ellipsis)
        named = [d for d in want if d is not Ellipsis]
        rest = [d for d in have if d not in named]
        want = want[:pos] + rest + want[pos + 1 :]
    if missing_dims == "ignore":
        want = [d for d in want if d in have]
    missing = [d for d in want if d not in have]
    if missing:
        raise ValueError(f"dimensions {missing} do not exist")
    tail = [d for d in have if d not in want]
    return tuple(want + tail)


class _VarView(dict):
    """Mapping returned by ``Dataset.variables`` / ``data_vars`` / ``coords``."""


class Dataset:
    """Dictionary of :class:`DataArray` sharing dimension names."""

    def __init__(self, data_vars=None, coords=None):
        object.__setattr__(self, "_vars", {})
        object.__setattr__(self, "_coord_names", [])
        for name, value in (coords or {}).items():
            self._vars[name] = self._coerce(name, value)
            self._coord_names.append(name)
        for name, value in (data_vars or {}).items():
            self._vars[name] = self._coerce(name, value)
        self._attach_coords()

    @staticmethod
    def _coerce(name, value):
        if isinstance(value, DataArray):
            out = DataArray(value.data, value.dims, value.coords, name)
        elif isinstance(value, tuple):
            dims, data = value[0], value[1]
            out = DataArray(np.asarray(data) if not hasattr(data, "shape") else data, dims, name=name)
        else:
            arr = np.asarray(value)
            if arr.ndim == 1:
                out = DataArray(arr, (name,), name=name)
            elif arr.ndim == 0:
                out = DataArray(arr, (), name=name)
            else:
                raise ValueError(f"cannot infer dims for {name}")
        return out

    def _attach_coords(self):
        """Give every variable access to its 1-D index coordinates (for ``.loc``)."""
        index = {
            name: np.asarray(var.data)
            for name, var in self._vars.items()
            if var.dims == (name,)
        }
        for var in self._vars.values():
            var.coords = {d: index[d] for d in var.dims if d in index}

    # ---- mapping surface -----------------------------------------------------
    @property
    def variables(self):
        return _VarView(self._vars)

    @property
    def data_vars(self):
        return _VarView({k: v for k, v in self._vars.items() if k not in self._coord_names})

    @property
    def coords(self):
        return _VarView({k: v for k, v in self._vars.items() if k in self._coord_names})

    @property
    def dims(self):
        out = {}
        for var in self._vars.values():
            for d, n in zip(var.dims, var.shape):
                out[d] = n
        return out

    sizes = dims

    def keys(self):
        return self._vars.keys()

    def __iter__(self):
        return iter(self.data_vars)

    def __contains__(self, key):
        return key in self._vars

    def __len__(self):
        return len(self.data_vars)

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            names = list(key)
            sub = Dataset()
            dims_used = set()
            for n in names:
                sub._vars[n] = self._vars[n]
                dims_used.update(self._vars[n].dims)
                if n in self._coord_names:
                    sub._coord_names.append(n)
            # keep the index coordinates of the selected variables (like xarray does)
            for c in self._coord_names:
                if c not in sub._vars and set(self._vars[c].dims) <= dims_used:
                    sub._vars[c] = self._vars[c]
                    sub._coord_names.append(c)
            sub._attach_coords()
            return sub
        try:
            return self._vars[key]
        except KeyError:
            # a dimension without coordinate variable: xarray hands out its integer index
            dims = self.dims
            if key in dims:
                return DataArray(np.arange(dims[key]), dims=(key,), name=key)
            raise KeyError(key) from None

    def __setitem__(self, key, value):
        self._vars[key] = self._coerce(key, value)
        self._attach_coords()

    def __getattr__(self, attr):
        try:
            return object.__getattribute__(self, "_vars")[attr]
        except KeyError:
            raise AttributeError(attr) from None

    def __repr__(self):
        lines = ["<minixr.Dataset>"]
        for k, v in self._vars.items():
            tag = "*" if k in self._coord_names else " "
            lines.append(f" {tag} {k}: {dict(zip(v.dims, v.shape))} {v.dtype}")
        return "\n".join(lines)

    # ---- structural ops ------------------------------------------------------
    def _like(self, new_vars):
        out = Dataset()
        out._vars.update(new_vars)
        out._coord_names.extend(c for c in self._coord_names if c in new_vars)
        out._attach_coords()
        return out

    def transpose(self, *dims, missing_dims="raise"):
        return self._like(
            {k: v.transpose(*dims, missing_dims="ignore") for k, v in self._vars.items()}
        )

    def isel(self, **indexers):
        return self._like({k: v.isel(**indexers) for k, v in self._vars.items()})

    def copy(self, deep=True):
        return self._like({k: v.copy(deep=deep) for k, v in self._vars.items()})


def zeros_like(obj):
    if isinstance(obj, DataArray):
        return DataArray(np.zeros_like(np.asarray(obj.data)), obj.dims, obj.coords, obj.name)
    new = {}
    for k, v in obj._vars.items():
        if k in obj._coord_names:
            new[k] = v
        else:
            new[k] = DataArray(np.zeros_like(np.asarray(v.data)), v.dims, v.coords, k)
    return obj._like(new)


def concat(objs, dim):
    first = objs[0]
    if isinstance(first, DataArray):
        axis = first.dims.index(dim)
        data = np.concatenate([np.asarray(o.data) for o in objs], axis=axis)
        return DataArray(data, first.dims, name=first.name)
    new = {}
    for k, v in first._vars.items():
        if dim in v.dims:
            axis = v.dims.index(dim)
            new[k] = DataArray(
                np.concatenate([np.asarray(o._vars[k].data) for o in objs], axis=axis), v.dims, name=k
            )
        else:
            new[k] = v
    return first._like(new)


def merge(objs):
    out = Dataset()
    for o in objs:
        for k, v in o._vars.items():
            if k not in out._vars:
                out._vars[k] = v
                if k in o._coord_names:
                    out._coord_names.append(k)
    out._attach_coords()
    return out
