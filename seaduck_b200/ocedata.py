"""Ocean dataset wrapper + device staging of the grid.

Same role and public surface as the reference's ``seaduck.ocedata`` (``OceData`` ``ocedata.py:147``,
``RelCoord`` families ``:34-143``): variable aliasing, readiness checks, grid extraction to float32
arrays, rel-coordinate lookups.  What differs is where things live: the grid arrays are uploaded
once to the GPU (``sdk_grid_create``), the KD-tree is replaced by a device bucket grid
(``sdk_grid_build_locator``) and every ``_find_rel_*`` call runs the ``sdk_locate`` kernel.
"""

from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib, minixr
from .topology import Topology

NO_ALIAS = {
    "dXC": "dxC",
    "dYC": "dyC",
    "dZ": "drC",
    "dXG": "dxG",
    "dYG": "dyG",
    "dZl": "drF",
}


def _general_len(thing):
    try:
        return len(thing)
    except TypeError:
        return 1


class Deferred:
    """A rel-coordinate that still lives on the GPU: fetched (and converted to the reference's numpy
    dtype) the first time somebody looks at it.  ``from_latlon`` of 10^7 points would otherwise copy
    ~1.5 GB of rel-coords to the host that a following ``interpolate`` never reads."""

    __slots__ = ("tensor", "dtype")

    def __init__(self, tensor, dtype=None):
        self.tensor, self.dtype = tensor, dtype

    def fetch(self):
        a = self.tensor.cpu().numpy()
        return a.astype(self.dtype) if self.dtype is not None else a


class DeviceVar:
    """A data variable that lives on the GPU only (``OceData.stage_device_field``): what ``od[name]`` returns for it.
    Carries the dimension names and the shape, like the labelled arrays of the dataset do."""

    def __init__(self, tensor, dims):
        self.tensor, self.dims = tensor, tuple(dims)
        if len(self.dims) != tensor.ndim:
            raise ValueError(f"dims {self.dims} do not match a tensor of shape {tuple(tensor.shape)}")

    shape = property(lambda self: tuple(self.tensor.shape))
    ndim = property(lambda self: self.tensor.ndim)
    nbytes = property(lambda self: self.tensor.numel() * self.tensor.element_size())
    dtype = property(lambda self: np.dtype(str(self.tensor.dtype).replace("torch.", "")))

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.cpu().numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


class RelCoord(dict):
    """Relative coordinates: dictionary with attribute access (reference ``ocedata.py:34``).
    Values may be :class:`Deferred`; every read access materialises them."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, Deferred):
            v = v.fetch()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def has(self, key):
        """True if ``key`` holds something (without fetching it)."""
        return dict.get(self, key) is not None

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getattr__(self, attr):
        try:
            return self[attr]
        except KeyError as exc:
            raise AttributeError(f"'RelCoord' object has no attribute '{attr}'") from exc

    def __setattr__(self, attr, value):
        self[attr] = value

    def subset(self, which):
        new = RelCoord()
        for var in self.keys():
            new[var] = None if self[var] is None else self[var][which]
        return new

    def update_from_subset(self, sub, which):
        for var in sub.keys():
            if var not in self.keys():
                raise KeyError(f"Key {var} is in subset, but not in the one to update.")
            if self[var] is not None:
                self[var][which] = sub[var]

    @classmethod
    def create_class(cls, class_name, fields):
        """A named family of rel-coordinates (reference ``ocedata.py:95``): a ``RelCoord`` whose positional arguments
        fill ``fields`` in order, with ``_make`` / ``_fields`` like a namedtuple."""
        fields = tuple(fields)

        def fill(self, *values):
            if len(values) > len(fields):
                raise TypeError(f"{class_name} takes {len(fields)} positional arguments but {len(values)} were given")
            self.update(zip(fields, values))

        family = type(class_name, (cls,), {
            "__slots__": (),
            "_fields": fields,
            "__init__": fill,
            "_make": classmethod(lambda kind, values: kind(*values)),
            "__repr__": lambda self: class_name + "(" + ", ".join(f"{k}={self[k]!r}" for k in fields if k in self) + ")",
        })
        return family


HRel = RelCoord.create_class("HRel", ["face", "iy", "ix", "rx", "ry", "cs", "sn", "dx", "dy", "bx", "by"])
VRel = RelCoord.create_class("VRel", ["iz", "rz", "dz", "bz"])
VLinRel = RelCoord.create_class("VLRel", ["iz_lin", "rz_lin", "dz_lin", "bz_lin"])
VlRel = RelCoord.create_class("VlRel", ["izl", "rzl", "dzl", "bzl"])
VlLinRel = RelCoord.create_class("VlLinRel", ["izl_lin", "rzl_lin", "dzl_lin", "bzl_lin"])
TRel = RelCoord.create_class("TRel", ["it", "rt", "dt", "bt"])
TLinRel = RelCoord.create_class("TLinRel", ["it_lin", "rt_lin", "dt_lin", "bt_lin"])

_TYP = {"box": 0, "x_periodic": 1, "LLC": 2}
_HMODE = {"oceanparcel": 0, "local_cartesian": 1, "rectilinear": 2}


def _xr_module(data):
    if isinstance(data, minixr.Dataset):
        return minixr
    import xarray as xr  # noqa: PLC0415

    return xr


class OceData:
    """Ocean dataset: aliasing, topology, grid staging on the GPU, 4-D coordinate translation.

    Parameters follow the reference (``ocedata.py:164``).  ``device`` selects the CUDA device
    (default: current).
    """

    # canonical order of the dimensions every array of the dataset is brought into (reference ocedata.py:165-179)
    _DIM_ORDER = ("time", "time_midp", "time_outer", "Z", "Zl", "Zp1", "face", "Y", "Yp1", "X", "Xp1")
    # what each horizontal scheme needs to be in the dataset, in order of preference (reference ocedata.py:231)
    _SCHEMES = (
        ("oceanparcel", ("XC", "YC", "XG", "YG")),
        ("local_cartesian", ("XC", "YC", "dxG", "dyG", "CS", "SN")),
        ("rectilinear", ("lon", "lat")),
    )

    def __init__(self, data, alias=None, memory_limit=1e7, device=None, device_memory_limit=None):
        self._xr = _xr_module(data)
        self._ds = data.transpose(*self._DIM_ORDER, ..., missing_dims="ignore")
        self.tp = Topology(self._ds)
        if alias == "auto":
            raise NotImplementedError("auto alias not yet implemented")
        if alias is None:
            self.alias = NO_ALIAS
        elif isinstance(alias, dict):
            self.alias = alias
        self.too_large = "XC" in self._ds.variables and self._ds["XC"].nbytes > memory_limit
        self.device = device
        # Out-of-core fields (the reference's `too_large` route, ocedata.py:188-191 + smart_read.py:73-107, which
        # gathers only the needed values from a dask array): a data variable (or time slab of one) larger than
        # `device_memory_limit` bytes -- default: 70 % of the free device memory at the time it is first needed -- is not
        # uploaded; it stays in page-locked host memory and the kernels gather from it in place over the bus.
        self.device_memory_limit = device_memory_limit
        self._host_resident = set()  # names forced to stay on the host (keep_on_host)
        self._dev = {}  # name -> torch tensor on the GPU
        self._field_cache = {}
        self._interp_cache = {}  # rim tables and masks of the interpolation kernel
        self._grid_handle = None
        self._device_vars = {}  # name -> DeviceVar (fields staged on the GPU without a host copy)
        self.readiness = self.check_readiness()
        self._grid2array()
        if self.readiness["Zl"]:
            self._append_zero_bottom_level()

    def _append_zero_bottom_level(self):
        """Every variable on ``Zl`` gets one more level of zeros below the deepest one, at ``Zl[-1]`` of the extended
        axis (reference ``ocedata.py:197-210``): the vertical velocity at the sea floor, which the particle kernels read
        as the lower wall of the deepest cell."""
        xr, ds = self._xr, self._ds
        on_zl = [name for name in ds.data_vars if "Zl" in ds[name].dims]
        if not on_zl:
            return
        elsewhere = [name for name in ds.data_vars if name not in on_zl]
        floor = xr.zeros_like(ds[on_zl].isel(Zl=slice(1)))
        floor["Zl"] = [self.Zl[-1]]
        self._ds = xr.merge([xr.concat([ds[on_zl], floor], dim="Zl"), ds[elsewhere]])

    # ---- mapping surface (reference ocedata.py:212-229) ---------------------------------------
    def _dataset_name(self, key):
        return self.alias.get(key, key)

    def __setitem__(self, key, item):
        # labelled arrays go into the dataset (under the aliased name), anything else becomes an attribute
        if isinstance(item, self._xr.DataArray):
            self._ds[self._dataset_name(key)] = item
        else:
            object.__setattr__(self, key, item)

    def __getitem__(self, key):
        # attributes (the float32 grid arrays, Vol ...) shadow staged device fields, which shadow the dataset
        own = vars(self)
        if key in own:
            return own[key]
        staged = own.get("_device_vars", {})
        if key in staged:
            return staged[key]
        return self._ds[self._dataset_name(key)]

    def check_readiness(self):
        """Which interpolation schemes the dataset supports (reference ``ocedata.py:231``): ``h`` names the horizontal
        scheme, ``Z`` / ``Zl`` / ``time`` say whether that axis exists with more than one entry."""
        present = set(self._ds.variables)
        scheme = next((name for name, needs in self._SCHEMES if present.issuperset(needs)), None)
        if scheme is None:
            raise ValueError(
                """
                Need a full set of the following to create the object:
                {XC,YC,XG,YG} or {XC,YC,dxG,dyG,CS,SN} for curvilinear grid;
                or {lon, lat} for rectilinear grid.

                See add_missing variable or set_alias.
                """
            )
        if scheme == "rectilinear":
            metres_per_degree = 6371e3 * np.pi / 180
            self.dlon = np.gradient(self["lon"]) * metres_per_degree
            self.dlat = np.gradient(self["lat"]) * metres_per_degree
        readiness = {"h": scheme}
        for axis in ("time", "Z", "Zl"):
            readiness[axis] = axis in present and _general_len(self[axis]) > 1
        return readiness

    def _add_missing_vol(self, as_numpy=False):
        """``Vol = drF * rA * hFacC`` (reference ``ocedata.py:303``)."""
        if self.readiness["Zl"]:
            vol = self._ds["drF"] * self._ds["rA"]
            if "HFacC" in self._ds.variables:
                vol = vol * self._ds["HFacC"]
            elif "hFacC" in self._ds.variables:
                vol = vol * self._ds["hFacC"]
            vol = vol.transpose("Z", "face", "Y", "X", missing_dims="ignore")
        else:
            vol = self._ds["rA"]
        vol = vol.fillna(0)
        if as_numpy:
            self["Vol"] = np.array(vol)
        else:
            self["Vol"] = vol

    # ---- grid extraction ----------------------------------------------------------------------
    def _grid2array(self):
        if self.readiness["h"] == "oceanparcel":
            # (np.asarray: no second host copy of a grid that already is float32 -- LLC4320 is 1 GB per array)
            for var in ["XC", "YC", "XG", "YG"]:
                self[var] = np.asarray(self[var], dtype="float32")
            for var in ["rA", "CS", "SN"]:
                try:
                    self[var] = np.asarray(self[var], dtype="float32")
                except KeyError:
                    logging.info("no %s in dataset, skip", var)
                    self[var] = None
            try:
                self.dX = np.asarray(self["dXG"], dtype="float32")
                self.dY = np.asarray(self["dYG"], dtype="float32")
            except KeyError:
                self.dX = None
                self.dY = None
        elif self.readiness["h"] == "local_cartesian":
            for var in ["XC", "YC", "CS", "SN"]:
                self[var] = np.array(self[var], dtype="float32")
            self.dX = np.array(self["dXG"], dtype="float32")
            self.dY = np.array(self["dYG"], dtype="float32")
            for var in ["XG", "YG", "dXC", "dYC", "rA"]:
                try:
                    self[var] = np.array(self[var], dtype="float32")
                except KeyError:
                    self[var] = None
        elif self.readiness["h"] == "rectilinear":
            self.lon = np.array(self["lon"], dtype="float32")
            self.lat = np.array(self["lat"], dtype="float32")
        if self.readiness["Z"]:
            self.Z = np.array(self["Z"], dtype="float32")
            try:
                self.dZ = np.array(self["dZ"], dtype="float32")
            except KeyError:
                self.dZ = np.diff(self.Z)
                self.dZ = np.append(self.dZ, self.dZ[-1]).astype("float32")
        if self.readiness["Zl"]:
            if "Zp1" in self._ds.variables:
                self.Zl = np.array(self["Zp1"], dtype="float32")
            else:
                self.Zl = np.zeros(len(self["Zl"]) + 1)
                self.Zl[:-1] = np.array(self._ds["Zl"])
                self.Zl[-1] = 2 * self.Zl[-2] - self.Zl[-3]
                self.Zl = self.Zl.astype("float32")
            try:
                self.dZl = np.array(self["dZl"], dtype="float32")
            except KeyError:
                self.dZl = np.diff(self.Zl).astype("float32")
        if self.readiness["time"]:
            self.t_base = 0
            self.ts = np.array(self["time"])
            if self["time"].dtype != "float":
                self.ts = (self.ts).astype(float) / 1e9
            try:
                self.time_midp = np.array(self["time_midp"])
                self.time_midp = (self.time_midp).astype(float) / 1e9
            except KeyError:
                self.time_midp = (self.ts[1:] + self.ts[:-1]) / 2

    # ---- device side ----------------------------------------------------------------------------
    def _torch_device(self):
        import torch  # noqa: PLC0415

        if not torch.cuda.is_available():
            raise _lib.SdkError("seaduck_b200 needs a CUDA device (B200); there is no CPU fallback")
        if self.device is None:
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device(self.device)

    def to_device(self, name, array, dtype=None):
        """Upload (once) a host array; returns the cached torch tensor."""
        import torch  # noqa: PLC0415

        if name not in self._dev:
            arr = np.ascontiguousarray(array, dtype=dtype)
            self._dev[name] = torch.from_numpy(arr).to(self._torch_device())
        return self._dev[name]

    def grid_handle(self):
        """Create (once) the device grid: uploads the float32 grid and the topology tables."""
        if self._grid_handle is not None:
            return self._grid_handle
        L = _lib.lib()
        tp = self.tp
        d = _lib.GridDesc()
        d.typ = _TYP[tp.typ]
        d.has_face = int(tp.has_face)
        d.nface = tp.h_shape[0] if tp.has_face else 1
        d.ny, d.nx = tp.h_shape[-2:]
        d.gny, d.gnx = tp.g_shape[-2:]
        d.hmode = _HMODE[self.readiness["h"]]
        keep = []
        if tp.typ == "LLC":
            nf = np.ascontiguousarray(tp.nbr_face, np.int32)
            ne = np.ascontiguousarray(tp.nbr_edge, np.int32)
            keep += [nf, ne]
            d.nbr_face = nf.ctypes.data
            d.nbr_edge = ne.ctypes.data
        for name in ("XC", "YC", "XG", "YG", "dX", "dY", "CS", "SN"):
            arr = getattr(self, name, None)
            if arr is not None:
                setattr(d, name, self.to_device(name, arr, np.float32).data_ptr())
        hmode = self.readiness["h"]
        if hmode == "rectilinear":
            # the 1-D axes, their metric spacing (ocedata.py:255-259) and numpy's float32 cosine of the float32
            # latitudes, which Particle._cross_cell_wall_read / _rel evaluate (lagrangian.py:671, :693-703)
            d.lon = self.to_device("lon", self.lon, np.float32).data_ptr()
            d.lat = self.to_device("lat", self.lat, np.float32).data_ptr()
            d.dlon_is_f32 = int(np.asarray(self.dlon).dtype == np.float32)
            d.dlon = self.to_device("dlon", self.dlon, np.float64).data_ptr()
            d.dlat = self.to_device("dlat", self.dlat, np.float64).data_ptr()
            d.coslat32 = self.to_device("coslat32", self._coslat32(self.lat), np.float32).data_ptr()
        elif hmode == "local_cartesian":
            d.coslat32 = self.to_device("coslat32", self._coslat32(self.YC), np.float32).data_ptr()
        if self.readiness["Zl"]:
            d.n_zl = len(self.Zl)
            d.n_z = len(self.dZl)
            d.Zl = self.to_device("Zl", self.Zl, np.float32).data_ptr()
            d.dZl = self.to_device("dZl", self.dZl, np.float32).data_ptr()
        if self.readiness["Z"]:
            d.n_zc = len(self.Z)
            d.Z = self.to_device("Z", self.Z, np.float32).data_ptr()
            d.dZ = self.to_device("dZ", self.dZ, np.float32).data_ptr()
        if hmode == "oceanparcel" and (d.gny <= d.ny or d.gnx <= d.nx):
            d.corner_tab = self.to_device("corner_tab", self._corner_table(), np.int32).data_ptr()
        handle = C.c_void_p()
        _lib.check(L.sdk_grid_create(C.byref(d), C.byref(handle)), "sdk_grid_create")
        self._grid_handle = handle
        self._keep = keep
        self._build_locator()
        return handle

    @staticmethod
    def _coslat32(lat32):
        """``np.cos(by * np.pi / 180)`` exactly as the reference's plain-numpy code gets it for a float32
        latitude array (float32 product, quotient and cosine)."""
        lat32 = np.asarray(lat32)
        assert lat32.dtype == np.float32
        out = np.cos(lat32 * np.pi / 180)
        assert out.dtype == np.float32
        return out

    def _corner_table(self):
        """Flat XG indices of the LR, UR, UL corners of the cells on the top row / right column."""
        tp = self.tp
        ny, nx = tp.h_shape[-2:]
        gny, gnx = tp.g_shape[-2:]
        nface = tp.h_shape[0] if tp.has_face else 1
        f = np.repeat(np.arange(nface), nx + ny)
        slot = np.tile(np.arange(nx + ny), nface)
        iy = np.where(slot < nx, ny - 1, slot - nx)
        ix = np.where(slot < nx, slot, nx - 1)
        ind = (f, iy, ix) if tp.has_face else (iy, ix)
        n = len(f)
        table = np.full((n, 3), -1, np.int64)
        ind1, st1 = tp.ind_tend_vec(ind, np.full(n, 3), cuvwg="G", swallow=True, return_status=True)
        ind2, st2 = tp.ind_tend_vec(tuple(ind1), np.full(n, 0), cuvwg="G", swallow=True, return_status=True)
        ind3, st3 = tp.ind_tend_vec(ind, np.full(n, 0), cuvwg="G", swallow=True, return_status=True)
        for k, (idx, st) in enumerate(((ind1, st1), (ind2, st2 | st1), (ind3, st3))):
            idx = [np.asarray(a) for a in idx]
            yy = np.where(idx[-2] < 0, idx[-2] + gny, idx[-2])
            xx = np.where(idx[-1] < 0, idx[-1] + gnx, idx[-1])
            ff = idx[0] if tp.has_face else 0
            flat = (ff * gny + yy) * gnx + xx
            abnormal = (st == 2) | (yy >= gny) | (xx >= gnx)
            table[:, k] = np.where(abnormal, -1, flat)
        return table.astype(np.int32)

    def _build_locator(self):
        if self.readiness["h"] == "rectilinear":
            return  # two 1-D searches on the device, nothing to build
        L = _lib.lib()
        xc = np.asarray(self.XC, np.float64)
        yc = np.asarray(self.YC, np.float64)
        ncell = xc.size
        lat0, lat1 = float(np.nanmin(yc)), float(np.nanmax(yc))
        lon_min, lon_max = float(np.nanmin(xc)), float(np.nanmax(xc))
        if self.tp.typ in ("LLC", "x_periodic") or lon_max - lon_min > 300:
            lon0, span_lon = -180.0, 360.0
        else:
            pad = 1e-6 + 0.01 * (lon_max - lon_min)
            lon0, span_lon = lon_min - pad, (lon_max - lon_min) + 2 * pad
        pad = 1e-6 + 0.01 * (lat1 - lat0)
        lat0, span_lat = lat0 - pad, (lat1 - lat0) + 2 * pad
        # about one cell per bucket
        aspect = max(span_lon / span_lat, 1e-3)
        nlat = int(np.clip(np.sqrt(ncell / aspect), 4, 8192))
        nlon = int(np.clip(nlat * aspect, 4, 16384))
        _lib.check(
            L.sdk_grid_build_locator(self._grid_handle, lon0, lat0, span_lon, span_lat, nlon, nlat, _lib.current_stream()),
            "sdk_grid_build_locator",
        )

    def __del__(self):
        try:
            if self._grid_handle is not None:
                _lib.lib().sdk_grid_destroy(self._grid_handle)
        except Exception:  # noqa: BLE001
            pass

    @staticmethod
    def _slab_key(name, time_slice):
        return (name, None if time_slice is None else (time_slice.start, time_slice.stop))

    def _host_slab(self, name, time_slice):
        da = self[name]
        arr = np.asarray(da[time_slice] if time_slice is not None else da)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        return np.ascontiguousarray(arr)

    def _remember_slab(self, key, value):
        # keep at most two slabs per variable (current + previous / prefetched)
        same = [k for k in self._field_cache if k[0] == key[0]]
        for k in same[:-1]:
            del self._field_cache[k]
        self._field_cache[key] = value

    def stage_device_field(self, name, tensor, dims):
        """Register a data variable that already is on the GPU (e.g. written there by a model or generated there):
        no host copy is made or needed; ``dims`` as in the dataset (``("time", "face", "Y", "Xp1")`` ...)."""
        if tensor.device != self._torch_device():
            raise ValueError(f"{name}: tensor on {tensor.device}, the dataset lives on {self._torch_device()}")
        self._device_vars[name] = DeviceVar(tensor.contiguous(), dims)
        self._field_cache = {k: v for k, v in self._field_cache.items() if k[0] != name}

    def device_field(self, name, time_slice=None):
        """Device copy of a data variable (optionally a slab of snapshots), cached.

        Replaces ``Particle.update_uvw_array`` / ``smart_read`` gathers from host memory
        (reference ``lagrangian.py:175``, ``smart_read.py:22``): fields stay resident in HBM.
        A slab that :meth:`prefetch_field` is still copying is waited for on the current stream.
        """
        import torch  # noqa: PLC0415

        staged = self._device_vars.get(name)
        if staged is not None:
            return staged.tensor if time_slice is None else staged.tensor[time_slice]
        key = self._slab_key(name, time_slice)
        hit = self._field_cache.get(key)
        if hit is not None:
            if isinstance(hit, tuple):  # (tensor, event, pinned staging buffer) from prefetch_field
                t, ev, _ = hit
                cur = torch.cuda.current_stream(t.device)
                cur.wait_event(ev)
                t.record_stream(cur)  # allocated under the copy stream, used on this one from now on
                self._field_cache[key] = t
                return t
            return hit
        host = self._host_slab(name, time_slice)
        if self._stays_on_host(name, host.nbytes):
            # page-locked and mapped: under unified addressing the kernels read it through the same pointer
            t = torch.from_numpy(host).pin_memory()
        else:
            t = torch.from_numpy(host).to(self._torch_device())
        self._remember_slab(key, t)
        return t

    def keep_on_host(self, *names):
        """Never upload these data variables: the kernels read them in place from page-locked host memory (for
        fields that do not fit the device next to everything else; every gather then crosses the bus)."""
        self._host_resident.update(names)
        self._field_cache = {k: v for k, v in self._field_cache.items() if k[0] not in names}
        return self

    def _stays_on_host(self, name, nbytes):
        import torch  # noqa: PLC0415

        if name in self._host_resident:
            return True
        limit = self.device_memory_limit
        if limit is None:
            free, _ = torch.cuda.mem_get_info(self._torch_device())
            limit = 0.7 * free
        return nbytes > limit

    def prefetch_field(self, name, time_slice):
        """Start copying a slab to the device on a side stream and return at once (streaming ingest for
        time-varying runs: the copy of the next velocity snapshot overlaps the current leg).  The
        staging buffer is pinned; :meth:`device_field` picks the slab up later."""
        import torch  # noqa: PLC0415

        key = self._slab_key(name, time_slice)
        if key in self._field_cache or name in self._device_vars or name in self._host_resident:
            return
        dev = self._torch_device()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        host = torch.from_numpy(self._host_slab(name, time_slice)).pin_memory()
        if self._stays_on_host(name, host.numel() * host.element_size()):
            self._remember_slab(key, host)  # too large to upload: read in place
            return
        with torch.cuda.stream(self._copy_stream):
            t = torch.empty(host.shape, dtype=host.dtype, device=dev)
            t.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._remember_slab(key, (t, ev, host))

    # ---- rel-coordinate lookups (reference ocedata.py:422-488), all on the device ---------------
    def locate(self, x=None, y=None, z=None, t=None):
        """Run ``sdk_locate``; returns a dict of torch tensors on the device."""
        import torch  # noqa: PLC0415

        L = _lib.lib()
        handle = self.grid_handle()
        dev = self._torch_device()
        n = max(_general_len(v) for v in (x, y, z, t) if v is not None)
        # (non_blocking: arrays from interp.pinned() go up by DMA; pageable ones are staged by the driver as before)
        up = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, np.float64)).to(dev, non_blocking=True)  # noqa: E731
        dx, dy, dz, dt = up(x), up(y), up(z), up(t)
        out = {}
        loc = _lib.Located()

        def alloc(name, dtype, mult=1):
            out[name] = torch.empty(n * mult, dtype=dtype, device=dev)
            setattr(loc, name, out[name].data_ptr())

        out["status"] = torch.zeros(n, dtype=torch.int32, device=dev)
        loc.status = out["status"].data_ptr()
        if dx is not None:
            if self.tp.has_face:
                alloc("face", torch.int32)
            for nme in ("iy", "ix"):
                alloc(nme, torch.int32)
            for nme in ("rx", "ry"):
                alloc(nme, torch.float64)
            rectilinear = self.readiness["h"] == "rectilinear"
            for nme in ("cs", "sn", "bx", "by") + (() if rectilinear else ("dx", "dy")):
                alloc(nme, torch.float32)
            alloc("px", torch.float64, 4)
            alloc("py", torch.float64, 4)
            if rectilinear:
                # dx, dy = cos(by) * deg2m * spacing are float64 in the reference (utils.py:602): rows 2 of px / py
                # (no float32 copy is asked for: sdk_located.dx / dy stay NULL)
                out["dx"], out["dy"] = out["px"][2 * n : 3 * n], out["py"][2 * n : 3 * n]
        if dz is not None:
            if self.readiness["Z"]:
                alloc("iz", torch.int32)
                alloc("iz_lin", torch.int32)
                for nme in ("rz", "dz", "bz", "rz_lin", "dz_lin", "bz_lin"):
                    alloc(nme, torch.float64)
            if self.readiness["Zl"]:
                alloc("izl_near", torch.int32)
                alloc("izl", torch.int32)
                for nme in ("rzl_near", "dzl_near", "bzl_near", "rzl", "dzl", "bzl"):
                    alloc(nme, torch.float64)
        ts_dev = None
        n_t = 0
        if dt is not None and self.readiness["time"]:
            ts_dev = self.to_device("ts", self.ts, np.float64)
            n_t = len(self.ts)
            alloc("it", torch.int32)
            alloc("it_lin", torch.int32)
            for nme in ("rt", "dt", "bt", "rt_lin", "dt_lin", "bt_lin"):
                alloc(nme, torch.float64)
        _lib.check(
            L.sdk_locate(
                handle, n, _lib.ptr(dx), _lib.ptr(dy), _lib.ptr(dz), _lib.ptr(dt), _lib.ptr(ts_dev), n_t,
                C.byref(loc), _lib.current_stream(),
            ),
            "sdk_locate",
        )
        out["_inputs"] = (dx, dy, dz, dt)
        word = torch.zeros(1, dtype=torch.int32, device=dev)  # OR of everybody's status bits
        _lib.check(L.sdk_status_or(out["status"].data_ptr(), n, word.data_ptr(), _lib.current_stream()), "sdk_status_or")
        flags = int(word.item())
        if flags & 16:
            raise ValueError("Value out of bound.")  # reference find_ind, utils.py:301
        out["_flags"] = flags
        return out

    @staticmethod
    def _np(t, dtype=None):
        a = t.cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def _find_rel_h(self, x, y):
        o = self.locate(x=x, y=y)
        face = self._np(o["face"], np.int64) if "face" in o else None
        ints = [self._np(o[k], np.int64) for k in ("iy", "ix")]
        rest = [self._np(o[k]) for k in ("rx", "ry", "cs", "sn", "dx", "dy", "bx", "by")]
        if self.readiness["h"] == "rectilinear":
            rest[2], rest[3] = rest[2].astype(np.float64), rest[3].astype(np.float64)
            rest[6], rest[7] = rest[6].astype(np.float64), rest[7].astype(np.float64)
        else:
            if self.CS is None:
                rest[2] = rest[3] = None
            if self.dX is None:
                rest[4] = rest[5] = None
        return HRel._make([face] + ints + rest)

    def _find_rel_v(self, z):
        o = self.locate(z=z)
        return VRel(self._np(o["iz"], np.int64), self._np(o["rz"]), self._np(o["dz"]), self._np(o["bz"]))

    def _find_rel_v_lin(self, z):
        o = self.locate(z=z)
        if o["_flags"] & 32:
            raise ValueError("Value out of bound.")
        return VLinRel(self._np(o["iz_lin"], np.int64), self._np(o["rz_lin"]), self._np(o["dz_lin"]), self._np(o["bz_lin"]))

    def _find_rel_t_lin(self, t):
        o = self.locate(t=t)
        if o["_flags"] & 64:
            raise ValueError("Value out of bound.")
        return TLinRel(self._np(o["it_lin"], np.int64), self._np(o["rt_lin"]), self._np(o["dt_lin"]), self._np(o["bt_lin"]))

    def _find_rel_vl(self, z):
        o = self.locate(z=z)
        return VlRel(
            self._np(o["izl_near"], np.int64), self._np(o["rzl_near"]), self._np(o["dzl_near"]), self._np(o["bzl_near"])
        )

    def _find_rel_vl_lin(self, z):
        o = self.locate(z=z)
        return VlLinRel(self._np(o["izl"], np.int64), self._np(o["rzl"]), self._np(o["dzl"]), self._np(o["bzl"]))

    def _find_rel_t(self, t):
        o = self.locate(t=t)
        return TRel(self._np(o["it"], np.int64), self._np(o["rt"]), self._np(o["dt"]), self._np(o["bt"]))
