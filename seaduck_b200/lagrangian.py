"""Lagrangian ``Particle``: a ``Position`` that knows how to move itself.

Same constructor and methods as the reference's ``seaduck.lagrangian.Particle``
(``lagrangian.py:56``): ``update_uvw_array`` :175, ``get_vol`` :202, ``get_u_du`` :218,
``note_taking`` :301 / ``empty_lists`` :355, ``to_next_stop`` :741, ``to_list_of_time`` :804,
``deepcopy`` :725.  The particle state lives on the GPU as a structure of arrays; one call of
``to_next_stop`` is one launch of the fused ``sdk_advect`` kernel.  numpy attributes
(``pt.lon``, ``pt.ix`` ...) are host mirrors fetched lazily after the device state changed.
"""

from __future__ import annotations

import ctypes as C
import logging
import warnings

import numpy as np

from . import _lib
from .eulerian import Position
from .kernel_weight import KnW
from .ocedata import RelCoord

DEG2M = 6271e3 * np.pi / 180  # sic (reference lagrangian.py:22)

uvkernel = np.array([[0, 0], [1, 0], [0, 1]])
ukernel = np.array([[0, 0], [1, 0]])
vkernel = np.array([[0, 0], [0, 1]])
wkernel = np.array([[0, 0]])
udoll = [[0, 1]]
vdoll = [[0, 2]]
wdoll = [[0]]
wknw = KnW(kernel=wkernel, inheritance=None, vkernel="linear", ignore_mask=True)
uknw = KnW(kernel=uvkernel, inheritance=udoll, ignore_mask=True)
vknw = KnW(kernel=uvkernel, inheritance=vdoll, ignore_mask=True)
dwknw = KnW(kernel=wkernel, inheritance=None, vkernel="dz", ignore_mask=True)
duknw = KnW(kernel=uvkernel, inheritance=udoll, hkernel="dx", h_order=1, ignore_mask=True)
dvknw = KnW(kernel=uvkernel, inheritance=vdoll, hkernel="dy", h_order=1, ignore_mask=True)
uknw_scalar = KnW(kernel=ukernel, inheritance=None, ignore_mask=True)
vknw_scalar = KnW(kernel=vkernel, inheritance=None, ignore_mask=True)
duknw_scalar = KnW(kernel=ukernel, inheritance=None, hkernel="dx", h_order=1, ignore_mask=True)
dvknw_scalar = KnW(kernel=vkernel, inheritance=None, hkernel="dy", h_order=1, ignore_mask=True)

# python attribute -> (device array name, lives in rel?, numpy dtype of the host mirror)
_STATE = {
    "lon": ("lon", False, np.float64),
    "lat": ("lat", False, np.float64),
    "dep": ("dep", False, np.float64),
    "t": ("t", False, np.float64),
    "face": ("face", True, np.int64),
    "iy": ("iy", True, np.int64),
    "ix": ("ix", True, np.int64),
    "izl_lin": ("izl", True, np.int64),
    "it": ("it", True, np.int64),
    "rx": ("rx", True, np.float64),
    "ry": ("ry", True, np.float64),
    "rzl_lin": ("rzl", True, np.float64),
    "u": ("u", False, np.float64),
    "v": ("v", False, np.float64),
    "w": ("w", False, np.float64),
    "du": ("du", False, np.float64),
    "dv": ("dv", False, np.float64),
    "dw": ("dw", False, np.float64),
    "px": ("px", False, np.float64),
    "py": ("py", False, np.float64),
    "dx": ("dx", True, np.float32),
    "dy": ("dy", True, np.float32),
    "bzl_lin": ("bzl", True, np.float64),
    "dzl_lin": ("dzl", True, np.float64),
    "vol": ("vol", False, np.float64),
}
_LIST_NAMES = {
    "fclist": "fc", "itlist": "it", "izlist": "izl", "rzlist": "rzl", "iylist": "iy", "ixlist": "ix",
    "rxlist": "rx", "rylist": "ry", "ttlist": "tt", "uulist": "uu", "vvlist": "vv", "wwlist": "ww",
    "dulist": "du", "dvlist": "dv", "dwlist": "dw", "xxlist": "xx", "yylist": "yy", "zzlist": "zz",
    "vslist": "vs",
}


class RecordSnapshot:
    """What a stop of ``to_list_of_time`` leaves behind when full clones of the particle state would not fit
    (``Particle.SNAPSHOT_CLONE_BUDGET``): position and cell of every particle in the caller's order, 32 bytes each
    (``sdk_snapshot_record``), kept in HBM until somebody reads them.  ``lon, lat, dep, face, iy, ix, izl_lin, t, N``
    read like the attributes of a ``Particle`` snapshot (reference ``lagrangian.py:725``); the rest of the state
    (rel-coords, velocities) is not kept."""

    _FIELDS = {"lon": "lon", "lat": "lat", "dep": "dep", "face": "face", "iy": "iy", "ix": "ix", "izl_lin": "izl"}

    def __init__(self, ocedata, t, records, has_face, has_dep):
        self.ocedata, self.N = ocedata, int(records.shape[0])
        self._t, self._records, self._host, self._event = float(t), records, None, None
        self._has_face, self._has_dep = has_face, has_dep

    @property
    def t(self):
        return np.full(self.N, self._t)

    def prefetch_to_host(self, stream=None):
        """Start the device-to-host copy of the records into pinned memory (on ``stream``) and return at once."""
        import torch  # noqa: PLC0415

        if self._host is None and self._event is None:
            host = torch.empty(self._records.shape, dtype=self._records.dtype).pin_memory()
            with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
                host.copy_(self._records, non_blocking=True)
                self._event = torch.cuda.Event()
                self._event.record()
            self._pinned = host

    def records(self):
        """numpy structured array of the records (``_lib.SNAPSHOT_RECORD``)."""
        if self._event is not None:
            self._event.synchronize()
            return self._pinned.numpy().reshape(-1).view(np.dtype(_lib.SNAPSHOT_RECORD))
        return self._records.cpu().numpy().reshape(-1).view(np.dtype(_lib.SNAPSHOT_RECORD))

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._FIELDS:
            raise AttributeError(
                f"'{name}': a record snapshot keeps lon, lat, dep, face, iy, ix, izl_lin only "
                "(to_list_of_time(..., snapshots='full') keeps the whole state per stop)"
            )
        if self._host is None:
            from .parallel import unpack_records  # noqa: PLC0415

            self.__dict__["_host"] = unpack_records(self.records())
        if (name == "face" and not self._has_face) or (name in ("izl_lin",) and not self._has_dep):
            return None
        return self._host[self._FIELDS[name]]


class Particle(Position):
    """Lagrangian particles advected on the GPU.  Parameters as in the reference (``lagrangian.py:93``).

    ``sort`` (not in the reference): keep the device state in cell order -- sorted by (level, face, iy, ix) with
    ``sdk_cell_sort`` -- so that the lanes of a warp read neighbouring memory.  ``"auto"`` sorts from
    ``SORT_THRESHOLD`` particles on (never with ``save_raw`` or a ``callback``, whose records / masks are in
    particle order).  Nothing changes for the caller: every attribute and snapshot comes back in the order the
    particles were given in.

    ``lazy = True`` (not in the reference) keeps ``to_next_stop`` asynchronous: the end-of-stop
    bookkeeping (``t := t_stop`` unless nothing was live, the time index, the crossing counters) is
    done on the device by ``sdk_finish_stop`` and the "maximum iteration count reached" warning is
    raised at the next :meth:`sync` / read of :attr:`crossings` instead of inside the call, so the
    host can queue the next stop, the snapshot exchange and the read-back while the GPU works.
    """

    lazy = False
    SORT_THRESHOLD = 100_000
    # to_list_of_time keeps full clones of the state per stop (what the reference returns) while they fit this many
    # bytes, 32-byte records per particle beyond that
    SNAPSHOT_CLONE_BUDGET = 8 << 30

    def __init__(
        self,
        uname="UVELMASS",
        vname="VVELMASS",
        wname="WVELMASS",
        free_surface="noflux",
        save_raw=False,
        transport=False,
        callback=None,
        max_iteration=200,
        sort="auto",
        **kwarg,
    ):
        Position.__init__(self)
        self.__dict__["_sort_request"] = sort
        if "bool_array" in kwarg.keys():
            self.from_bool_array(**kwarg)
        else:
            self.from_latlon(**kwarg)
        od = self.ocedata
        self.uname, self.vname, self.wname = uname, vname, wname
        self.wknw, self.dwknw = wknw, dwknw
        if self.face is not None:
            self.uknw, self.duknw, self.vknw, self.dvknw = uknw, duknw, vknw, dvknw
        else:
            self.uknw, self.duknw, self.vknw, self.dvknw = uknw_scalar, duknw_scalar, vknw_scalar, dvknw_scalar
        self.callback = callback
        self.transport = transport
        if self.transport:
            try:
                assert isinstance(od["Vol"], np.ndarray)
            except (AssertionError, KeyError):
                od._add_missing_vol(as_numpy=True)
        self.free_surface = free_surface
        if free_surface == "noflux" and wname is not None:
            logging.warning("Setting the surface velocity to zero. Dataset might be modified. ")
            od[self.wname].loc[{"Zl": 0}] = 0
            od._field_cache = {k: v for k, v in od._field_cache.items() if k[0] != self.wname}
        self.too_large = od.too_large
        self.max_iteration = max_iteration
        self.save_raw = save_raw
        self._has_dep = self.dep is not None and od.readiness["Zl"]
        self._init_device_state()
        self.update_uvw_array()
        if self.transport:
            self.get_vol()
        self.get_u_du()
        if self.save_raw:
            self.empty_lists()

    # ------------------------------------------------------------------------------------------
    # device state
    # ------------------------------------------------------------------------------------------
    def _init_device_state(self):
        import torch  # noqa: PLC0415

        od = self.ocedata
        dev = od._torch_device()
        n = self.N
        loc = self.__dict__.pop("_located")
        dlon, dlat, ddep, dt = loc["_inputs"]
        st = {}
        f64 = lambda: torch.zeros(n, dtype=torch.float64, device=dev)  # noqa: E731
        st["lon"], st["lat"] = dlon.clone(), dlat.clone()
        st["dep"] = ddep.clone() if ddep is not None else (
            torch.as_tensor(self.dep, dtype=torch.float64).to(dev) if self.dep is not None else f64()
        )
        st["t"] = dt.clone() if dt is not None else (
            torch.as_tensor(self.t, dtype=torch.float64).to(dev) if self.t is not None else f64()
        )
        st["face"] = loc["face"] if "face" in loc else None
        st["iy"], st["ix"] = loc["iy"], loc["ix"]
        st["rx"], st["ry"] = loc["rx"], loc["ry"]
        st["px"], st["py"] = loc["px"], loc["py"]
        # (schemes without corner points keep dx, dy as float64 rows of px / py)
        st["dx"], st["dy"] = (None, None) if od.readiness["h"] != "oceanparcel" else (loc["dx"], loc["dy"])
        if self._has_dep:
            st["izl"], st["rzl"], st["dzl"], st["bzl"] = loc["izl"], loc["rzl"], loc["dzl"], loc["bzl"]
        else:
            st["izl"] = torch.zeros(n, dtype=torch.int32, device=dev)
            st["rzl"], st["bzl"] = f64(), f64()
            st["dzl"] = torch.ones(n, dtype=torch.float64, device=dev)
        st["it"] = loc["it"] if "it" in loc else torch.zeros(n, dtype=torch.int32, device=dev)
        for name in ("u", "v", "w", "du", "dv", "dw"):
            st[name] = f64()
        st["vol"] = torch.ones(n, dtype=torch.float64, device=dev)
        for name in ("status", "niter", "ncross"):
            st[name] = torch.zeros(n, dtype=torch.int32, device=dev)
        self.__dict__["_perm"] = None  # device int32: state slot k holds the caller's particle _perm[k]
        self.__dict__["_inv"] = None   # device int64: the caller's particle i sits in state slot _inv[i]
        want = self.__dict__.pop("_sort_request", False)
        if want == "auto":
            want = n >= self.SORT_THRESHOLD
        if want and not self.save_raw and self.callback is None and n > 1:
            st = self._cell_sorted(st, n, dev)
        self.__dict__["_st"] = st
        self.__dict__["_crossings"] = 0
        self.__dict__["_stats"] = torch.zeros(4, dtype=torch.int64, device=dev)
        self.__dict__["_total"] = torch.zeros(4, dtype=torch.int64, device=dev)  # lazy mode: running totals
        self.__dict__["_warned"] = 0
        self.__dict__["_stale"] = set()
        self.__dict__["_raw"] = []
        self.__dict__["_local_h"] = od.readiness["h"] != "oceanparcel"
        from .parallel import check_grid_fits_record  # noqa: PLC0415

        check_grid_fits_record(od.tp, len(od.Zl) if od.readiness["Zl"] else 0)  # (the 32-byte output record's cell word)
        # host mirrors of what the constructor of the reference sets as float32 zeros
        for name in ("u", "v", "w", "du", "dv", "dw", "vol") + (() if self._local_h else ("px", "py")):
            self._stale.add(name)
        for key in ("izl_lin", "rzl_lin", "dzl_lin", "bzl_lin"):
            if not self._has_dep and key not in self.rel:
                self.rel[key] = None

    def _cell_sorted(self, st, n, dev):
        """Permute the freshly located state into cell order (``sdk_cell_sort`` + ``sdk_particles_gather``)."""
        import torch  # noqa: PLC0415

        L = _lib.lib()
        stream = _lib.current_stream()
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        nbytes = int(L.sdk_cell_sort_scratch_bytes(n))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        _lib.check(
            L.sdk_cell_sort(self.ocedata.grid_handle(), n, ptr(st.get("face")), ptr(st["iy"]), ptr(st["ix"]),
                            ptr(st["izl"]) if self._has_dep else None, perm.data_ptr(), scratch.data_ptr(), nbytes, stream),
            "sdk_cell_sort",
        )
        new = {k: (None if v is None else torch.empty_like(v)) for k, v in st.items()}
        src, dst = self._cparticles(state=st), self._cparticles(state=new)
        _lib.check(L.sdk_particles_gather(C.byref(src), C.byref(dst), perm.data_ptr(), stream), "sdk_particles_gather")
        inv = torch.empty(n, dtype=torch.int64, device=dev)
        inv[perm.long()] = torch.arange(n, dtype=torch.int64, device=dev)
        self.__dict__["_perm"], self.__dict__["_inv"] = perm, inv
        return new

    def _out_index(self):
        """int32 device tensor that puts packed records back into the caller's order, or None."""
        return self._perm

    def _cparticles(self, active=None, state=None):
        st = self._st if state is None else state
        p = _lib.Particles()
        p.n = self.N
        for name in _lib.PARTICLE_F64 + _lib.PARTICLE_I32 + _lib.PARTICLE_F64B + _lib.PARTICLE_F32 + _lib.PARTICLE_F64C + ["status", "niter", "ncross"]:
            t = st.get(name)
            setattr(p, name, None if t is None else t.data_ptr())
        p.active = None if active is None else active.data_ptr()
        return p

    def _pull(self, attr):
        """Refresh the host mirror of one attribute from the device."""
        dev_name, in_rel, dtype = _STATE[attr]
        t = self._st.get(dev_name)
        if self._local_h and attr in ("px", "py", "dx", "dy"):
            # schemes without corner points: no px / py on the particle (reference lagrangian.py:549), and the
            # device keeps bx, cs, dx / by, sn, dy as float64 rows of the px / py arrays (include/seaduck_b200.h)
            if attr in ("px", "py"):
                self._stale.discard(attr)
                raise AttributeError(attr)
            n = self.N
            t = self._st["px" if attr == "dx" else "py"][2 * n : 3 * n]
            dtype = np.float64 if self.ocedata.readiness["h"] == "rectilinear" else np.float32
        if t is None:
            val = None
        else:
            inv = self._inv
            if inv is not None:
                t = t.view(4, -1)[:, inv] if attr in ("px", "py") else t[inv]
            val = t.cpu().numpy().astype(dtype, copy=False)
            if attr in ("px", "py"):
                val = val.reshape(4, -1)
        if in_rel:
            self.rel[attr] = val
        else:
            self.__dict__[attr] = val
        self._stale.discard(attr)
        return val

    def _mark_stale(self, names=None):
        names = _STATE.keys() if names is None else names
        for attr in names:
            _, in_rel, _ = _STATE[attr]
            if attr in ("izl_lin", "rzl_lin", "dzl_lin", "bzl_lin") and not self._has_dep:
                continue
            if attr == "face" and self._st.get("face") is None:
                continue
            if attr == "it" and not self.ocedata.readiness["time"]:
                continue
            if attr in ("dx", "dy") and not self._local_h and self.ocedata.dX is None:
                continue
            if attr in ("px", "py") and self._local_h:
                continue
            self._stale.add(attr)
            if not in_rel:
                self.__dict__.pop(attr, None)
        # derived rel-coords are recomputed on demand
        self.__dict__["_derived_ok"] = False

    def __getattr__(self, attr):
        d = object.__getattribute__(self, "__dict__")
        if "_st" in d:
            if attr in _STATE and attr in d["_stale"]:
                return self._pull(attr)
            if attr in _LIST_NAMES:
                return self._raw_list(attr)
            if attr in ("bx", "by", "cs", "sn", "iz", "rz", "dz", "bz", "izl", "rzl", "dzl", "bzl", "rt", "dt", "bt") and not d.get("_derived_ok", True):
                self._refresh_derived()
        return Position.__getattr__(self, attr)

    def __getattribute__(self, attr):
        # rel-backed state attributes are found by __getattr__; plain ones that went stale were
        # removed from __dict__, so the normal path is unchanged and cheap
        return object.__getattribute__(self, attr)

    def __setattr__(self, attr, value):
        d = self.__dict__
        if "_st" in d and attr in _STATE and value is not None:
            import torch  # noqa: PLC0415

            dev_name, in_rel, dtype = _STATE[attr]
            tgt = d["_st"].get(dev_name)
            arr = np.ascontiguousarray(value)
            tdtype = tgt.dtype if tgt is not None else torch.float64
            new = torch.as_tensor(arr).to(device=self.ocedata._torch_device(), dtype=tdtype)
            if d.get("_perm") is not None:
                new = new[..., d["_perm"].long()]
            d["_st"][dev_name] = new.reshape(-1).contiguous()
            d["_stale"].discard(attr)
        Position.__setattr__(self, attr, value)

    def subset(self, which):
        """Host-side subset of the particles (reference ``eulerian.py:319``): what a ``callback`` receives.
        The host mirrors are refreshed first; the result holds numpy arrays only."""
        for attr in list(self._stale):
            self._pull(attr)
        if not self.__dict__.get("_derived_ok", True):
            self._refresh_derived()
        return Position.subset(self, which)

    def _rel_get(self, key):
        if key in self._stale:
            return self._pull(key)
        return self.rel.get(key)

    def _refresh_derived(self):
        """bx, by, cs, sn and the nearest-level vertical rel-coords follow the current cell / depth
        (the reference refreshes them in cross_cell_wall, lagrangian.py:640-686)."""
        od = self.ocedata
        self.__dict__["_derived_ok"] = True
        if self._local_h:
            # bx, cs / by, sn follow the particle on the device (rows 0, 1 of px / py)
            n = self.N
            as32 = od.readiness["h"] == "local_cartesian"  # float32 table values there, float64 arrays in rectilinear
            for key, (arr, row) in dict(bx=("px", 0), cs=("px", 1), by=("py", 0), sn=("py", 1)).items():
                val = self._st[arr][row * n : (row + 1) * n]
                val = (val if self._inv is None else val[self._inv]).cpu().numpy()
                self.rel[key] = val.astype(np.float32) if as32 else val
        else:
            face, iy, ix = self._rel_get("face"), self._rel_get("iy"), self._rel_get("ix")
            ind = (face, iy, ix) if face is not None else (iy, ix)
            self.rel["bx"] = od.XC[ind]
            self.rel["by"] = od.YC[ind]
        # cs, sn keep the values of the starting cell: the reference only refreshes bx, by, px, py in the
        # oceanparcel scheme (_cross_cell_wall_read, lagrangian.py:660-665), and vec_transform uses them
        if self._has_dep:
            # get_u_du leaves iz = izl_lin - 1 (lagrangian.py:247) while rz, dz, bz come from the nearest
            # cell centre of the current depth (_cross_cell_wall_rel, lagrangian.py:683)
            self.rel["iz"] = self._rel_get("izl_lin") - 1
            if od.readiness["Z"]:
                v = od._find_rel_v(self.dep)
                self.rel["rz"], self.rel["dz"], self.rel["bz"] = v.rz, v.dz, v.bz

    # ------------------------------------------------------------------------------------------
    # velocity staging
    # ------------------------------------------------------------------------------------------
    def update_uvw_array(self):
        """Stage the velocity slab needed by the particles on the device (reference :175)."""
        od = self.ocedata
        if "time" not in od[self.uname].dims:
            if "_vel" in self.__dict__:
                return
            time_slice = None
            it0 = 0
        else:
            it = self._st["it"]
            self.itmin = int(it.min().item())
            self.itmax = int(it.max().item())
            time_slice = slice(self.itmin, self.itmax + 1)
            it0 = self.itmin
        ut = od.device_field(self.uname, time_slice)
        vt = od.device_field(self.vname, time_slice)
        wt = od.device_field(self.wname, time_slice) if self.wname is not None else None
        udims, vdims = od[self.uname].dims, od[self.vname].dims
        v = _lib.Velocity()
        v.U, v.V = ut.data_ptr(), vt.data_ptr()
        v.W = wt.data_ptr() if wt is not None else None
        import torch  # noqa: PLC0415

        if ut.dtype != vt.dtype or (wt is not None and wt.dtype != ut.dtype):
            raise TypeError("u, v, w must share one storage dtype")
        v.dtype = 0 if ut.dtype == torch.float64 else 1
        v.has_time = int("time" in udims)
        v.it0 = it0
        v.nt = ut.shape[0] if v.has_time else 1
        v.has_z = int("Z" in udims)
        v.nzU = ut.shape[udims.index("Z")] if v.has_z else 1
        v.nyU, v.nxU = ut.shape[-2:]
        v.nyV, v.nxV = vt.shape[-2:]
        v.u_stag = int("Xp1" in udims)
        v.v_stag = int("Yp1" in vdims)
        v.nzW = wt.shape[1 if v.has_time else 0] if wt is not None else 1
        v.transport = int(self.transport)
        keep = [ut, vt, wt]
        if self.transport:
            vol = np.asarray(od["Vol"])
            vt_ = od.to_device("Vol", vol, np.float64 if vol.dtype != np.float32 else np.float32)
            v.Vol = vt_.data_ptr()
            v.vol_dtype = 0 if vt_.dtype == torch.float64 else 1
            v.vol_has_z = int(vol.ndim == len(od.tp.h_shape) + 1)
            keep.append(vt_)
        self.__dict__["_vel"] = v
        self.__dict__["_vel_keep"] = keep

    def _prefetch_slab_for(self, t_stop):
        """After a stop every particle sits at ``t_stop``, i.e. in one known time level: start the
        upload of that snapshot of u, v, w (streaming ingest)."""
        od = self.ocedata
        midp = np.asarray(od.time_midp, float)
        if t_stop < midp[0]:
            it = 0
        else:
            k = int(np.argmin(np.abs(midp - t_stop)))
            if midp[k] > t_stop:
                k -= 1
            it = max(k, 0) + 1
        sl = slice(it, it + 1)
        for name in (self.uname, self.vname, self.wname):
            if name is not None:
                od.prefetch_field(name, sl)

    def get_vol(self):
        """Cell volume at the particles (reference :202); done together with get_u_du on the device."""
        self._mark_stale(["vol"])

    def get_u_du(self):
        """Velocity and its gradient at the particles with the default kernels (reference :218)."""
        L = _lib.lib()
        p = self._cparticles()
        _lib.check(
            L.sdk_get_u_du(self.ocedata.grid_handle(), C.byref(self._vel), C.byref(p), int(self._has_dep), _lib.current_stream()),
            "sdk_get_u_du",
        )
        self._mark_stale(["u", "v", "w", "du", "dv", "dw", "vol"])

    def fillna(self):
        """nan -> 0 for the velocities: already part of the device kernels (reference :289)."""

    # ------------------------------------------------------------------------------------------
    # crossing log
    # ------------------------------------------------------------------------------------------
    def empty_lists(self):
        self.__dict__["_raw"] = []
        self.__dict__["_pending_start"] = False

    def note_taking(self, subset_index=None, stamp=-1):
        """Only the start-of-stop record (stamp 15) is taken from the host; the rest is written by
        the kernel (reference :301)."""
        if not self.save_raw:
            raise AttributeError("This is not a particle_rawlist object")
        if stamp != 15 or subset_index is not None:
            raise NotImplementedError("note_taking is done inside the advection kernel")
        self.__dict__["_pending_start"] = True

    @staticmethod
    def _host_chunk(ch):
        """Host view of one chunk of the log: numpy structured array + offsets (fetched once)."""
        if "host" not in ch:
            rec = ch["records"].cpu().numpy().view(np.dtype(_lib.LOG_RECORD)).reshape(-1)
            ch["host"] = (rec, ch["offset"].cpu().numpy())
        return ch["host"]

    def _raw_list(self, attr):
        if not self.save_raw:
            raise AttributeError("This is not a particle_rawlist object")
        key = _LIST_NAMES[attr]
        if key == "fc" and self._st.get("face") is None:
            raise AttributeError(attr)
        out = None
        for ch in self._raw:
            rec, off = self._host_chunk(ch)
            parts = np.split(rec[key], off[1:-1])  # one view per particle, no Python loop over the records
            out = [p.tolist() for p in parts] if out is None else [a + p.tolist() for a, p in zip(out, parts)]
        return out if out is not None else [[] for _ in range(self.N)]

    def raw_arrays(self, merged=False):
        """Crossing log of the last stop(s): per chunk (one per launch) a dict of flat numpy arrays
        (views into the 128-byte records) + per-particle ``offset``.  ``merged=True`` returns ONE such
        dict with the records of every particle concatenated in launch order (what the reference's
        per-particle lists hold after a stop that took several launches, e.g. with a ``callback``)."""
        out = []
        for ch in self._raw:
            rec, off = self._host_chunk(ch)
            d = {k: rec[k] for k in _lib.LOG_I32 + _lib.LOG_F64}
            d["offset"] = off
            out.append(d)
        if not merged:
            return out
        if len(out) == 1:
            return out[0]
        counts = [np.diff(d["offset"]) for d in out]
        total = np.sum(counts, axis=0)
        offset = np.zeros(self.N + 1, np.int64)
        np.cumsum(total, out=offset[1:])
        merged_rec = {k: np.empty(int(offset[-1]), out[0][k].dtype) for k in _lib.LOG_I32 + _lib.LOG_F64}
        start = offset[:-1].copy()
        for d, cnt in zip(out, counts):
            # destination of every record of this chunk: start of its particle + position inside the chunk
            pid = np.repeat(np.arange(self.N), cnt)
            dst = start[pid] + (np.arange(len(pid)) - d["offset"][:-1][pid])
            for k in merged_rec:
                merged_rec[k][dst] = d[k]
            start += cnt
        merged_rec["offset"] = offset
        return merged_rec

    # ------------------------------------------------------------------------------------------
    # time stepping
    # ------------------------------------------------------------------------------------------
    def _params(self, max_iteration):
        par = _lib.AdvectParams()
        par.tol = 1e-4
        par.trim_tol = 1e-14
        par.max_iteration = int(max_iteration)
        par.kick_back = int(self.free_surface == "kick_back")
        par.has_dep = int(self._has_dep)
        tune = getattr(self, "tuning", None) or {}
        par.refill_threshold = int(tune.get("refill_threshold", 0))
        par.blocks_per_sm = int(tune.get("blocks_per_sm", 0))
        par.threads_per_block = int(tune.get("threads_per_block", 0))
        par.reserve_sms = int(tune.get("reserve_sms", 0))
        par.stats = self._stats.data_ptr()
        return par

    def ship_to(self, exchange, order="caller"):
        """Multi-GPU runs (``parallel.RecordExchange``): the advection kernel of every leg whose end is an output stop
        also stores the particles' 32-byte output records into every rank's receive buffer (``sdk_advect_params.ship``).
        The caller waits for and releases each shipped step (``exchange.wait(step, release=True)``;
        ``last_ship_step`` says which).  ``order="caller"``: record i of the block is particle i as the caller numbered
        them; ``order="state"``: the records go out in the order of the device state (cell order when the particles are
        sorted -- contiguous stores; the receiver needs this rank's permutation, see ``parallel.ShardedParticle``).
        ``None`` switches it off."""
        if exchange is not None and (self.save_raw or self.callback is not None):
            raise ValueError("the fused output exchange does not go with save_raw or a callback")
        self.__dict__["_exchange"] = exchange
        self.__dict__["_ship_order"] = order
        self.__dict__["_ship_leg"] = exchange is not None
        self.__dict__["last_ship_step"] = None

    def _read_stats(self):
        """(live at start, crossings, hit max_iteration, flagged) of the launches since the last
        call -- one small device-to-host copy, the only synchronisation of a stop."""
        vals = self._stats.cpu().numpy().astype(np.int64)
        self._stats.zero_()
        return vals

    def _launch(self, t_stop, max_iteration, active=None, emit_start=False, continue_mode=0):
        import torch  # noqa: PLC0415

        L = _lib.lib()
        handle = self.ocedata.grid_handle()
        p = self._cparticles(active)
        par = self._params(max_iteration)
        stream = _lib.current_stream()
        if not self.save_raw:
            ex = self.__dict__.get("_exchange")
            if ex is not None and self.__dict__.get("_ship_leg", False):
                par.ship = C.addressof(ex.x)
                scatter = self._perm is not None and self.__dict__.get("_ship_order", "caller") == "caller"
                par.ship_index = self._perm.data_ptr() if scatter else None
                par.ship_step = ex.next_step()
                self.__dict__["last_ship_step"] = par.ship_step
            _lib.check(L.sdk_advect(handle, C.byref(self._vel), C.byref(p), t_stop, C.byref(par), None, 1, stream), "sdk_advect")
            return
        dev = self.ocedata._torch_device()
        cnt = torch.zeros(self.N, dtype=torch.int32, device=dev)
        lg = _lib.Log()
        lg.count = cnt.data_ptr()
        lg.emit_start = int(emit_start)
        lg.continue_mode = int(continue_mode)
        # pass 1: count the records of every particle without committing the state
        par_count = self._params(max_iteration)
        par_count.stats = None
        _lib.check(L.sdk_advect(handle, C.byref(self._vel), C.byref(p), t_stop, C.byref(par_count), C.byref(lg), 0, stream), "sdk_advect(count)")
        offset = torch.zeros(self.N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=offset[1:])
        total = int(offset[-1].item())
        # one array of 128-byte records (every record up to `total` is written by the fill pass)
        records = torch.empty((total, 128), dtype=torch.uint8, device=dev)
        lg.capacity = total
        lg.offset = offset.data_ptr()
        lg.records = records.data_ptr()
        # pass 2: same trajectory, now writing the records
        _lib.check(L.sdk_advect(handle, C.byref(self._vel), C.byref(p), t_stop, C.byref(par), C.byref(lg), 1, stream), "sdk_advect(log)")
        self._raw.append({"records": records, "offset": offset})

    def to_next_stop(self, t_stop):
        """Integrate all particles to ``t_stop`` (reference ``lagrangian.py:741``).

        One kernel launch, then one 32-byte read-back (how many particles were live, crossings,
        who ran out of iterations) that also orders the host with the device.
        """
        import torch  # noqa: PLC0415

        t_stop = float(t_stop)
        tol = 1e-4
        st = self._st
        emit_start = bool(self.__dict__.get("_pending_start", False))
        self.__dict__["_pending_start"] = False
        if self.lazy and self.callback is None and not self.save_raw:
            # no host round trip: one memset, the advection kernel, one bookkeeping kernel
            self._stats.zero_()
            self._launch(t_stop, self.max_iteration)
            od = self.ocedata
            midp = od.to_device("time_midp", od.time_midp, np.float64) if od.readiness["time"] else None
            _lib.check(
                _lib.lib().sdk_finish_stop(
                    st["t"].data_ptr(), st["it"].data_ptr() if midp is not None else None,
                    None if midp is None else midp.data_ptr(), 0 if midp is None else len(od.time_midp), t_stop,
                    self._stats.data_ptr(), self._total.data_ptr(), self.N, _lib.current_stream(),
                ),
                "sdk_finish_stop",
            )
            self._mark_stale()
            return
        if self.callback is None:
            self._launch(t_stop, self.max_iteration, emit_start=emit_start)
            n_live, n_cross, n_max, n_flag = self._read_stats()
            self.__dict__["_crossings"] += int(n_cross)
            if n_flag > n_max:
                self._raise_for_status()
            if n_live == 0:
                # the kernel touched nothing; like the reference, return before t := t_stop
                if self.save_raw and self._raw and not emit_start:
                    self._raw.pop()
                logging.info("Nothing left to simulate")
                return
            self._mark_stale()
            hit_max = n_max > 0
        else:
            # a host callback decides after every analytic step who keeps going
            todo = (t_stop - st["t"]).abs() > tol
            alive = todo.cpu().numpy() & np.asarray(self.callback(self), bool)
            if not alive.any():
                logging.info("Nothing left to simulate")
                return
            dev = self.ocedata._torch_device()
            hit_max = False
            for i in range(self.max_iteration):
                active = torch.as_tensor(alive.astype(np.int32)).to(dev)
                # the host decides who continues: the kernel leaves the stamp-7 record of a crossing to the
                # next launch, which writes it for the particles that are still active
                self._launch(t_stop, 1, active=active, emit_start=emit_start and i == 0, continue_mode=1 if i == 0 else 3)
                self.__dict__["_crossings"] += int(self._read_stats()[1])
                self._mark_stale()
                still = ((t_stop - st["t"]).abs() > tol).cpu().numpy()
                idx = np.where(alive)[0]
                keep = still[idx] & np.asarray(self.callback(self.subset(idx)), bool)
                alive = np.zeros_like(alive)
                alive[idx[keep]] = True
                if not alive.any():
                    break
            if i == self.max_iteration - 1:
                hit_max = True
                if self.save_raw and alive.any():
                    # out of iterations with particles still going: the reference has already logged
                    # their stamp-7 record (lagrangian.py:789-792); a zero-iteration launch writes it
                    active = torch.as_tensor(alive.astype(np.int32)).to(dev)
                    self._launch(t_stop, 0, active=active, continue_mode=3)
                    self._stats.zero_()
        if hit_max:
            warnings.warn("maximum iteration count reached")
        st["t"].fill_(t_stop)
        od = self.ocedata
        if od.readiness["time"]:
            midp = od.to_device("time_midp", od.time_midp, np.float64)
            _lib.check(
                _lib.lib().sdk_time_index(st["t"].data_ptr(), midp.data_ptr(), len(od.time_midp), st["it"].data_ptr(), self.N, _lib.current_stream()),
                "sdk_time_index",
            )
        self._mark_stale(["t", "it"])

    def _raise_for_status(self):
        """Status bits the reference answers with an exception: a particle that left through the deepest staged
        level (``Zl[izl_lin]`` raises IndexError, reference ``lagrangian.py:675``)."""
        if bool((self._st["status"] & 128).any().item()):  # SDK_ST_OOB_LEVEL
            raise IndexError("a particle left through the deepest vertical level (non-zero vertical velocity at the bottom)")

    def sync(self):
        """Lazy mode: wait for the queued stops, fold their counters in and raise deferred warnings."""
        tot = self._total.cpu().numpy().astype(np.int64)
        self._total.zero_()
        self.__dict__["_crossings"] += int(tot[1])
        if tot[2] > 0:
            warnings.warn("maximum iteration count reached")
        if tot[3] > tot[2]:
            self._raise_for_status()
        return tot

    @property
    def crossings(self):
        """Total number of cell-wall crossings so far (iterations that ended on a wall)."""
        if self.lazy:
            self.sync()
        return int(self._crossings)

    def deepcopy(self):
        """Snapshot of the particles (reference ``lagrangian.py:725``): device-side clone."""
        import torch  # noqa: PLC0415

        p = object.__new__(type(self))
        object.__setattr__(p, "rel", RelCoord())
        for key, item in self.__dict__.items():
            if key in ("_st", "_stale", "_raw", "_crossings", "_stats", "_total"):
                continue  # (_perm / _inv are shared: they never change)
            if isinstance(item, np.ndarray):
                if item.ndim == 1:
                    p.__dict__[key] = item.copy()
            elif key == "rel":
                continue
            else:
                p.__dict__[key] = item
        for k, v in self.rel.items():
            p.rel[k] = None if v is None else (v.copy() if isinstance(v, np.ndarray) else v)
        p.__dict__["_st"] = {k: (None if v is None else v.clone()) for k, v in self._st.items()}
        p.__dict__["_stale"] = set(self._stale)
        p.__dict__["_raw"] = list(self._raw)
        p.__dict__["_crossings"] = self._crossings
        p.__dict__["_stats"] = torch.zeros_like(self._stats)
        p.__dict__["_total"] = torch.zeros_like(self._stats)
        p.__dict__["_derived_ok"] = False
        return p

    def reserve_records(self, n_stops):
        """Set aside HBM for the 32-byte records of ``n_stops`` future stops in one allocation (``cudaMalloc`` of a
        block per ``to_list_of_time`` call is slow, and serialised across the processes of a node); the calls that
        follow take their blocks from it in order."""
        import torch  # noqa: PLC0415

        arena = torch.empty((int(n_stops), self.N, 32), dtype=torch.uint8, device=self.ocedata._torch_device())
        self.__dict__["_record_arena"] = [arena, 0]

    def _record_blocks(self, n_keep):
        """``n_keep`` blocks [N, 32] for record snapshots: from the reserved arena while it lasts, else a fresh one."""
        import torch  # noqa: PLC0415

        arena = self.__dict__.get("_record_arena")
        if arena is not None and arena[1] + n_keep <= arena[0].shape[0]:
            blocks = arena[0][arena[1] : arena[1] + n_keep]
            arena[1] += n_keep
            return blocks
        return torch.empty((max(n_keep, 1), self.N, 32), dtype=torch.uint8, device=self.ocedata._torch_device())

    def record_snapshot(self, t_stop, out=None):
        """Position and cell of every particle, now, as a :class:`RecordSnapshot` (one ``sdk_pack_records`` launch);
        ``out``: uint8 device tensor [N, 32] to pack into."""
        from .parallel import pack_records  # noqa: PLC0415

        rec = pack_records(self._cparticles(), self.N, out=out, out_index=self._out_index(), device=self.ocedata._torch_device())
        return RecordSnapshot(self.ocedata, t_stop, rec, self._st.get("face") is not None, self._has_dep)

    def to_list_of_time(
        self, normal_stops, update_stops="default", return_in_between=True, dump_filename=False, store_kwarg={},
        snapshots="auto", _on_stop=None,
    ):
        """Integrate to a list of times, returning a snapshot per stop (reference ``lagrangian.py:804``).

        ``snapshots`` (not in the reference): ``"full"`` returns clones of the whole particle state like the
        reference does, ``"records"`` returns :class:`RecordSnapshot` objects (32 bytes per particle and stop, in
        HBM), ``"auto"`` switches to records when the clones would take more than ``SNAPSHOT_CLONE_BUDGET`` bytes.
        """
        if snapshots not in ("auto", "full", "records"):
            raise ValueError("snapshots must be 'auto', 'full' or 'records'")
        od = self.ocedata
        t0 = float(self._st["t"][0].item())
        t_min = np.minimum(np.min(normal_stops), t0)
        t_max = np.maximum(np.max(normal_stops), t0)
        if not self.save_raw and dump_filename:
            raise ValueError("saving to file only available if save_raw = True")
        has_time = "time" in od[self.uname].dims
        if has_time:
            data_tmin = od.ts.min()
            data_tmax = od.ts.max()
            if t_min < data_tmin or t_max > data_tmax:
                raise ValueError("time range not within bound" + f"({data_tmin},{data_tmax})")
        if isinstance(update_stops, str) and update_stops == "default":
            if not has_time:
                update_stops = []
            else:
                try:
                    update_stops = od.time_midp[np.logical_and(t_min < od.time_midp, od.time_midp < t_max)]
                except AttributeError as exc:
                    raise AttributeError(
                        "time_midp is required for update_stops = default, but it is not in the dataset, "
                        "either create it or specify the update stops."
                    ) from exc
        if dump_filename:
            raise NotImplementedError("dumping the crossing log to zarr is post-processing (SURVEY §8f-1)")
        temp = list(zip(normal_stops, np.zeros_like(normal_stops))) + list(zip(update_stops, np.ones_like(update_stops)))
        temp.sort(key=lambda x: abs(x[0] - t0))
        stops, update = list(zip(*temp))
        self.get_u_du()
        to_return = []
        state_bytes = sum(v.numel() * v.element_size() for v in self._st.values() if v is not None)
        if snapshots == "auto":
            snapshots = "records" if (state_bytes * len(stops) > self.SNAPSHOT_CLONE_BUDGET and not self.save_raw) else "full"

        ring = None
        if snapshots == "records" and _on_stop is None:
            # one allocation for the records of every stop that is kept (a fresh 32 N-byte block per stop would go to
            # the driver every time: cudaMalloc is slow and serialised across the processes of a node)
            n_keep = sum(1 for i in range(len(stops)) if return_in_between or not (update[i] or i == 0))
            ring = self._record_blocks(n_keep)

        def keep(t_stop):
            if _on_stop is not None:
                _on_stop(t_stop)
            elif snapshots == "records":
                to_return.append(self.record_snapshot(t_stop, out=ring[len(to_return)]))
            else:
                to_return.append(self.deepcopy())

        for i, tl in enumerate(stops):
            timestr = str(np.datetime64(round(tl), "s"))
            logging.info(timestr)
            if self.save_raw:
                self.note_taking(stamp=15)
            if self.__dict__.get("_exchange") is not None:
                # the fused output exchange ships the legs that end in a stop somebody keeps
                self.__dict__["_ship_leg"] = return_in_between or not (update[i] or i == 0)
            self.to_next_stop(tl)
            if has_time and self.lazy and (update[i] or i == 0):
                # the leg is queued, not finished: copy the snapshot it ends in while it runs
                self._prefetch_slab_for(tl)
            if update[i] or i == 0:
                if has_time:
                    self.update_uvw_array()
                    self.get_u_du()
                if return_in_between:
                    keep(tl)
            else:
                keep(tl)
            if self.save_raw:
                self.empty_lists()
        return stops, to_return
