"""TEST INFRASTRUCTURE -- Python face of the CPU oracle.

``OracleGrid`` / ``OracleParticle`` wrap ``libsd_oracle.so`` (plain C restatement of the reference's
Lagrangian loop, ``oracle/sd_oracle.c``) and restate in numpy the parts of the reference that set a
particle up (``Position.from_latlon`` -> ``find_rel_h_oceanparcel`` / ``find_rel``; reference
``seaduck/eulerian.py:138``, ``seaduck/utils.py:321,614``).  Parity status: PINNED against golden
vectors generated from the unmodified reference (``oracle/make_golden.py`` ->
``tests/golden/*.npz``) and the reference's own known-answer tests.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline legs may import this.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)


class _Grid(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("typ", "has_face", "nface", "ny", "nx", "gny", "gnx", "hmode", "n_zl", "n_z")]
        + [("face_connect", c_i32p)]
        + [(n, c_f32p) for n in ("XC", "YC", "XG", "YG", "dX", "dY", "CS", "SN", "Zl", "dZl", "Z", "dZ", "lon", "lat")]
        + [("dlon", c_f64p), ("dlat", c_f64p), ("coslat32", c_f32p)]
    )


class _Vel(C.Structure):
    _fields_ = (
        [("U", C.c_void_p), ("V", C.c_void_p), ("W", C.c_void_p)]
        + [
            (n, C.c_int32)
            for n in (
                "dtype", "has_time", "it0", "nt", "has_z", "nzU", "nzW",
                "nyU", "nxU", "nyV", "nxV", "u_stag", "v_stag", "has_w", "transport",
            )
        ]
        + [("Vol", C.c_void_p), ("vol_dtype", C.c_int32), ("vol_has_z", C.c_int32)]
    )


_P_F64 = ("lon", "lat", "dep", "t")
_P_I32 = ("face", "iy", "ix", "izl", "it")
_P_F64B = ("rx", "ry", "rzl", "u", "v", "w", "du", "dv", "dw", "px", "py", "dx", "dy", "bzl", "dzl", "vol", "bx", "by", "cs", "sn")
_P_OUT = ("niter", "ncross", "status")


class _Particles(C.Structure):
    _fields_ = (
        [("n", C.c_int64)]
        + [(n, c_f64p) for n in _P_F64]
        + [(n, c_i32p) for n in _P_I32]
        + [(n, c_f64p) for n in _P_F64B]
        + [(n, c_i32p) for n in _P_OUT]
    )


class _Params(C.Structure):
    _fields_ = [
        ("tol", C.c_double),
        ("trim_tol", C.c_double),
        ("max_iteration", C.c_int32),
        ("kick_back", C.c_int32),
        ("has_dep", C.c_int32),
        ("nthreads", C.c_int32),
    ]


_LOG_I32 = ("fc", "it", "izl", "iy", "ix", "vs")
_LOG_F64 = ("rzl", "rx", "ry", "tt", "uu", "vv", "ww", "du", "dv", "dw", "xx", "yy", "zz")


class _Log(C.Structure):
    _fields_ = (
        [("capacity", C.c_int64), ("offset", c_i64p), ("count", c_i32p)]
        + [(n, c_i32p) for n in _LOG_I32]
        + [(n, c_f64p) for n in _LOG_F64]
        + [("emit_start", C.c_int32)]
    )


def build_oracle(force=False):
    """Compile ``libsd_oracle.so`` with gcc (a few hundred ms)."""
    so = os.path.join(_HERE, "libsd_oracle.so")
    src = os.path.join(_HERE, "sd_oracle.c")
    hdr = os.path.join(_HERE, "sd_oracle.h")
    stale = not os.path.exists(so) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr) if os.path.exists(f)
    )
    if force or stale:
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so, src, "-lm", "-lpthread"],
            cwd=_HERE,
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_oracle())
        _LIB.ora_advect.restype = C.c_int
        _LIB.ora_ind_tend.restype = C.c_int
        _LIB.ora_ind_moves.restype = C.c_int
        _LIB.ora_corners.restype = C.c_int
    return _LIB


def _ptr(arr, typ):
    if arr is None:
        return typ()
    return arr.ctypes.data_as(typ)


_TYP = {"box": 0, "x_periodic": 1, "LLC": 2}
_HMODE = {"oceanparcel": 0, "local_cartesian": 1, "rectilinear": 2}

DEG2M = 6271e3 * np.pi / 180  # sic, reference lagrangian.py:22 / utils.py:306


# ------------------------------------------------------------------------------------------------
# numpy restatements of reference helpers used while setting a particle up
# ------------------------------------------------------------------------------------------------
def to_180(x, peri=360):
    """reference utils.py:199"""
    x = x % peri % peri
    return x + (-1) * (x // (peri / 2)) * peri


def spherical2cartesian(lat, lon, R=6371.0):
    """reference utils.py:166"""
    y_rad = np.deg2rad(lat)
    x_rad = np.deg2rad(lon)
    return R * np.cos(y_rad) * np.cos(x_rad), R * np.cos(y_rad) * np.sin(x_rad), R * np.sin(y_rad)


def find_rel(value, array, darray=None, ascending=1, above=True, peri=None, dx_right=True):
    """reference utils.py:321 (with find_ind :273 inlined)."""
    array = np.asarray(array)
    if darray is None:
        darray = np.abs(array[1:] - array[:-1])
        darray = np.append(darray, darray[-1])
        dx_right = True
    n = len(value)
    ix_arr = np.zeros(n, dtype="int")
    rx_arr = np.zeros(n)
    dx_arr = np.zeros(n)
    bx_arr = np.zeros(n)
    dx_offset = int(not dx_right) + int(ascending > 0) - 1
    for i, x in enumerate(value):
        if peri is None:
            index = int(np.argmin(np.abs(array - x)))
        else:
            index = int(np.argmin(np.abs(to_180(array - x, peri=peri))))
        if above and array[index] > x and peri is None:
            index -= ascending * 1
        if index < 0 or index >= len(array):
            raise ValueError("Value out of bound.")
        bx = array[index]
        dx = x - bx if peri is None else to_180(x - bx, peri=peri)
        if not above or peri is not None:
            dx_offset = int(not dx_right) + int(ascending * dx > 0) - 1
        j = index + dx_offset
        ix_arr[i] = index
        rx_arr[i] = dx / darray[j]
        dx_arr[i] = darray[j]
        bx_arr[i] = bx
    return ix_arr, rx_arr, dx_arr, bx_arr


def find_rx_ry_oceanparcel(x, y, px, py):
    """reference utils.py:532"""
    x0 = px[0]
    x = to_180(x - x0)
    px = to_180(px - x0)
    a = [px[0], -px[0] + px[1], -px[0] + px[3], ((px[0] - px[1]) + px[2]) - px[3]]
    b = [py[0], -py[0] + py[1], -py[0] + py[3], ((py[0] - py[1]) + py[2]) - py[3]]
    with np.errstate(all="ignore"):
        aa = a[3] * b[2] - a[2] * b[3]
        bb = a[3] * b[0] - a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + x * b[3] - y * a[3]
        cc = a[1] * b[0] - a[0] * b[1] + x * b[1] - y * a[1]
        det2 = bb * bb - 4 * aa * cc
        order1 = np.abs(aa) < 1e-12
        order2 = np.logical_and(~order1, det2 >= 0)
        ry = -(cc / bb)
        ry[order2] = ((-bb + np.sqrt(det2)) / (2 * aa))[order2]
        rot_rectilinear = np.abs(a[1] + a[3] * ry) < 1e-12
        rx = np.zeros_like(ry)
        rx[rot_rectilinear] = (
            (y - py[0]) / (py[1] - py[0]) + (y - py[3]) / (py[2] - py[3])
        )[rot_rectilinear] * 0.5
        rx[~rot_rectilinear] = ((x - a[0] - a[2] * ry) / (a[1] + a[3] * ry))[~rot_rectilinear]
    return rx - 1 / 2, ry - 1 / 2


class OracleGrid:
    """Grid arrays in the reference's in-memory form (``OceData._grid2array``, ocedata.py:411)."""

    def __init__(self, od):
        """``od``: anything with the reference OceData attribute surface (the reference's own
        ``OceData`` or the product's host object): ``XC YC XG YG dX dY CS SN Z dZ Zl dZl ts
        time_midp tp readiness``."""
        self.od = od
        tp = od.tp
        self.typ = tp.typ
        self.readiness = dict(od.readiness)
        self.h_shape = tuple(tp.h_shape)
        self.g_shape = tuple(tp.g_shape)
        self.has_face = len(self.h_shape) == 3
        get = lambda n: getattr(od, n, None)  # noqa: E731
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
        self.XC, self.YC, self.XG, self.YG = (f32(get(n)) for n in ("XC", "YC", "XG", "YG"))
        self.dX, self.dY, self.CS, self.SN = (f32(get(n)) for n in ("dX", "dY", "CS", "SN"))
        self.Z = f32(get("Z")) if self.readiness["Z"] else None
        self.dZ = f32(get("dZ")) if self.readiness["Z"] else None
        self.Zl = f32(get("Zl")) if self.readiness["Zl"] else None
        self.dZl = f32(get("dZl")) if self.readiness["Zl"] else None
        self.ts = np.asarray(od.ts, float) if self.readiness["time"] else None
        self.time_midp = np.asarray(od.time_midp, float) if self.readiness["time"] else None
        fc = getattr(tp, "face_connect", None)
        self.face_connect = None if fc is None else np.ascontiguousarray(fc, dtype=np.int32)
        self._tree = None
        g = _Grid()
        g.typ = _TYP[self.typ]
        g.has_face = int(self.has_face)
        g.nface = self.h_shape[0] if self.has_face else 1
        g.ny, g.nx = self.h_shape[-2:]
        g.gny, g.gnx = self.g_shape[-2:]
        g.hmode = _HMODE[self.readiness["h"]]
        g.n_zl = 0 if self.Zl is None else len(self.Zl)
        g.n_z = 0 if self.dZl is None else len(self.dZl)
        g.face_connect = _ptr(self.face_connect, c_i32p)
        for n in ("XC", "YC", "XG", "YG", "dX", "dY", "CS", "SN", "Zl", "dZl", "Z", "dZ"):
            setattr(g, n, _ptr(getattr(self, n), c_f32p))
        if self.readiness["h"] == "rectilinear":
            # ocedata.py:255-259, 368-370: lon / lat as float32, their spacing in metres from np.gradient (R = 6371 km)
            self.lon, self.lat = f32(get("lon")), f32(get("lat"))
            self.dlon = np.ascontiguousarray(get("dlon"), dtype=np.float64)
            self.dlat = np.ascontiguousarray(get("dlat"), dtype=np.float64)
            # what lagrangian.py:671 / :693 evaluate for a float32 latitude (float32 arithmetic and numpy's float32
            # cosine): taken from numpy on the same array, element by element
            self.coslat32 = np.ascontiguousarray(np.cos(self.lat * np.pi / 180), dtype=np.float32)
            assert self.coslat32.dtype == np.float32 and (self.lat * np.pi / 180).dtype == np.float32
            g.lon, g.lat = _ptr(self.lon, c_f32p), _ptr(self.lat, c_f32p)
            g.dlon, g.dlat = _ptr(self.dlon, c_f64p), _ptr(self.dlat, c_f64p)
            g.coslat32 = _ptr(self.coslat32, c_f32p)
        elif self.readiness["h"] == "local_cartesian":
            self.coslat32 = np.ascontiguousarray(np.cos(self.YC * np.pi / 180), dtype=np.float32)
            assert (self.YC * np.pi / 180).dtype == np.float32
            g.coslat32 = _ptr(self.coslat32, c_f32p)
        self.c = g

    # -- topology restatement exposed for tests -------------------------------------------------
    def ind_tend(self, ind, tend, cuvwg="C"):
        i3 = (C.c_int32 * 3)(*([0] * (3 - len(ind)) + [int(i) for i in ind]))
        out = (C.c_int32 * 3)()
        err = lib().ora_ind_tend(C.byref(self.c), i3, int(tend), C.c_char(cuvwg.encode()), out)
        return err, tuple(out)[3 - len(ind):]

    def ind_moves(self, ind, moves, cuvwg="C"):
        i3 = (C.c_int32 * 3)(*([0] * (3 - len(ind)) + [int(i) for i in ind]))
        out = (C.c_int32 * 3)()
        mv = (C.c_int * max(1, len(moves)))(*moves)
        err = lib().ora_ind_moves(C.byref(self.c), i3, mv, len(moves), C.c_char(cuvwg.encode()), out)
        return err, tuple(out)[3 - len(ind):]

    def corners(self, ind):
        i3 = (C.c_int32 * 3)(*([0] * (3 - len(ind)) + [int(i) for i in ind]))
        out = ((C.c_int32 * 3) * 4)()
        err = lib().ora_corners(C.byref(self.c), i3, out)
        return err, [tuple(o)[3 - len(ind):] for o in out]

    # -- Position.from_latlon restatement ----------------------------------------------------------
    def tree(self):
        """reference utils.py:225 create_tree (scipy cKDTree on float32 XC/YC, R = 6371 km)."""
        if self._tree is None:
            from scipy import spatial  # noqa: PLC0415

            x, y, z = spherical2cartesian(lat=self.YC, lon=self.XC)
            pts = np.vstack((x.ravel(), y.ravel(), z.ravel())).T
            pts = np.nan_to_num(pts, nan=777777)  # the reference's "ridiculous value" for NaN centres (utils.py:243-253)
            self._tree = spatial.cKDTree(pts, leafsize=16)
        return self._tree

    def find_ind_h(self, lons, lats):
        """reference utils.py:309"""
        x, y, z = spherical2cartesian(lats, lons)
        _, index1d = self.tree().query(np.array([x, y, z]).T)
        return np.unravel_index(index1d, self.h_shape)

    def px_py(self, ind, lon):
        """reference utils.py:495 find_px_py + eulerian.py:618"""
        n = len(ind[0])
        px = np.zeros((4, n))
        py = np.zeros((4, n))
        k = len(self.h_shape)
        for i in range(n):
            err, cs = self.corners(tuple(int(a[i]) for a in ind))
            if err:
                raise ValueError(f"cannot find the corners of cell {tuple(int(a[i]) for a in ind)}")
            for j in range(4):
                px[j, i] = self.XG[cs[j][-k:]]
                py[j, i] = self.YG[cs[j][-k:]]
        px = lon + to_180(px - lon)
        return px, py


class OracleParticle:
    """Restates ``Particle.__init__`` + ``to_next_stop`` + ``to_list_of_time`` (lagrangian.py:93,741,804)."""

    def __init__(
        self,
        grid,
        x,
        y,
        z,
        t,
        uarr,
        varr,
        warr=None,
        vol=None,
        u_dims=None,
        v_dims=None,
        transport=False,
        free_surface="noflux",
        max_iteration=200,
        save_raw=False,
        it0=0,
        nthreads=1,
    ):
        self.grid = grid
        g = grid
        self.N = n = len(x)
        self.transport = transport
        self.free_surface = free_surface
        self.max_iteration = max_iteration
        self.save_raw = save_raw
        self.nthreads = nthreads
        self.lon = np.array(x, float)
        self.lat = np.array(y, float)
        self.t = np.array(t, float)
        self.has_dep = z is not None and g.readiness["Zl"]
        self.dep = np.array(z, float) if z is not None else np.zeros(n)
        if g.readiness["h"] == "rectilinear":
            # reference find_rel_h_rectilinear, utils.py:602 (two periodic 1-D lookups; spacing from np.diff, R = 6271 km)
            ix, rx, dxd, bx = find_rel(self.lon, g.lon, peri=360.0)
            iy, ry, dyd, by = find_rel(self.lat, g.lat, peri=360.0)
            self.face = np.zeros(n, np.int32)
            self.iy, self.ix = np.ascontiguousarray(iy, np.int32), np.ascontiguousarray(ix, np.int32)
            self.rx, self.ry, self.bx, self.by = rx, ry, bx, by
            self.dx = np.cos(by * np.pi / 180) * DEG2M * dxd
            self.dy = DEG2M * dyd
            self.cs, self.sn = np.ones(n), np.zeros(n)
            self.px, self.py = np.zeros((4, n)), np.zeros((4, n))
        elif g.readiness["h"] == "local_cartesian":
            # reference find_rel_h_naive, utils.py:577 (find_rx_ry_naive :522 is compiled: float64 cosine)
            ind = g.find_ind_h(self.lon, self.lat)
            if g.has_face:
                self.face, self.iy, self.ix = (np.ascontiguousarray(a, np.int32) for a in ind)
            else:
                self.face = np.zeros(n, np.int32)
                self.iy, self.ix = (np.ascontiguousarray(a, np.int32) for a in ind)
            self.bx, self.by = g.XC[ind].astype(float), g.YC[ind].astype(float)
            self.cs, self.sn = g.CS[ind].astype(float), g.SN[ind].astype(float)
            self.dx, self.dy = g.dX[ind].astype(float), g.dY[ind].astype(float)
            dlon, dlat = to_180(self.lon - self.bx), to_180(self.lat - self.by)
            cosby = np.cos(self.by * np.pi / 180)
            self.rx = (dlon * cosby * self.cs + dlat * self.sn) * DEG2M / self.dx
            self.ry = (dlat * self.cs - dlon * self.sn * cosby) * DEG2M / self.dy
            self.px, self.py = np.zeros((4, n)), np.zeros((4, n))
        else:
            self._init_h_oceanparcel(g, n)
        if self.has_dep:
            # reference OceData._find_rel_vl_lin, ocedata.py:475
            izl, rzl, dzl, bzl = find_rel(self.dep, g.Zl, g.dZl, ascending=-1, dx_right=True)
            self.izl = izl.astype(np.int32)
            self.rzl, self.dzl, self.bzl = rzl, dzl.astype(float), bzl.astype(float)
        else:
            self.izl = np.zeros(n, np.int32)
            self.rzl, self.dzl, self.bzl = np.zeros(n), np.ones(n), np.zeros(n)
        if g.readiness["time"]:
            it, _, _, _ = find_rel(self.t, g.ts, above=False)
            self.it = it.astype(np.int32)
        else:
            self.it = np.zeros(n, np.int32)
        (self.u, self.v, self.w, self.du, self.dv, self.dw) = (np.zeros(n) for _ in range(6))
        self.vol = np.ones(n)
        self.niter = np.zeros(n, np.int32)
        self.ncross = np.zeros(n, np.int32)
        self.status = np.zeros(n, np.int32)
        self.set_velocity(uarr, varr, warr, vol, u_dims, v_dims, it0)
        if transport:
            self.get_vol()
        self.get_u_du()
        self.records = None

    def _init_h_oceanparcel(self, g, n):
        # horizontal rel-coords: reference find_rel_h_oceanparcel, utils.py:614
        ind = g.find_ind_h(self.lon, self.lat)
        if g.has_face:
            self.face, self.iy, self.ix = (np.ascontiguousarray(a, np.int32) for a in ind)
        else:
            self.face = np.zeros(n, np.int32)
            self.iy, self.ix = (np.ascontiguousarray(a, np.int32) for a in ind)
        self.bx = g.XC[ind].astype(float)
        self.by = g.YC[ind].astype(float)
        self.dx = g.dX[ind].astype(float) if g.dX is not None else np.ones(n)
        self.dy = g.dY[ind].astype(float) if g.dY is not None else np.ones(n)
        self.cs = g.CS[ind].astype(float) if g.CS is not None else np.ones(n)
        self.sn = g.SN[ind].astype(float) if g.SN is not None else np.zeros(n)
        px, py = g.px_py(ind, self.lon)
        self.rx, self.ry = find_rx_ry_oceanparcel(self.lon, self.lat, px, py)
        self.px, self.py = np.ascontiguousarray(px), np.ascontiguousarray(py)

    # -- velocity staging -------------------------------------------------------------------------
    def set_velocity(self, uarr, varr, warr, vol, u_dims, v_dims, it0=0):
        g = self.grid
        self._keep = [np.ascontiguousarray(a) if a is not None else None for a in (uarr, varr, warr, vol)]
        uarr, varr, warr, vol = self._keep
        v = _Vel()
        v.dtype = 0 if uarr.dtype == np.float64 else 1
        if uarr.dtype not in (np.float64, np.float32):
            raise TypeError("velocity must be float32 or float64")
        v.U = uarr.ctypes.data
        v.V = varr.ctypes.data
        v.W = warr.ctypes.data if warr is not None else None
        v.has_time = int("time" in u_dims)
        v.it0 = it0
        v.nt = uarr.shape[0] if v.has_time else 1
        v.has_z = int("Z" in u_dims)
        v.nzU = uarr.shape[u_dims.index("Z")] if v.has_z else 1
        v.nyU, v.nxU = uarr.shape[-2:]
        v.nyV, v.nxV = varr.shape[-2:]
        v.u_stag = int("Xp1" in u_dims)
        v.v_stag = int("Yp1" in v_dims)
        v.has_w = int(warr is not None)
        v.nzW = warr.shape[1 if v.has_time else 0] if warr is not None else 1
        v.transport = int(self.transport)
        if vol is not None:
            v.Vol = vol.ctypes.data
            v.vol_dtype = 0 if vol.dtype == np.float64 else 1
            v.vol_has_z = int(vol.ndim == len(g.h_shape) + 1)
        self.cvel = v

    def _cparticles(self):
        p = _Particles()
        p.n = self.N
        for n in _P_F64 + _P_F64B:
            setattr(p, n, _ptr(getattr(self, n), c_f64p))
        for n in _P_I32 + _P_OUT:
            setattr(p, n, _ptr(getattr(self, n), c_i32p))
        return p

    def get_vol(self):
        """reference lagrangian.py:202"""
        g = self.grid
        vol = self._keep[3]
        index = []
        if vol.ndim == len(g.h_shape) + 1:
            index.append(self.izl - 1)
        if g.has_face:
            index.append(self.face)
        index += [self.iy, self.ix]
        self.vol = np.ascontiguousarray(vol[tuple(index)], float)

    def get_u_du(self):
        p = self._cparticles()
        f = lib().ora_get_u_du
        for i in range(self.N):
            f(C.byref(self.grid.c), C.byref(self.cvel), C.byref(p), C.c_int64(i), int(self.has_dep))

    def to_next_stop(self, t_stop, emit_start=False):
        """reference lagrangian.py:741; returns the number of wall crossings."""
        par = _Params(1e-4, 1e-14, self.max_iteration, int(self.free_surface == "kick_back"), int(self.has_dep), self.nthreads)
        p = self._cparticles()
        log = None
        if self.save_raw:
            # pass 1: count records on a scratch copy of the state
            saved = {n: getattr(self, n).copy() for n in _P_F64 + _P_I32 + _P_F64B}
            cnt = np.zeros(self.N, np.int32)
            lg = _Log()
            lg.count = _ptr(cnt, c_i32p)
            lg.emit_start = int(emit_start)
            lib().ora_advect(C.byref(self.grid.c), C.byref(self.cvel), C.byref(p), C.c_double(t_stop), C.byref(par), C.byref(lg))
            for n, a in saved.items():
                getattr(self, n)[...] = a
            offset = np.zeros(self.N + 1, np.int64)
            np.cumsum(cnt, out=offset[1:])
            total = int(offset[-1])
            rec = {n: np.zeros(total, np.int32) for n in _LOG_I32}
            rec.update({n: np.zeros(total) for n in _LOG_F64})
            lg = _Log()
            lg.capacity = total
            lg.offset = _ptr(offset, c_i64p)
            lg.count = _ptr(cnt, c_i32p)
            lg.emit_start = int(emit_start)
            for n in _LOG_I32:
                setattr(lg, n, _ptr(rec[n], c_i32p))
            for n in _LOG_F64:
                setattr(lg, n, _ptr(rec[n], c_f64p))
            rec["offset"] = offset
            self.records = rec
            log = C.byref(lg)
            self._lg = lg
        self.status[:] = 0
        if not (np.abs(t_stop - self.t) > 1e-4).any():
            return 0  # "Nothing left to simulate": the reference returns before touching t / it
        err = lib().ora_advect(C.byref(self.grid.c), C.byref(self.cvel), C.byref(p), C.c_double(t_stop), C.byref(par), log)
        if err:
            raise RuntimeError(f"oracle advect failed: {err}")
        # reference lagrangian.py:795-802
        self.t = np.ones(self.N) * t_stop
        g = self.grid
        if g.readiness["time"]:
            before_first = self.t < g.time_midp[0]
            self.it[before_first] = 0
            if (~before_first).any():
                it, _, _, _ = find_rel(self.t[~before_first], g.time_midp)
                self.it[~before_first] = it + 1
        return int(self.ncross.sum())


def oracle_particle_from_od(od, x, y, z, t, uname="UVELMASS", vname="VVELMASS", wname="WVELMASS", transport=False,
                            save_raw=False, nthreads=1, **kw):
    """Build an :class:`OracleParticle` the way ``Particle.__init__`` does (reference lagrangian.py:93):
    Vol from drF*rA*hFacC when transport, W surface level zeroed for ``free_surface="noflux"``,
    velocity slab ``[itmin:itmax+1]`` when the data has a time axis."""
    grid = OracleGrid(od)
    full_u = np.asarray(od[uname])
    full_v = np.asarray(od[vname])
    full_w = np.asarray(od[wname]) if wname is not None else None
    u_dims, v_dims = tuple(od[uname].dims), tuple(od[vname].dims)
    if full_w is not None and kw.get("free_surface", "noflux") == "noflux":
        full_w = full_w.copy()
        if "time" in u_dims:
            full_w[:, 0] = 0
        else:
            full_w[0] = 0
    vol = None
    if transport:
        try:
            vol = np.asarray(od["Vol"])
        except (KeyError, AttributeError):
            od._add_missing_vol(as_numpy=True)
            vol = np.asarray(od["Vol"])
    has_time = "time" in u_dims
    # first pass to know `it`
    op = OracleParticle.__new__(OracleParticle)
    op._full = (full_u, full_v, full_w, vol, u_dims, v_dims, has_time)
    if has_time:
        it, _, _, _ = find_rel(np.asarray(t, float), grid.ts, above=False)
        itmin, itmax = int(it.min()), int(it.max())
        sl = slice(itmin, itmax + 1)
        ua, va, wa = full_u[sl], full_v[sl], None if full_w is None else full_w[sl]
    else:
        itmin = 0
        ua, va, wa = full_u, full_v, full_w
    OracleParticle.__init__(op, grid, x, y, z, t, ua, va, wa, vol, u_dims, v_dims, transport=transport,
                            save_raw=save_raw, it0=itmin, nthreads=nthreads,
                            free_surface=kw.get("free_surface", "noflux"), max_iteration=kw.get("max_iteration", 200))
    return op


def oracle_update_uvw(op):
    """reference Particle.update_uvw_array (lagrangian.py:175) + get_u_du."""
    full_u, full_v, full_w, vol, u_dims, v_dims, has_time = op._full
    if not has_time:
        return
    itmin, itmax = int(op.it.min()), int(op.it.max())
    sl = slice(itmin, itmax + 1)
    op.set_velocity(full_u[sl], full_v[sl], None if full_w is None else full_w[sl], vol, u_dims, v_dims, it0=itmin)
    op.get_u_du()


def oracle_to_list_of_time(op, normal_stops, update_stops="default"):
    """reference Particle.to_list_of_time (lagrangian.py:804); yields (stop, records, state) per snapshot."""
    g = op.grid
    has_time = op._full[6]
    t0 = op.t[0]
    t_min = min(np.min(normal_stops), t0)
    t_max = max(np.max(normal_stops), t0)
    if isinstance(update_stops, str):
        update_stops = g.time_midp[(t_min < g.time_midp) & (g.time_midp < t_max)] if has_time else []
    temp = [(s, 0) for s in normal_stops] + [(s, 1) for s in update_stops]
    temp.sort(key=lambda x: abs(x[0] - t0))
    op.get_u_du()
    out = []
    for i, (tl, upd) in enumerate(temp):
        op.to_next_stop(tl, emit_start=True)
        if upd or i == 0:
            oracle_update_uvw(op)
        rec = None if op.records is None else {k: v.copy() for k, v in op.records.items()}
        state = {k: getattr(op, k).copy() for k in _P_F64 + _P_I32 + ("rx", "ry", "rzl", "u", "v", "w", "du", "dv", "dw")}
        out.append((tl, rec, state))
    return out
