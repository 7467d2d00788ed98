"""TEST INFRASTRUCTURE -- CPU restatement (numpy + the C oracle's topology walk) of the reference's
Eulerian path: ``Position.from_latlon`` (``seaduck/eulerian.py:138``) and ``Position.interpolate``
(``eulerian.py:1268``) with its stages ``_fatten_h`` :386, ``_fatten_v/_vl/_t`` :451-518,
``_transform_vector_and_register`` :887, ``_mask_value_and_register`` :959 (``get_masks.py:12-110,169``),
``_read_data_and_register`` :1064, ``_compute_weight_and_register`` :1161 and the weights of
``kernel_weight.py:84-428,481-584,772-874``.

Parity status: PINNED -- ``tests/test_interp_oracle.py`` checks it against the golden vectors made by
running the unmodified reference (``oracle/make_golden_interp.py`` -> ``tests/golden/interp_*.npz``).
Only ``tests/`` and ``bench_interp.py``'s cpu-baseline leg may import this; the product never does.
A ``KnW`` here is any object with the reference's attribute surface (``kernel, inheritance, hkernel,
h_order, vkernel, tkernel, ignore_mask``) -- it is only read as data.
"""

from __future__ import annotations

import ctypes as C
from itertools import combinations
from math import factorial

import numpy as np

from . import oracle as ora


# ------------------------------------------------------------------------------------------------
# topology (C oracle)
# ------------------------------------------------------------------------------------------------
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fatten_h(og, face, iy, ix, kernel, cuvwg):
    """eulerian.py:386 -> (faces | None, iys, ixs, err), each (N, M)."""
    lib = ora.lib()
    n, m = len(iy), len(kernel)
    kx, ky = _i32(kernel[:, 0]), _i32(kernel[:, 1])
    f = _i32(face) if face is not None else np.zeros(n, np.int32)
    yy, xx = _i32(iy), _i32(ix)
    of, oy, ox, oe = (np.zeros((n, m), np.int32) for _ in range(4))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    lib.ora_fatten_h.argtypes = [C.c_void_p, C.c_int64] + [C.POINTER(C.c_int32)] * 3 + [C.c_int] + [C.POINTER(C.c_int32)] * 2 + [C.c_char] + [C.POINTER(C.c_int32)] * 4
    lib.ora_fatten_h(C.byref(og.c), n, p(f), p(yy), p(xx), m, p(kx), p(ky), C.c_char(cuvwg.encode()), p(of), p(oy), p(ox), p(oe))
    return (of.astype(np.int64) if og.has_face else None), oy.astype(np.int64), ox.astype(np.int64), oe


def four_matrix_for_uv(og, faces):
    """topology.py:689 with the rounding of eulerian.py:934-937."""
    lib = ora.lib()
    n, m = faces.shape
    f = _i32(faces)
    out = [np.zeros((n, m)) for _ in range(4)]
    err = np.zeros((n, m), np.int32)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    lib.ora_four_matrix_for_uv.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int32)]
    lib.ora_four_matrix_for_uv(C.byref(og.c), n, m, f.ctypes.data_as(C.POINTER(C.c_int32)), *(pd(a) for a in out), err.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, err


# ------------------------------------------------------------------------------------------------
# Position.from_latlon
# ------------------------------------------------------------------------------------------------
class OraclePosition:
    """Indices and rel-coords of points (eulerian.py:138; ocedata.py:422-488)."""

    def __init__(self, od, x=None, y=None, z=None, t=None):
        self.od = od
        self.og = ora.OracleGrid(od)
        og = self.og
        self.N = max(len(v) for v in (x, y, z, t) if v is not None)
        self.face = self.iy = self.ix = self.rx = self.ry = self.cs = self.sn = None
        if x is not None and y is not None:
            self.lon, self.lat = np.asarray(x, float), np.asarray(y, float)
            if og.readiness["h"] == "rectilinear":
                # find_rel_h_rectilinear, utils.py:602: two periodic 1-D lookups, grid aligned with the meridians
                ix, self.rx, _, _ = ora.find_rel(self.lon, og.lon, peri=360.0)
                iy, self.ry, _, _ = ora.find_rel(self.lat, og.lat, peri=360.0)
                self.iy, self.ix = np.asarray(iy, np.int64), np.asarray(ix, np.int64)
                self.cs, self.sn = np.ones_like(self.lon), np.zeros_like(self.lon)
            else:
                ind = og.find_ind_h(self.lon, self.lat)
                if og.has_face:
                    self.face, self.iy, self.ix = (np.asarray(a, np.int64) for a in ind)
                else:
                    self.iy, self.ix = (np.asarray(a, np.int64) for a in ind)
                if og.readiness["h"] == "local_cartesian":
                    # find_rel_h_naive, utils.py:577 (find_rx_ry_naive :522 is compiled: float64 cosine of a float32 YC)
                    bx, by = og.XC[ind].astype(float), og.YC[ind].astype(float)
                    cs, sn = og.CS[ind].astype(float), og.SN[ind].astype(float)
                    dlon, dlat = ora.to_180(self.lon - bx), ora.to_180(self.lat - by)
                    cosby = np.cos(by * np.pi / 180)
                    self.rx = (dlon * cosby * cs + dlat * sn) * ora.DEG2M / og.dX[ind].astype(float)
                    self.ry = (dlat * cs - dlon * sn * cosby) * ora.DEG2M / og.dY[ind].astype(float)
                else:
                    px, py = og.px_py(ind, self.lon)
                    self.rx, self.ry = ora.find_rx_ry_oceanparcel(self.lon, self.lat, px, py)
                if og.CS is not None:
                    self.cs, self.sn = og.CS[ind], og.SN[ind]
        self.dep = None if z is None else np.asarray(z, float)
        self.t = None if t is None else np.asarray(t, float)
        rd = og.readiness
        nothing = (None, None, None, None)
        self.iz, self.rz, _, _ = ora.find_rel(self.dep, og.Z, above=False) if (z is not None and rd["Z"]) else nothing
        self.izl, self.rzl, _, _ = ora.find_rel(self.dep, og.Zl, above=False) if (z is not None and rd["Zl"]) else nothing
        self.izl_lin, self.rzl_lin, _, _ = (
            ora.find_rel(self.dep, og.Zl, darray=og.dZl, ascending=-1, dx_right=True) if (z is not None and rd["Zl"]) else nothing
        )
        self.it, self.rt, _, _ = ora.find_rel(self.t, og.ts, above=False) if (t is not None and rd["time"]) else nothing
        self._lin = {}

    def lin(self, which):
        """Lazily evaluated linear rel-coords (eulerian.py:465,508)."""
        if which not in self._lin:
            og = self.og
            if which == "z":
                self._lin[which] = ora.find_rel(self.dep, og.Z, darray=og.dZ, ascending=-1, dx_right=False)[:2]
            else:
                self._lin[which] = ora.find_rel(self.t, og.ts)[:2]
        return self._lin[which]


# ------------------------------------------------------------------------------------------------
# weights (kernel_weight.py)
# ------------------------------------------------------------------------------------------------
def _axis_weight(nodes, target, r, order):
    """Lagrange weight of ``target`` among 1-D ``nodes`` for the ``order``-th derivative
    (kernel_weight.py:136-230: sums over subsets of the other nodes / product of differences; at
    the maximum order the numerator is (order - 1)!)."""
    others = [o for o in nodes if o != target]
    denom = 1.0
    for o in others:
        denom *= target - o
    if order == len(nodes) - 1:
        return np.full(np.shape(r), float(factorial(max(order - 1, 0)))) / denom
    num = np.zeros(np.shape(r))
    for sub in combinations(others, len(others) - order):
        term = np.ones(np.shape(r))
        for o in sub:
            term = term * (r - o)
        num = num + term
    return num / denom


def stencil_weight(kernel, hkernel, h_order, rx, ry):
    """kernel_weight (kernel_weight.py:431): cross (:84) or rectangular (:313) kernels -> (N, M)."""
    kernel = np.asarray(kernel).astype(int)
    xs = sorted({int(v) for v in kernel[:, 0]})
    ys = sorted({int(v) for v in kernel[:, 1]})
    m = len(kernel)
    w = np.zeros((len(rx), m))
    if m == len(xs) + len(ys) - 1:  # cross
        typ = hkernel[1:] if "d" in hkernel else hkernel
        if len(xs) == 1:
            typ = "y"
        elif len(ys) == 1:
            typ = "x"
        for i, (x, y) in enumerate(kernel):
            if typ == "interp":
                if x != 0:
                    w[:, i] = _axis_weight(xs, x, rx, 0)
                elif y != 0:
                    w[:, i] = _axis_weight(ys, y, ry, 0)
                else:
                    w[:, i] = _axis_weight(xs, 0, rx, 0) + _axis_weight(ys, 0, ry, 0) - 1
            elif typ == "x":
                if y == 0:
                    w[:, i] = _axis_weight(xs, x, rx, h_order)
            elif typ == "y":
                if x == 0:
                    w[:, i] = _axis_weight(ys, y, ry, h_order)
    elif m == len(xs) * len(ys):  # rectangle
        xo, yo = {"interp": (0, 0), "dx": (h_order, 0), "dy": (0, h_order)}[hkernel]
        for i, (x, y) in enumerate(kernel):
            w[:, i] = _axis_weight(xs, x, rx, xo) * _axis_weight(ys, y, ry, yo)
    else:
        raise NotImplementedError("The shape of the kernel is neither cross or square")
    return w


def cascade_weight(knw, rx, ry, masked):
    """get_weight_cascade (kernel_weight.py:525) for one (N, M) wet matrix, or the full stencil."""
    kernel = np.asarray(knw.kernel).astype(int)
    n, m = len(rx), len(kernel)
    weight = np.zeros((n, m))
    if masked is None:
        doll = list(knw.inheritance[0])
        weight[:, doll] = stencil_weight(kernel[doll], knw.hkernel, knw.h_order, rx, ry)
        return weight
    weight[:, 0] = np.nan
    chosen = np.full(n, -1)
    for level, doll in enumerate(knw.inheritance):
        wet = masked[:, list(doll)].all(axis=1)
        chosen[(chosen < 0) & wet] = level  # the largest stencil that is all wet (:481-523)
    for level, doll in enumerate(knw.inheritance):
        sel = np.where(chosen == level)[0]
        if len(sel) == 0:
            continue
        sub = np.zeros((len(sel), m))
        sub[:, list(doll)] = stencil_weight(kernel[list(doll)], knw.hkernel, knw.h_order, rx[sel], ry[sel])
        weight[sel] = sub
    return weight


def get_weight(knw, rx, ry, rz, rt, masked, bottom_scheme):
    """KnW.get_weight (kernel_weight.py:772) -> (N, M, nz, nt); ``masked`` is (N, M, nz, nt) or None."""
    nz = 2 if knw.vkernel in ("linear", "dz") else 1
    nt = 2 if knw.tkernel in ("linear", "dt") else 1
    n = len(rx)
    weight = np.zeros((n, len(knw.kernel), nz, nt))
    if np.isscalar(rz):
        rz = np.full(n, float(rz))
    if np.isscalar(rt):
        rt = np.full(n, float(rt))
    tweight = {"linear": [(1 - rt).reshape(n, 1, 1), rt.reshape(n, 1, 1)], "dt": [-1, 1], "nearest": [1]}[knw.tkernel]
    rpz = np.array(rz, float, copy=True)
    zweight = {"linear": [(1 - rpz).reshape(n, 1), rpz.reshape(n, 1)], "dz": [-1, 1], "nearest": [1]}[knw.vkernel]
    for jt in range(nt):
        for jz in range(nz):
            weight[:, :, jz, jt] = cascade_weight(knw, rx, ry, None if masked is None else masked[:, :, jz, jt])
    for jt in range(nt):
        if knw.vkernel == "linear" and bottom_scheme == "no flux":
            second = np.isnan(weight[:, :, 0, jt]).any(axis=1)
            weight[second, :, 0, jt] = 0
            weight[np.logical_and(second, rz < 1 / 2), :, 1, jt] = 0
            zweight[1][second] = 1
        for jz in range(nz):
            weight[:, :, jz, jt] *= zweight[jz]
        weight[:, :, :, jt] *= tweight[jt]
    return weight


# ------------------------------------------------------------------------------------------------
# masks (get_masks.py)
# ------------------------------------------------------------------------------------------------
class OracleMasks:
    def __init__(self, od, og):
        self.od, self.og = od, og
        ds = od._ds
        self.has = "maskC" in list(ds.keys())
        self.cache = {}
        if self.has:
            self.maskC = np.asarray(ds["maskC"])
            self.mdims = tuple(ds["maskC"].dims)
            if "Z" not in self.mdims:
                raise NotImplementedError("2D land mask")

    def _neighbour(self, node):
        og = self.og
        shape = og.h_shape
        idx = np.indices(shape).reshape(len(shape), -1)
        face = idx[0] if og.has_face else None
        ff, fy, fx, _ = fatten_h(og, face, idx[-2], idx[-1], np.array([node]), "C")
        if og.has_face:
            return ff[:, 0].reshape(shape), fy[:, 0].reshape(shape), fx[:, 0].reshape(shape)
        return fy[:, 0].reshape(shape), fx[:, 0].reshape(shape)

    def get(self, cuvw):
        """maskC | mask_u_node :12 | mask_v_node :46 | mask_w_node :79, zero-padded like get_masked :195-214."""
        if cuvw in self.cache:
            return self.cache[cuvw]
        ds = self.od._ds
        mc = self.maskC
        if cuvw == "C":
            out = mc
        elif "mask" + cuvw in list(ds.keys()):
            out = np.asarray(ds["mask" + cuvw])
        else:
            if cuvw in ("U", "V"):
                nb = self._neighbour((-1, 0) if cuvw == "U" else (0, -1))
                other = mc[(slice(None),) + tuple(nb)]
                small = np.where(mc == 0, (other != 0).astype(mc.dtype), mc)
            else:
                small = np.zeros_like(mc)
                small[1:] = mc[:-1]
                small = np.logical_or(small, mc).astype(int)
            rename = {
                "U": lambda d: d if d != "X" else "Xp1",
                "V": lambda d: d if d != "Y" else "Xp1",
                "Wvel": lambda d: d if d != "Z" else "Zl",
            }[cuvw]
            sizes = tuple(len(ds[rename(d)]) for d in self.mdims)
            out = np.zeros(sizes)
            out[tuple(slice(0, s) for s in small.shape)] = small
        self.cache[cuvw] = out
        return out


def _mask_kind(dims):
    if "Zl" in dims:
        return "Wvel"
    if "Z" in dims:
        if "Xp1" in dims and "Yp1" in dims:
            raise NotImplementedError("corner mask")
        if "Xp1" in dims:
            return "U"
        if "Yp1" in dims:
            return "V"
    return "C"


# ------------------------------------------------------------------------------------------------
# Position.interpolate
# ------------------------------------------------------------------------------------------------
class InterpOracle:
    def __init__(self, od, pos):
        self.od, self.pos, self.og = od, pos, pos.og
        self.masks = OracleMasks(od, self.og)

    # -- index sets ----------------------------------------------------------------------------------
    def _vertical_time(self, dims, knw):
        """(kz (N, nz) | None, rz, kt (N, nt) | None, rt) -- eulerian.py:451-518 and :1196-1224."""
        p = self.pos
        kz = rz = kt = rt = None
        if "Z" in dims or "Zl" in dims:
            if "Z" in dims:
                near, rnear = p.iz, p.rz
                lin = p.lin("z") if knw.vkernel != "nearest" else None
            else:
                near, rnear = p.izl, p.rzl
                lin = (p.izl_lin, p.rzl_lin) if knw.vkernel != "nearest" else None
            if knw.vkernel == "nearest":
                kz, rz = near.reshape(-1, 1), rnear
            else:
                kz, rz = np.vstack([lin[0], lin[0] - 1]).T, lin[1]
        if "time" in dims:
            if knw.tkernel == "nearest":
                kt, rt = p.it.reshape(-1, 1), p.rt
            else:
                it_lin, rt_lin = p.lin("t")
                kt, rt = np.vstack([it_lin, it_lin + 1]).T, rt_lin
        return kz, rz, kt, rt

    @staticmethod
    def _gather(arr, dims, kt, kz, ff, fy, fx):
        """arr[(t), (z), (face), y, x] broadcast to (N, M, nz, nt) (smart_read on in-memory data)."""
        n, m = fy.shape
        nz = 1 if kz is None else kz.shape[1]
        nt = 1 if kt is None else kt.shape[1]
        idx = []
        if "time" in dims:
            idx.append(kt.reshape(n, 1, 1, nt))
        if "Z" in dims or "Zl" in dims:
            idx.append(kz.reshape(n, 1, nz, 1))
        if "face" in dims:
            idx.append(ff.reshape(n, m, 1, 1))
        idx += [fy.reshape(n, m, 1, 1), fx.reshape(n, m, 1, 1)]
        out = arr[tuple(idx)]
        return np.broadcast_to(out, (n, m, nz, nt))

    def _wet(self, cuvw, dims, kz, ff, fy, fx):
        """get_masked (get_masks.py:169) at the fattened indices -> (N, M, nz, 1); a variable without a
        vertical dimension is masked with level 0 (_ind_for_mask, eulerian.py:82)."""
        mask = self.masks.get(cuvw)
        n, m = fy.shape
        z = np.zeros((n, 1), np.int64) if kz is None else kz
        idx = [z.reshape(n, 1, -1, 1)]
        if self.og.has_face:
            idx.append(ff.reshape(n, m, 1, 1))
        idx += [fy.reshape(n, m, 1, 1), fx.reshape(n, m, 1, 1)]
        return mask[tuple(idx)]

    # -- one scalar ----------------------------------------------------------------------------------
    def scalar(self, name, knw, arr=None):
        od, p = self.od, self.pos
        dims = tuple(od[name].dims)
        arr = np.asarray(od[name]) if arr is None else arr
        cuvwg = "G" if ("Xp1" in dims or "Yp1" in dims) else "C"
        kernel = np.asarray(knw.kernel).astype(int)
        ff, fy, fx, err = fatten_h(self.og, p.face, p.iy, p.ix, kernel, cuvwg)
        kz, rz, kt, rt = self._vertical_time(dims, knw)
        bad = err.any(axis=1)
        sfy, sfx = np.where(err != 0, 0, fy), np.where(err != 0, 0, fx)
        sff = None if ff is None else np.where(err != 0, 0, ff)
        data = np.nan_to_num(self._gather(arr, dims, kt, kz, sff, sfy, sfx))
        masked = None
        long_dims = "".join(dims)
        if not knw.ignore_mask and "X" in long_dims and "Y" in long_dims and self.masks.has:
            masked = np.broadcast_to(self._wet(_mask_kind(dims), dims, kz, sff, sfy, sfx), data.shape)
        rx = p.rx + 0.5 if "Xp1" in dims else p.rx
        ry = p.ry + 0.5 if "Yp1" in dims else p.ry
        bottom = "no_flux" if "Z" in dims else None  # sic: never equals "no flux" (eulerian.py:1193)
        weight = get_weight(knw, rx, ry, 0 if rz is None else rz, 0 if rt is None else rt, masked, bottom)
        out = np.einsum("nijk,nijk->n", data, weight)
        out[bad] = np.nan  # the reference raises for these points
        return out

    # -- one horizontal vector on a grid with faces --------------------------------------------------
    def vector(self, names, knws, vec_transform=True):
        od, p = self.od, self.pos
        uname, vname = names
        uknw, vknw = knws
        udims, vdims = tuple(od[uname].dims), tuple(od[vname].dims)
        uarr, varr = np.asarray(od[uname]), np.asarray(od[vname])
        kz, rz, kt, rt = self._vertical_time(udims, uknw)
        res = []
        bad = np.zeros(p.N, bool)
        ignore = uknw.ignore_mask or vknw.ignore_mask
        for which, knw in (("U", uknw), ("V", vknw)):
            kernel = np.asarray(knw.kernel).astype(int)
            ff, fy, fx, err = fatten_h(self.og, p.face, p.iy, p.ix, kernel, which)
            sfy, sfx = np.where(err != 0, 0, fy), np.where(err != 0, 0, fx)
            sff = np.where(err != 0, p.face[:, None], ff)
            (ufu, ufv, vfu, vfv), err2 = four_matrix_for_uv(self.og, sff)
            bad |= err.any(axis=1) | err2.any(axis=1)
            own, other = (ufu, ufv) if which == "U" else (vfv, vfu)
            from_own = np.abs(own).astype(bool)
            from_other = np.abs(other).astype(bool)
            d_u = np.nan_to_num(self._gather(uarr, udims, kt, kz, sff, np.minimum(sfy, uarr.shape[-2] - 1), np.minimum(sfx, uarr.shape[-1] - 1)))
            d_v = np.nan_to_num(self._gather(varr, vdims, kt, kz, sff, np.minimum(sfy, varr.shape[-2] - 1), np.minimum(sfx, varr.shape[-1] - 1)))
            d_own, d_other = (d_u, d_v) if which == "U" else (d_v, d_u)
            temp = np.full(d_own.shape, np.nan)
            temp[from_own] = d_own[from_own]
            temp[from_other] = d_other[from_other]
            data = temp * own[:, :, None, None] + temp * other[:, :, None, None]  # eulerian.py:1152-1157
            masked = None
            if not ignore and self.masks.has:
                m_u = self._wet("U", udims, kz, sff, sfy, sfx)
                m_v = self._wet("V", vdims, kz, sff, sfy, sfx)
                m_own, m_other = (m_u, m_v) if which == "U" else (m_v, m_u)
                mk = np.full(m_own.shape, np.nan)
                mk[from_own] = m_own[from_own]
                mk[from_other] = m_other[from_other]
                masked = np.broadcast_to(mk, data.shape)
            rx = p.rx + 0.5 if which == "U" else p.rx
            ry = p.ry if which == "U" else p.ry + 0.5
            weight = get_weight(knw, rx, ry, 0 if rz is None else rz, 0 if rt is None else rt, masked, "no flux")
            res.append(np.einsum("nijk,nijk->n", data, weight))
        u, v = res
        if vec_transform:
            cs, sn = p.cs.astype(float), p.sn.astype(float)
            u, v = u * cs - v * sn, u * sn + v * cs  # local_to_latlon, utils.py:206
        u[bad] = np.nan
        v[bad] = np.nan
        return u, v

    # -- the public call -----------------------------------------------------------------------------
    def interpolate(self, var_name, knw, vec_transform=True):
        """Same nesting rules as the reference for the cases the tests use (str | tuple | list)."""
        if isinstance(var_name, list):
            knws = knw if isinstance(knw, list) else [knw] * len(var_name)
            return [self.interpolate(v, k, vec_transform) for v, k in zip(var_name, knws)]
        if isinstance(var_name, tuple):
            if self.pos.face is None:
                return tuple(self.scalar(v, k) for v, k in zip(var_name, knw))
            return self.vector(var_name, knw, vec_transform)
        return self.scalar(var_name, knw)
