"""Eulerian ``Position``: where the points are on the grid, and interpolation at those points.

Public surface of the reference's ``seaduck.eulerian.Position`` (``eulerian.py:110``):
``from_latlon`` (:138), ``subset`` (:319), ``update_from_subset`` (:353), ``get_px_py`` (:592),
``get_f_node_weight`` (:621), ``interpolate`` (:1268).  The work is done by CUDA kernels
(``sdk_locate``, ``sdk_interp``); numpy arrays on the object are host mirrors.
"""

from __future__ import annotations

import copy
import logging

import numpy as np

from .ocedata import Deferred, HRel, OceData, RelCoord, TRel, VlLinRel, VlRel, VRel, _general_len


def to_180(x, peri=360):
    """Convert any longitude scale to [-180,180) (reference ``utils.py:199``)."""
    x = x % peri % peri
    return x + (-1) * (x // (peri / 2)) * peri


def weight_f_node(rx, ry):
    """Bilinear weights of the four corners LL, LR, UR, UL (reference ``utils.py:632``)."""
    return np.vstack(
        [
            (0.5 - rx) * (0.5 - ry),
            (0.5 + rx) * (0.5 - ry),
            (0.5 + rx) * (0.5 + ry),
            (0.5 - rx) * (0.5 + ry),
        ]
    ).T


_LAZY_REL = {}
for _n in ("iz_lin", "rz_lin", "dz_lin", "bz_lin"):
    _LAZY_REL[_n] = ("_find_rel_v_lin", "dep")
for _n in ("it_lin", "rt_lin", "dt_lin", "bt_lin"):
    _LAZY_REL[_n] = ("_find_rel_t_lin", "t")


class Position:
    """Points on the grid.  Create empty, then ``from_latlon``."""

    def __new__(cls, *arg, **kwarg):
        new_position = object.__new__(cls)
        object.__setattr__(new_position, "rel", RelCoord())
        return new_position

    def __getattr__(self, attr):
        # only called when normal lookup fails
        if attr == "rel":
            raise AttributeError(attr)
        rel = object.__getattribute__(self, "rel")
        if attr in rel.keys():
            return rel[attr]
        if attr in _LAZY_REL:
            # linear Z / time rel-coords are looked up on first use (reference eulerian.py:465,508)
            func, src = _LAZY_REL[attr]
            coord = self.__dict__.get(src)
            if coord is not None:
                rel.update(getattr(self.ocedata, func)(coord))
                return rel[attr]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{attr}'")

    def __setattr__(self, attr, value):
        if attr == "rel":
            object.__setattr__(self, attr, value)
        elif attr in self.rel.keys():
            self.rel[attr] = value
            # the device copy made by from_latlon no longer describes these points
            self.__dict__.pop("_located", None)
        else:
            object.__setattr__(self, attr, value)

    def from_latlon(self, x=None, y=None, z=None, t=None, data=None):
        """Fill in the coordinate info from lat-lon-dep-time (reference ``eulerian.py:138``)."""
        if data is not None:
            if not isinstance(data, OceData):
                raise ValueError("Input data must be OceData")
            self.ocedata = data
        elif "ocedata" not in self.__dict__:
            raise ValueError("data not provided")
        self.tp = self.ocedata.tp
        x, y, z, t = self._coordinate_arrays(x, y, z, t)
        od = self.ocedata
        # one fused device pass for all four coordinates
        loc = od.locate(
            x=x if (x is not None and y is not None) else None,
            y=y if (x is not None and y is not None) else None,
            z=z if (z is not None and (od.readiness["Z"] or od.readiness["Zl"])) else None,
            t=t if (t is not None and od.readiness["time"]) else None,
        ) if any(v is not None for v in (x, y, z, t)) else {}
        self._located = loc  # device tensors, reused by Particle
        npf = Deferred  # host mirrors are fetched on first use
        if (x is not None) and (y is not None):
            self.lon = np.array(x, float)
            self.lat = np.array(y, float)
            face = npf(loc["face"], np.int64) if "face" in loc else None
            vals = [npf(loc[k], np.int64) for k in ("iy", "ix")] + [
                npf(loc[k]) for k in ("rx", "ry", "cs", "sn", "dx", "dy", "bx", "by")
            ]
            if od.readiness["h"] == "rectilinear":
                # find_rel_h_rectilinear (utils.py:602) returns float64 arrays throughout
                for j in (4, 5, 8, 9):
                    vals[j] = npf(vals[j].tensor, np.float64)
            else:
                if od.CS is None:
                    vals[4] = vals[5] = None
                if od.dX is None:
                    vals[6] = vals[7] = None
            self.rel.update(HRel._make([face] + vals))
        else:
            self.rel.update(HRel._make([None for i in range(11)]))
            self.lon = None
            self.lat = None
        if z is not None:
            self.dep = np.array(z, float)
            if od.readiness["Z"]:
                self.rel.update(VRel(npf(loc["iz"], np.int64), npf(loc["rz"]), npf(loc["dz"]), npf(loc["bz"])))
            else:
                self.rel.update(VRel._make(None for i in range(4)))
            if od.readiness["Zl"]:
                self.rel.update(
                    VlRel(npf(loc["izl_near"], np.int64), npf(loc["rzl_near"]), npf(loc["dzl_near"]), npf(loc["bzl_near"]))
                )
                self.rel.update(VlLinRel(npf(loc["izl"], np.int64), npf(loc["rzl"]), npf(loc["dzl"]), npf(loc["bzl"])))
            else:
                self.rel.update(VlRel._make(None for i in range(4)))
                self.rel.update(VlLinRel._make(None for i in range(4)))
        else:
            self.rel.update(VRel._make(None for i in range(4)))
            self.rel.update(VlRel._make(None for i in range(4)))
            self.rel.update(VlLinRel._make(None for i in range(4)))
            self.dep = None
        if t is not None:
            self.t = np.array(t, float)
            if od.readiness["time"]:
                self.rel.update(TRel(npf(loc["it"], np.int64), npf(loc["rt"]), npf(loc["dt"]), npf(loc["bt"])))
            else:
                self.rel.update(TRel._make(None for i in range(4)))
        else:
            self.rel.update(TRel._make(None for i in range(4)))
            self.t = None
        return self

    def _coordinate_arrays(self, *coords):
        """The four coordinate arguments as 1-D arrays of one common length ``self.N`` (scalars are repeated; ``None``
        stays ``None``), with the reference's complaints about shapes (``eulerian.py:160-181``)."""
        sizes = [_general_len(c) for c in coords]
        self.N = max(sizes)
        if any(size > 1 and size != self.N for size in sizes):
            raise ValueError("Shapes of input coordinates are not compatible")
        out = []
        for c in coords:
            if isinstance(c, (int, float, np.floating)):
                c = np.full(self.N, float(c))
            if c is not None and np.ndim(c) > 1:
                raise ValueError("Input need to be 1D numpy arrays")
            out.append(c)
        return out

    def from_bool_array(self, t=None, data=None, bool_array=None, num=None, random_seed=None):
        """Random points in the grid boxes where ``bool_array`` is true, in numbers proportional to the
        (weighted) cell volume (reference ``eulerian.py:219``).

        Seeding is host work on purpose: it draws from ``np.random`` in the reference's order (rx, ry,
        rzl_lin), so a given ``random_seed`` yields the reference's particles.  The points are then
        staged on the device like the result of ``from_latlon``.
        """
        import torch  # noqa: PLC0415

        if random_seed is not None:
            np.random.seed(random_seed)
        if not isinstance(data, OceData):
            raise ValueError("Input data must be OceData")
        if data.readiness["h"] != "oceanparcel":
            raise NotImplementedError("This method only support datasets that has XG, YG in it.")
        if not (data.readiness["Z"] and data.readiness["Zl"]):
            raise NotImplementedError("This method only support datasets that has all Z coords.")
        self.ocedata = od = data
        try:
            od["Vol"]
        except KeyError:
            od._add_missing_vol()
        self.tp = od.tp
        bool_array = np.asarray(bool_array)
        inds = np.where(bool_array)
        np_vol = np.array(od["Vol"] * bool_array)
        vols = np.array(np_vol[inds])
        num_each = np.round(vols * num / np.sum(vols)).astype(int)
        num = int(np.sum(num_each))
        self.N = num
        inds = tuple(np.repeat(np.array(i), num_each) for i in inds)
        if len(inds) == 3:
            iz, iy, ix = inds
            ind, face = (iy, ix), None
        elif len(inds) == 4:
            iz, face, iy, ix = inds
            ind = (face, iy, ix)
        else:
            raise ValueError("bool_array must be 3 or 4 dimensional")
        pick = lambda arr: None if arr is None else arr[ind]  # noqa: E731
        cs, sn, dx, dy, bx, by = pick(od.CS), pick(od.SN), pick(od.dX), pick(od.dY), od.XC[ind], od.YC[ind]
        rx = np.random.random(num) - 0.5
        ry = np.random.random(num) - 0.5
        rzl_lin = np.random.random(num)
        self.rel.update(HRel(face, iy, ix, rx, ry, cs, sn, dx, dy, bx, by))
        izl_lin = iz + 1
        bzl_lin = od.Zl[izl_lin]
        dzl_lin = od.dZl[izl_lin]  # sic (reference eulerian.py:287)
        self.rel.update(VlLinRel(izl_lin, rzl_lin, dzl_lin, bzl_lin))
        self.lon = bx  # temporarily: the corners are unwrapped around it
        px, py = self.get_px_py()
        w = self.get_f_node_weight()
        self.lon = np.einsum("nj,nj->n", w, px.T)
        self.lat = np.einsum("nj,nj->n", w, py.T)
        self.dep = self.bzl_lin + self.dzl_lin * self.rzl_lin
        self.rel.update(od._find_rel_v(self.dep))
        self.rel.update(od._find_rel_vl(self.dep))
        if isinstance(t, (int, float, np.floating)):
            t = np.ones(self.N, float) * t
        elif isinstance(t, np.ndarray):
            if len(t) != self.N:
                raise ValueError("Mismatch between input t and final particle number")
        if t is not None:
            self.t = copy.deepcopy(t)
            if od.readiness["time"]:
                self.rel.update(od._find_rel_t(t))
            else:
                self.rel.update(TRel._make(None for i in range(4)))
        else:
            self.rel.update(TRel._make(None for i in range(4)))
            self.t = None
        # device copy in the layout sdk_locate produces (Particle takes it from here)
        dev = od._torch_device()
        px, py = self.get_px_py()
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)  # noqa: E731
        loc = {
            "iy": up(iy, np.int32), "ix": up(ix, np.int32), "rx": up(rx, np.float64), "ry": up(ry, np.float64),
            "px": up(px.reshape(-1), np.float64), "py": up(py.reshape(-1), np.float64),
            "dx": up(dx if dx is not None else np.zeros(num), np.float32), "dy": up(dy if dy is not None else np.zeros(num), np.float32),
            "izl": up(izl_lin, np.int32), "rzl": up(rzl_lin, np.float64), "dzl": up(dzl_lin, np.float64), "bzl": up(bzl_lin, np.float64),
            "_inputs": (up(self.lon, np.float64), up(self.lat, np.float64), up(self.dep, np.float64), None if self.t is None else up(self.t, np.float64)),
        }
        if face is not None:
            loc["face"] = up(face, np.int32)
        if self.t is not None and od.readiness["time"]:
            loc["it"] = up(self.it, np.int32)
        self._located = loc
        return self

    def subset(self, which):
        """Subset of the points (reference ``eulerian.py:319``)."""
        p = object.__new__(type(self))
        for key, item in vars(self).items():
            if key.startswith("_"):
                continue
            if isinstance(item, np.ndarray):
                if len(item.shape) == 1:
                    object.__setattr__(p, key, item[which])
                    object.__setattr__(p, "N", len(item[which]))
                elif key in ["px", "py"]:
                    object.__setattr__(p, key, item[:, which])
                else:
                    object.__setattr__(p, key, item)
            elif isinstance(item, RelCoord):
                object.__setattr__(p, key, item.subset(which))
            else:
                object.__setattr__(p, key, item)
        return p

    def update_from_subset(self, sub, which):
        """Write a subset back (reference ``eulerian.py:353``): per-particle arrays and rel-coordinates of ``sub`` go into
        rows ``which`` of this object; an attribute this object does not know yet is adopted as it is (with a warning)."""
        for key, item in vars(sub).items():
            if key.startswith("_"):
                continue
            if not hasattr(self, key):
                logging.warning(f"A new attribute {key} defined after updating from subset")
                setattr(self, key, item)
            mine = getattr(self, key)
            if mine is None:
                continue
            if isinstance(item, RelCoord):
                self.rel.update_from_subset(item, which)
            elif isinstance(item, list):
                setattr(self, key, item)
            elif isinstance(item, np.ndarray):
                if item.ndim == 1:
                    mine[which] = item
                elif key in ("px", "py"):
                    mine[:, which] = item

    # ---- neighbour index sets on the host (the device kernels do this per point internally) -----------
    def _fatten_h(self, knw, ind_moves_kwarg={}):
        """Horizontal indices of every kernel node of every point, ``(faces | None, iys, ixs)`` each
        ``(N, M)`` (reference ``eulerian.py:386``): naive offsets, topology walk where they leave the grid."""
        kernel = np.asarray(knw.kernel).astype(int)
        cuvwg = ind_moves_kwarg.get("cuvwg", "C")
        tp = self.ocedata.tp
        has_face = self.face is not None
        ind = (self.face, self.iy, self.ix) if has_face else (self.iy, self.ix)
        cols = []
        for node in kernel:
            out, status = tp.fatten_node(ind, node, cuvwg)
            bad = np.where(status != 0)[0]
            if len(bad):
                tp._raise(int(status[bad[0]]), tuple(int(a[bad[0]]) for a in ind))
            cols.append(out)
        stack = lambda k: np.stack([np.asarray(c[k]) for c in cols], axis=1).astype(int)  # noqa: E731
        if has_face:
            return stack(0), stack(1), stack(2)
        return None, stack(0), stack(1)

    def _fatten_v(self, knw):
        """``iz`` or ``[iz_lin, iz_lin - 1]`` (reference ``eulerian.py:451``)."""
        if self.iz is None:
            return None
        if knw.vkernel == "nearest":
            return copy.deepcopy(self.iz)
        if knw.vkernel in ["dz", "linear"]:
            return np.vstack([self.iz_lin, self.iz_lin - 1]).T
        raise ValueError("vkernel not supported")

    def _fatten_vl(self, knw):
        """``izl`` or ``[izl_lin, izl_lin - 1]`` (reference ``eulerian.py:474``)."""
        if self.izl is None:
            return None
        if knw.vkernel == "nearest":
            return copy.deepcopy(self.izl)
        if knw.vkernel in ["dz", "linear"]:
            return np.vstack([self.izl_lin, self.izl_lin - 1]).T
        raise ValueError("vkernel not supported")

    def _fatten_t(self, knw):
        """``it`` or ``[it_lin, it_lin + 1]`` (reference ``eulerian.py:497``)."""
        if self.it is None:
            return None
        if knw.tkernel == "nearest":
            return copy.deepcopy(self.it)
        if knw.tkernel in ["dt", "linear"]:
            return np.vstack([self.it_lin, self.it_lin + 1]).T
        raise ValueError("tkernel not supported")

    def fatten(self, knw, four_d=False, required="all", ind_moves_kwarg={}):
        """Indices of all the grid points a kernel touches, in the requested dimensions (reference
        ``eulerian.py:520``).  New dimensions are appended as trailing axes: ``(N, M[, nz][, nt])``."""
        if required != "all" and isinstance(required, str):
            required = (required,)
        elif required != "all" and not isinstance(required, tuple):
            required = tuple(required)
        want = lambda name: required == "all" or name in required  # noqa: E731

        def add_axis(new, arrays):
            # broadcast `new` (N,) or (N, k) against arrays (N, ...) -> all (N, ..., k)
            new = np.asarray(new)
            new = new.reshape(len(new), -1)
            shape = arrays[0].shape
            tail = (1,) * (len(shape) - 1)
            full = shape + (new.shape[1],)
            first = np.broadcast_to(new.reshape((shape[0],) + tail + (new.shape[1],)), full).astype(int)
            rest = [np.broadcast_to(a[..., None], full).astype(int) for a in arrays]
            return [first] + rest

        if want("X") or want("Y") or want("face"):
            ffc, fiy, fix = self._fatten_h(knw, ind_moves_kwarg=ind_moves_kwarg)
            if ffc is not None:
                arrays, keys = [ffc, fiy, fix], ["face", "Y", "X"]
            else:
                arrays, keys = [fiy, fix], ["Y", "X"]
        else:
            arrays, keys = [np.zeros(self.N)], ["place_holder"]
        vertical = None
        if want("Z"):
            vertical, vkey = self._fatten_v(knw), "Z"
        elif want("Zl"):
            vertical, vkey = self._fatten_vl(knw), "Zl"
        if vertical is not None:
            arrays = add_axis(vertical, arrays)
            keys.insert(0, vkey)
        elif four_d and not (want("Z") or want("Zl")):
            arrays = [np.expand_dims(a, axis=-1) for a in arrays]
        temporal = self._fatten_t(knw) if want("time") else None
        if temporal is not None:
            arrays = add_axis(temporal, arrays)
            keys.insert(0, "time")
        elif four_d and not want("time"):
            arrays = [np.expand_dims(a, axis=-1) for a in arrays]
        table = dict(zip(keys, arrays))
        if required == "all":
            required = [k for k in keys if k != "place_holder"]
        return tuple(table[k] for k in required)

    def get_px_py(self):
        """Longitude / latitude of the 4 corners around each point (reference ``eulerian.py:592``)."""
        od = self.ocedata
        if self.face is not None:
            ind = (self.face, self.iy, self.ix)
        else:
            ind = (self.iy, self.ix)
        corners = od.tp.corner_indices(ind)
        gny, gnx = od.tp.g_shape[-2:]
        xs, ys = [], []
        for c in corners:
            c = tuple(np.asarray(a) for a in c)
            xs.append(od.XG[c])
            ys.append(od.YG[c])
        px = np.vstack(xs).astype("float64")
        py = np.vstack(ys).astype("float64")
        px = self.lon + to_180(px - self.lon)
        return px, py

    def get_f_node_weight(self):
        return weight_f_node(self.rx, self.ry)

    def interpolate(self, var_name, knw, vec_transform=True, prefetched=None, prefetch_prefix=None):
        """Interpolate / differentiate fields at the points (reference ``eulerian.py:1268``)."""
        from .interp import interpolate  # noqa: PLC0415

        return interpolate(self, var_name, knw, vec_transform, prefetched, prefetch_prefix)

    def deepcopy_position(self):
        p = object.__new__(type(self))
        for key, item in vars(self).items():
            if isinstance(item, (np.ndarray, RelCoord, list)):
                object.__setattr__(p, key, copy.deepcopy(item))
            else:
                object.__setattr__(p, key, item)
        return p
