"""Host driver of the device interpolation kernel (K3, ``csrc/interp.cu``).

Replaces the body of ``Position.interpolate`` (reference ``eulerian.py:1268``) and its stages
``_register_interpolation_input`` :648, ``_fatten_required_index_and_register`` :839,
``_transform_vector_and_register`` :887, ``_mask_value_and_register`` :959,
``_read_data_and_register`` :1064, ``_compute_weight_and_register`` :1161.  What stays on the host
is bookkeeping: which variable goes with which kernel, the output format, and per (kernel, grid
position) a small *rim table* -- the neighbour indices of the cells whose stencil leaves a face,
built once with the vectorised ``Topology`` walk and kept on the device.  Everything per point
(indices, masks, stencil choice, Lagrange weights, gathers, dot product) is one kernel launch per
variable.

Rim table slot order (per face, ``r = reach``, C-shaped faces ``ny x nx``):
    top band     iy in [ny-r, ny)            slot = (iy-(ny-r))*nx + ix
    bottom band  iy in [0, r)                slot = r*nx + iy*nx + ix
    left band    ix in [0, r),  iy in [r, ny-r)   slot = 2*r*nx + (iy-r)*r + ix
    right band   ix in [nx-r, nx), iy in [r, ny-r) slot = 2*r*nx + (ny-2r)*r + (iy-r)*r + ix-(nx-r)
or, for tiny grids (``rim_dense``), every cell with ``slot = iy*nx + ix``.
"""

from __future__ import annotations

import copy
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .kernel_weight import KnW
from .topology import OK

NODE_RAISES = -1
NODE_LAST = -2


def _pack(f, iy, ix):
    return (np.asarray(f, np.int64) << 26) | (np.asarray(iy, np.int64) << 13) | np.asarray(ix, np.int64)


# ------------------------------------------------------------------------------------------------
# rim tables
# ------------------------------------------------------------------------------------------------
def rim_cells(ny, nx, r, dense):
    """(iy, ix) of the rim cells of one face in slot order."""
    if dense:
        iy, ix = np.divmod(np.arange(ny * nx), nx)
        return iy, ix
    ys, xs = [], []
    yy, xx = np.meshgrid(np.arange(ny - r, ny), np.arange(nx), indexing="ij")
    ys.append(yy.ravel()), xs.append(xx.ravel())
    yy, xx = np.meshgrid(np.arange(0, r), np.arange(nx), indexing="ij")
    ys.append(yy.ravel()), xs.append(xx.ravel())
    yy, xx = np.meshgrid(np.arange(r, ny - r), np.arange(0, r), indexing="ij")
    ys.append(yy.ravel()), xs.append(xx.ravel())
    yy, xx = np.meshgrid(np.arange(r, ny - r), np.arange(nx - r, nx), indexing="ij")
    ys.append(yy.ravel()), xs.append(xx.ravel())
    return np.concatenate(ys), np.concatenate(xs)


def build_rim(tp, kernel, cuvwg):
    """Neighbour table of the rim cells for one horizontal stencil and grid position.

    Returns ``(reach, dense, rim int32 (nface*slots, M), code uint8 | None)``; the walk is the
    vectorised restatement of the reference's redo loop (``eulerian.py:430-445``).
    """
    kernel = np.asarray(kernel).astype(int)
    ny, nx = tp.h_shape[-2:]
    nface = tp.h_shape[0] if tp.has_face else 1
    reach = int(np.abs(kernel).max()) if kernel.size else 0
    if reach == 0:
        return 0, 0, None, None
    dense = int(min(ny, nx) < 2 * reach + 1 or ny * nx <= 1024)
    iy1, ix1 = rim_cells(ny, nx, reach, dense)
    per_face = len(iy1)
    f = np.repeat(np.arange(nface), per_face)
    iy = np.tile(iy1, nface)
    ix = np.tile(ix1, nface)
    ind = (f, iy, ix) if tp.has_face else (iy, ix)
    m = len(kernel)
    rim = np.zeros((len(iy), m), np.int64)
    faces = np.zeros((len(iy), m), np.int64)
    for j, node in enumerate(kernel):
        out, st = tp.fatten_node(ind, node, cuvwg)
        of = out[0] if tp.has_face else np.zeros_like(iy)
        oy, ox = out[-2], out[-1]
        packed = _pack(np.maximum(of, 0), np.maximum(oy, 0), np.maximum(ox, 0))
        packed = np.where((oy < 0) | (ox < 0), NODE_LAST, packed)
        packed = np.where(st != OK, NODE_RAISES, packed)
        rim[:, j] = packed
        faces[:, j] = np.where((st != OK) | (of < 0), f, of)
    code = None
    if tp.has_face and cuvwg in ("U", "V"):
        code = np.zeros((len(iy), m), np.uint8)
        different = faces != faces[:, :1]
        rows = np.where(different.any(axis=1))[0]
        if len(rows):
            sub = faces[rows]
            try:
                ufu, ufv, vfu, vfv = tp.four_matrix_for_uv(sub)
            except ValueError:
                # a node two faces away from node 0: the reference raises for such a point
                ufu = np.ones(sub.shape)
                ufv = np.zeros(sub.shape)
                vfu, vfv = ufv, ufu
                for a, row in enumerate(sub):
                    try:
                        r4 = tp.four_matrix_for_uv(row[None, :])
                        ufu[a], ufv[a], vfu[a], vfv[a] = (q[0] for q in r4)
                    except ValueError:
                        rim[rows[a], row != row[0]] = NODE_RAISES
            own, other = (ufu, ufv) if cuvwg == "U" else (vfv, vfu)
            own, other = np.round(own), np.round(other)
            from_other = np.abs(other).astype(bool)
            negative = (own + other) < 0
            code[rows] = from_other.astype(np.uint8) | (negative.astype(np.uint8) << 1)
    return reach, dense, rim.astype(np.int32), code


def _rim_on_device(od, knw, cuvwg):
    key = ("rim", hash(tuple(map(tuple, knw.kernel.astype(int).tolist()))), cuvwg)
    hit = od._interp_cache.get(key)
    if hit is None:
        import torch  # noqa: PLC0415

        reach, dense, rim, code = build_rim(od.tp, knw.kernel, cuvwg)
        dev = od._torch_device()
        hit = (
            reach,
            dense,
            None if rim is None else torch.from_numpy(rim).to(dev),
            None if code is None else torch.from_numpy(code).to(dev),
        )
        od._interp_cache[key] = hit
    return hit


# ------------------------------------------------------------------------------------------------
# ctypes mirrors of the structures in include/seaduck_b200.h
# ------------------------------------------------------------------------------------------------
MAX_NODES, MAX_LEVELS, MAX_COEF = 32, 8, 9
W_A, W_C, W_A_PLUS_C, W_A_TIMES_C = 0, 1, 2, 3


class KernelLevel(C.Structure):
    _fields_ = [("node_mask", C.c_uint32), ("reserved", C.c_uint32), ("mode", C.c_uint8 * MAX_NODES)]


class KernelDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_levels", C.c_int32),
        ("vmode", C.c_int32),
        ("tmode", C.c_int32),
        ("ignore_mask", C.c_int32),
        ("bottom_no_flux", C.c_int32),
        ("reach", C.c_int32),
        ("rim_dense", C.c_int32),
        ("n_coef", C.c_int32),
        ("reserved", C.c_int32),
        ("dx", C.c_int8 * MAX_NODES),
        ("dy", C.c_int8 * MAX_NODES),
        ("level", KernelLevel * MAX_LEVELS),
        ("coef", C.c_void_p),
        ("rim", C.c_void_p),
        ("rim_code", C.c_void_p),
    ]


class Field(C.Structure):
    _fields_ = [("data", C.c_void_p)] + [
        (n, C.c_int32) for n in "dtype has_time nt it0 has_z nz ny nx xstag ystag".split()
    ] + [("mask", C.c_void_p)] + [(n, C.c_int32) for n in "mask_nz mask_ny mask_nx reserved".split()]


class Points(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(n, C.c_void_p) for n in "face iy ix rx ry kz rz kt rt cs sn out_index".split()]


MAX_MULTI_FIELDS = 4
SEL_NONE, SEL_Z, SEL_Z_LIN, SEL_ZL, SEL_ZL_LIN = range(5)
SEL_T, SEL_T_LIN = 1, 2
JOB_SCALARS, JOB_VECTOR = 0, 1


class InterpJob(C.Structure):
    _fields_ = [(n, C.c_int32) for n in "kind n_fields zsel tsel vec_transform reserved".split()] + [
        ("fields", Field * MAX_MULTI_FIELDS), ("knw", KernelDesc * 2), ("out", C.c_void_p * MAX_MULTI_FIELDS),
    ]


class InterpIO(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(n, C.c_void_p) for n in "lon lat dep t".split()] + [
        ("sort", C.c_int32), ("max_chunks", C.c_int32), ("out_flags", C.c_void_p),
    ]


def axis_polynomial(nodes, target, order):
    """Coefficients (ascending powers of r) of the Lagrange weight of ``target`` among the 1-D
    ``nodes`` for the ``order``-th derivative, as the reference defines it (kernel_weight.py:136-230:
    sum over the subsets of the other nodes of prod (r - o), over prod (target - o); at the maximum
    order the numerator is (order - 1)!).  Small-integer arithmetic: exact in float64."""
    from itertools import combinations  # noqa: PLC0415
    from math import factorial  # noqa: PLC0415

    others = [o for o in nodes if o != target]
    denom = 1.0
    for o in others:
        denom *= target - o
    if order > len(nodes) - 1:
        raise ValueError("Kernel is too small for this derivative")
    if order == len(nodes) - 1:
        return np.array([float(factorial(max(order - 1, 0))) / denom])
    num = np.zeros(1)
    for sub in combinations(others, len(others) - order):
        term = np.ones(1)
        for o in sub:
            term = np.convolve(term, np.array([-float(o), 1.0]))
        num = np.pad(num, (0, max(0, len(term) - len(num)))) + np.pad(term, (0, max(0, len(num) - len(term))))
    return num / denom


def level_polynomials(stencil, nodes_xy):
    """Per node of one inheritance level: (mode, A coefficients, C coefficients) so that the weight is
    A(rx), C(ry), A(rx) + C(ry) or A(rx) * C(ry) (``_Stencil.__call__`` in kernel_weight.py)."""
    from .kernel_weight import KIND_CROSS_INTERP, KIND_LINE_X, KIND_LINE_Y  # noqa: PLC0415

    xs, ys = stencil.xs, stencil.ys
    zero, one = np.zeros(1), np.ones(1)
    out = []
    for x, y in nodes_xy:
        x, y = int(x), int(y)
        if stencil.kind == KIND_CROSS_INTERP:
            if x != 0:
                out.append((W_A, axis_polynomial(xs, x, 0), zero))
            elif y != 0:
                out.append((W_C, zero, axis_polynomial(ys, y, 0)))
            else:
                c = axis_polynomial(ys, 0, 0).copy()
                c[0] -= 1.0
                out.append((W_A_PLUS_C, axis_polynomial(xs, 0, 0), c))
        elif stencil.kind == KIND_LINE_X:
            out.append((W_A, axis_polynomial(xs, x, stencil.xorder) if y == 0 else zero, zero))
        elif stencil.kind == KIND_LINE_Y:
            out.append((W_C, zero, axis_polynomial(ys, y, stencil.yorder) if x == 0 else zero))
        else:
            out.append((W_A_TIMES_C, axis_polynomial(xs, x, stencil.xorder), axis_polynomial(ys, y, stencil.yorder)))
    del one
    return out


def kernel_desc(od, knw, cuvwg, ignore_mask, bottom_no_flux):
    """``sdk_kernel_desc`` of a KnW (plus the tensors that must outlive the call)."""
    import torch  # noqa: PLC0415

    d = knw.describe()
    kern = d["kernel"]
    if len(kern) > MAX_NODES or len(d["levels"]) > MAX_LEVELS:
        raise NotImplementedError("kernel larger than the device limits (32 nodes, 8 levels)")
    k = KernelDesc()
    k.n_nodes = len(kern)
    k.n_levels = len(d["levels"])
    k.vmode, k.tmode = d["vmode"], d["tmode"]
    k.ignore_mask = int(ignore_mask)
    k.bottom_no_flux = int(bottom_no_flux)
    for j, (x, y) in enumerate(kern):
        k.dx[j], k.dy[j] = int(x), int(y)
    key = ("coef", hash(knw))
    hit = od._interp_cache.get(key)
    if hit is None:
        polys = []
        for lev, st in zip(d["levels"], knw.hfuncs):
            polys.append(level_polynomials(st, [kern[i] for i in lev["nodes"]]))
        n_coef = max(max(len(a), len(c)) for lp in polys for _, a, c in lp)
        if n_coef > MAX_COEF:
            raise NotImplementedError("more than 9 distinct node positions along one axis")
        table = np.zeros((k.n_levels, k.n_nodes, 2, n_coef))
        modes = np.zeros((k.n_levels, k.n_nodes), np.uint8)
        for li, (lev, lp) in enumerate(zip(d["levels"], polys)):
            for node, (mode, a, c) in zip(lev["nodes"], lp):
                modes[li, node] = mode
                table[li, node, 0, : len(a)] = a
                table[li, node, 1, : len(c)] = c
        hit = (n_coef, modes, torch.from_numpy(table).to(od._torch_device()))
        od._interp_cache[key] = hit
    n_coef, modes, table = hit
    k.n_coef = n_coef
    k.coef = table.data_ptr()
    for li, lev in enumerate(d["levels"]):
        L = k.level[li]
        L.node_mask = sum(1 << int(i) for i in lev["nodes"])
        for node in range(k.n_nodes):
            L.mode[node] = int(modes[li, node])
    reach, dense, rim, code = _rim_on_device(od, knw, cuvwg)
    k.reach, k.rim_dense = reach, dense
    k.rim = None if rim is None else rim.data_ptr()
    k.rim_code = None if code is None else code.data_ptr()
    return k, (rim, code, table)


# ------------------------------------------------------------------------------------------------
# masks and fields on the device
# ------------------------------------------------------------------------------------------------
def _device_masks(od):
    """uint8 maskC / maskU / maskV / maskWvel on the device, padded like the reference's
    ``get_masked`` (``get_masks.py:169``: wall masks are zero-padded to the staggered sizes)."""
    hit = od._interp_cache.get("masks")
    if hit is not None:
        return hit
    import torch  # noqa: PLC0415

    ds = od._ds
    keys = list(ds.keys())
    if "maskC" not in keys:
        od._interp_cache["masks"] = {}
        return {}
    mdims = ds["maskC"].dims
    if "Z" not in mdims:
        raise NotImplementedError(
            "2D land mask is not yet supported,"
            "you could potentially get around by adding a psuedo dimension"
            "or you could set knw.ignore_mask = True"
        )
    dev = od._torch_device()
    maskc = np.ascontiguousarray(np.asarray(ds["maskC"]) != 0).astype(np.uint8)
    tc = torch.from_numpy(maskc).to(dev)
    nz = maskc.shape[0]
    small = {"C": tc}
    need = [c for c in ("U", "V", "Wvel") if "mask" + c not in keys]
    if need:
        tu, tv, tw = (torch.empty_like(tc) for _ in range(3))
        _lib.check(
            _bind().sdk_build_masks(od.grid_handle(), tc.data_ptr(), nz, tu.data_ptr(), tv.data_ptr(), tw.data_ptr(), _lib.current_stream()),
            "sdk_build_masks",
        )
        small.update({"U": tu, "V": tv, "Wvel": tw})
    rename = {
        "U": lambda x: x if x != "X" else "Xp1",
        "V": lambda x: x if x != "Y" else "Xp1",  # sic (reference get_masks.py:205)
        "Wvel": lambda x: x if x != "Z" else "Zl",
    }
    out = {"C": tc}
    for c in ("U", "V", "Wvel"):
        if "mask" + c in keys:
            out[c] = torch.from_numpy(np.ascontiguousarray(np.asarray(ds["mask" + c]) != 0).astype(np.uint8)).to(dev)
            continue
        dims = tuple(map(rename[c], mdims))
        sizes = tuple(len(ds[dim]) for dim in dims)
        sm = small[c]
        if sizes == tuple(sm.shape):
            out[c] = sm
        else:
            big = torch.zeros(sizes, dtype=torch.uint8, device=dev)
            sl = tuple(slice(0, min(a, b)) for a, b in zip(sm.shape, sizes))
            big[sl] = sm[sl]
            out[c] = big
    od._interp_cache["masks"] = out
    return out


def _mask_kind(dims):
    if "Zl" in dims:
        return "Wvel"
    if "Z" in dims:
        if "Xp1" in dims and "Yp1" in dims:
            raise NotImplementedError(
                "The masking of corner points are open to "
                "interpretations thus not implemented, "
                "let knw.ignore_mask =True to go around"
            )
        if "Xp1" in dims:
            return "U"
        if "Yp1" in dims:
            return "V"
    return "C"


def _field(od, name, dims, prefetched, prefix, mask):
    """``sdk_field`` for a variable (or a prefetched array standing in for it)."""
    import torch  # noqa: PLC0415

    f = Field()
    it0 = 0
    if prefetched is not None:
        if prefix is None:
            raise ValueError("please pass value of the prefix of prefetched dataset, even if the prefix is zero")
        if any(int(p) != 0 for p in prefix[1:]) or ("time" not in dims and len(prefix) and int(prefix[0]) != 0):
            raise NotImplementedError("prefetch_prefix other than a leading time offset")
        it0 = int(prefix[0]) if len(prefix) else 0
        arr = np.asarray(prefetched)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(od._torch_device())
    else:
        t = od.device_field(name)
    if t.dim() != len(dims):
        raise ValueError(f"{name}: array with {t.dim()} dimensions for dims {dims}")
    f.data = t.data_ptr()
    f.dtype = 0 if t.dtype == torch.float64 else 1
    f.has_time = int("time" in dims)
    f.nt = t.shape[dims.index("time")] if f.has_time else 1
    f.it0 = it0
    zname = "Z" if "Z" in dims else ("Zl" if "Zl" in dims else None)
    f.has_z = int(zname is not None)
    f.nz = t.shape[dims.index(zname)] if zname else 1
    f.ny, f.nx = t.shape[-2], t.shape[-1]
    f.xstag = int("Xp1" in dims)
    f.ystag = int("Yp1" in dims)
    keep = [t]
    if mask is not None:
        f.mask = mask.data_ptr()
        f.mask_nz, f.mask_ny, f.mask_nx = mask.shape[0], mask.shape[-2], mask.shape[-1]
        keep.append(mask)
    return f, keep


# ------------------------------------------------------------------------------------------------
# points on the device
# ------------------------------------------------------------------------------------------------
_PARTICLE_DEV = {"face": "face", "iy": "iy", "ix": "ix", "rx": "rx", "ry": "ry", "izl_lin": "izl", "rzl_lin": "rzl", "it": "it"}
_LAZY = {
    "iz_lin": ("_find_rel_v_lin", "dep"), "rz_lin": ("_find_rel_v_lin", "dep"),
    "it_lin": ("_find_rel_t_lin", "t"), "rt_lin": ("_find_rel_t_lin", "t"),
    "izl_lin": ("_find_rel_vl_lin", "dep"), "rzl_lin": ("_find_rel_vl_lin", "dep"),
}
_LOCATED_NAME = {"izl": "izl_near", "rzl": "rzl_near", "izl_lin": "izl", "rzl_lin": "rzl"}
_FLAG = {"iz_lin": 32, "rz_lin": 32, "it_lin": 64, "rt_lin": 64}


class _DevicePoints:
    """Device tensors of the attributes of a Position (borrowed from ``from_latlon`` when the
    position has not been edited since, uploaded from the numpy mirrors otherwise)."""

    SORT_THRESHOLD = 200_000

    def __init__(self, pos):
        self.pos = pos
        self.od = pos.ocedata
        self.cache = {}
        self.perm = None
        self._sorted = None
        self._setup_sort()

    def _setup_sort(self):
        """Scattered points make every 8-byte gather cost a 32-byte sector.  For a large, located
        ``Position`` the kernels therefore run over a cell-sorted copy of the point arrays (threads
        of a warp then read neighbouring cells) and the results are scattered back to the caller's
        order.  Permutation and sorted copies are cached with the located state, so several
        variables (and repeated calls) share them."""
        import torch  # noqa: PLC0415

        d = self.pos.__dict__
        loc = d.get("_located")
        if "_st" in d or loc is None or self.pos.N < self.SORT_THRESHOLD or "iy" not in loc:
            return
        if "_perm" not in loc:
            tp = self.od.tp
            ny, nx = tp.h_shape[-2:]
            key = loc["iy"].long() * nx + loc["ix"].long()
            if "face" in loc:
                key += loc["face"].long() * (ny * nx)
            lev = loc.get("iz", loc.get("izl"))
            if lev is not None:
                key += lev.long() * (ny * nx * (tp.h_shape[0] if tp.has_face else 1))
            loc["_perm"] = torch.argsort(key).to(torch.int32)
            loc["_sorted"] = {}
        self.perm = loc["_perm"]
        self._sorted = loc["_sorted"]


    def get(self, name, integer=False):
        t = self._get(name, integer)
        if t is None or self.perm is None:
            return t
        if name not in self._sorted:
            self._sorted[name] = t[self.perm.long()]
        return self._sorted[name]

    def _get(self, name, integer=False):
        import torch  # noqa: PLC0415

        if name in self.cache:
            return self.cache[name]
        pos = self.pos
        d = pos.__dict__
        t = None
        if "_st" in d and name in _PARTICLE_DEV:
            t = d["_st"].get(_PARTICLE_DEV[name])
            if t is not None and d.get("_inv") is not None:
                t = t[d["_inv"]]  # the state is kept in cell order; everything else here is in the caller's
        loc = d.get("_located")
        if t is None and loc is not None:
            key = _LOCATED_NAME.get(name, name)
            if key in loc:
                if loc.get("_flags", 0) & _FLAG.get(name, 0):
                    raise ValueError("Value out of bound.")
                t = loc[key]
        if t is None:
            val = None
            if "_st" in d:
                # a Particle: attribute access knows which host mirrors are stale / derived
                try:
                    val = getattr(pos, name)
                except AttributeError:
                    val = None
            elif name in pos.rel and pos.rel[name] is not None:
                val = pos.rel[name]
            elif name in _LAZY and name not in pos.rel:
                func, src = _LAZY[name]
                coord = getattr(pos, src, None)
                if coord is not None:
                    pos.rel.update(getattr(self.od, func)(coord))
                    val = pos.rel.get(name)
            else:
                try:
                    val = getattr(pos, name)
                except AttributeError:
                    val = None
            if val is None:
                self.cache[name] = None
                return None
            dtype = np.int32 if integer else (np.float32 if name in ("cs", "sn") else np.float64)
            t = torch.from_numpy(np.ascontiguousarray(val, dtype)).to(self.od._torch_device())
        if t.shape[0] != pos.N:
            raise ValueError(f"{name}: {t.shape[0]} entries for {pos.N} points")
        self.cache[name] = t
        return t


def _vertical_time(pts, dims, knw, p):
    """Fill kz/rz/kt/rt of ``p`` for a variable with ``dims`` (reference ``_fatten_v/_vl/_t``
    ``eulerian.py:451-518`` and the rz/rt choice of ``_compute_weight_and_register`` :1196-1224)."""
    keep = []
    if "Z" in dims or "Zl" in dims:
        stem = "z" if "Z" in dims else "zl"
        suffix = "" if knw.vkernel == "nearest" else "_lin"
        if knw.vkernel not in ("nearest", "linear", "dz"):
            raise ValueError("vkernel not supported")
        kz = pts.get(f"i{stem}{suffix}", integer=True)
        if kz is None:
            raise ValueError("the variable has a vertical dimension but the points have no depth")
        rz = pts.get(f"r{stem}{suffix}")
        p.kz, p.rz = kz.data_ptr(), (None if rz is None else rz.data_ptr())
        keep += [kz, rz]
    if "time" in dims:
        if knw.tkernel not in ("nearest", "linear", "dt"):
            raise ValueError("tkernel not supported")
        suffix = "" if knw.tkernel == "nearest" else "_lin"
        kt = pts.get(f"it{suffix}", integer=True)
        if kt is None:
            raise ValueError("the variable has a time dimension but the points have no time")
        rt = pts.get(f"rt{suffix}")
        p.kt, p.rt = kt.data_ptr(), (None if rt is None else rt.data_ptr())
        keep += [kt, rt]
    return keep


def _points_struct(pts):
    p = Points()
    p.n = pts.pos.N
    keep = []
    for name, integer in (("face", True), ("iy", True), ("ix", True), ("rx", False), ("ry", False)):
        t = pts.get(name, integer)
        keep.append(t)
        setattr(p, name, None if t is None else t.data_ptr())
    if pts.perm is not None:
        p.out_index = pts.perm.data_ptr()  # results go straight back to the caller's order
        keep.append(pts.perm)
    return p, keep


# ------------------------------------------------------------------------------------------------
# the two launches
# ------------------------------------------------------------------------------------------------
def _bind():
    L = _lib.lib()
    if not getattr(L, "_interp_bound", False):
        vp = C.c_void_p
        L.sdk_interp_scalar.argtypes = [vp, C.POINTER(Points), C.POINTER(Field), C.POINTER(KernelDesc), vp, vp]
        L.sdk_interp_vector.argtypes = [
            vp, C.POINTER(Points), C.POINTER(Field), C.POINTER(Field), C.POINTER(KernelDesc), C.POINTER(KernelDesc), C.c_int, vp, vp, vp,
        ]
        L.sdk_interp_scalar_multi.argtypes = [vp, C.POINTER(Points), C.POINTER(Field), C.c_int32, C.POINTER(KernelDesc), vp, vp]
        L.sdk_build_masks.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp]
        L.sdk_interp_host.argtypes = [vp, C.POINTER(InterpIO), vp, C.c_int32, C.POINTER(InterpJob), C.c_int32, vp, vp]
        L._interp_bound = True
    return L


def _mask_for(od, dims, ignore_mask, kind=None):
    if ignore_mask or "X" not in "".join(map(str, dims)) or "Y" not in "".join(map(str, dims)):
        return None
    masks = _device_masks(od)
    if not masks:
        warnings.warn("No maskC in dataset. Perhaps set ds['maskC'] = ds.HFacC")
        return None
    return masks[kind or _mask_kind(dims)]


def interp_scalar_device(pts, name, dims, knw, prefetched=None, prefix=None):
    """One scalar variable -> device tensor of N float64."""
    import torch  # noqa: PLC0415

    od = pts.od
    L = _bind()
    cuvwg = "G" if ("Xp1" in dims or "Yp1" in dims) else "C"
    mask = _mask_for(od, dims, knw.ignore_mask)
    # NB the reference passes bottom_scheme "no_flux" (sic) for scalars, which get_weight does not
    # recognise ("no flux"): no ghost point below the bottom for scalars (eulerian.py:1193)
    k, keep_k = kernel_desc(od, knw, cuvwg, knw.ignore_mask or mask is None, bottom_no_flux=False)
    f, keep_f = _field(od, name, dims, prefetched, prefix, mask)
    p, keep_p = _points_struct(pts)
    keep_v = _vertical_time(pts, dims, knw, p)
    out = torch.empty(pts.pos.N, dtype=torch.float64, device=od._torch_device())
    _lib.check(
        L.sdk_interp_scalar(od.grid_handle(), C.byref(p), C.byref(f), C.byref(k), out.data_ptr(), _lib.current_stream()),
        "sdk_interp_scalar",
    )
    del keep_k, keep_f, keep_p, keep_v
    return out


def interp_scalars_device(pts, names, dims, knw):
    """Several scalar variables that share dims, dtype layout and kernel -> list of device tensors, one launch
    (``sdk_interp_scalar_multi``): indices, wet flags, stencil choice and weights are worked out once per point."""
    import torch  # noqa: PLC0415

    od = pts.od
    L = _bind()
    cuvwg = "G" if ("Xp1" in dims or "Yp1" in dims) else "C"
    mask = _mask_for(od, dims, knw.ignore_mask)
    k, keep_k = kernel_desc(od, knw, cuvwg, knw.ignore_mask or mask is None, bottom_no_flux=False)
    fields = (Field * len(names))()
    keep_f = []
    for j, name in enumerate(names):
        f, keep = _field(od, name, dims, None, None, mask)
        fields[j] = f
        keep_f.append(keep)
    p, keep_p = _points_struct(pts)
    keep_v = _vertical_time(pts, dims, knw, p)
    outs = [torch.empty(pts.pos.N, dtype=torch.float64, device=od._torch_device()) for _ in names]
    out_ptrs = (C.c_void_p * len(names))(*[o.data_ptr() for o in outs])
    _lib.check(
        L.sdk_interp_scalar_multi(od.grid_handle(), C.byref(p), fields, len(names), C.byref(k), out_ptrs, _lib.current_stream()),
        "sdk_interp_scalar_multi",
    )
    del keep_k, keep_f, keep_p, keep_v
    return outs


def interp_vector_device(pts, names, dims, knws, vec_transform, prefetched=None, prefix=None):
    """A horizontal vector on a grid with faces -> two device tensors."""
    import torch  # noqa: PLC0415

    od = pts.od
    L = _bind()
    uknw, vknw = knws
    udims, vdims = dims
    if uknw.vkernel != vknw.vkernel or uknw.tkernel != vknw.tkernel:
        raise NotImplementedError("the two kernels of a vector must share vkernel and tkernel")
    ignore = uknw.ignore_mask or vknw.ignore_mask
    umask = _mask_for(od, udims, ignore, "U")
    vmask = _mask_for(od, vdims, ignore, "V")
    ignore = ignore or umask is None or vmask is None
    # vectors use get_weight's default bottom_scheme="no flux" (eulerian.py:1259-1264)
    ku, keep_ku = kernel_desc(od, uknw, "U", ignore, bottom_no_flux=uknw.vkernel == "linear")
    kv, keep_kv = kernel_desc(od, vknw, "V", ignore, bottom_no_flux=vknw.vkernel == "linear")
    upre, vpre = prefetched if prefetched is not None else (None, None)
    fu, keep_fu = _field(od, names[0], udims, upre, prefix, umask)
    fv, keep_fv = _field(od, names[1], vdims, vpre, prefix, vmask)
    p, keep_p = _points_struct(pts)
    keep_v = _vertical_time(pts, udims, uknw, p)
    if vec_transform:
        cs, sn = pts.get("cs"), pts.get("sn")
        if cs is None or sn is None:
            raise ValueError("vec_transform needs CS and SN in the dataset")
        p.cs, p.sn = cs.data_ptr(), sn.data_ptr()
        keep_v += [cs, sn]
    dev = od._torch_device()
    out_u = torch.empty(pts.pos.N, dtype=torch.float64, device=dev)
    out_v = torch.empty(pts.pos.N, dtype=torch.float64, device=dev)
    _lib.check(
        L.sdk_interp_vector(
            od.grid_handle(), C.byref(p), C.byref(fu), C.byref(fv), C.byref(ku), C.byref(kv), int(bool(vec_transform)),
            out_u.data_ptr(), out_v.data_ptr(), _lib.current_stream(),
        ),
        "sdk_interp_vector",
    )
    del keep_ku, keep_kv, keep_fu, keep_fv, keep_p, keep_v
    return out_u, out_v


# ------------------------------------------------------------------------------------------------
# Position.interpolate
# ------------------------------------------------------------------------------------------------
def _broadcast_inputs(pos, var_name, knw, prefetched, prefetch_prefix):
    """Input normalisation of the reference (``_register_interpolation_input``, eulerian.py:648-800)."""
    single = False
    if isinstance(var_name, (str, tuple)):
        var_name = [var_name]
        single = True
    elif not isinstance(var_name, list):
        raise ValueError("var_name needs to be string, tuple, or a list of the above.")
    num_var = len(var_name)
    if isinstance(knw, KnW):
        knw = [knw for _ in range(num_var)]
    if isinstance(knw, tuple):
        if len(knw) != 2:
            raise ValueError(
                "When knw is a tuple, we assume it to be kernels for a horizontal vector,"
                "thus, it has to have 2 elements"
            )
        knw = [knw for _ in range(num_var)]
    elif isinstance(knw, list):
        if len(knw) != num_var:
            raise ValueError("Mismatch between the number of kernels and variables")
    elif isinstance(knw, dict):
        temp = []
        for var in var_name:
            a_knw = knw.get(var)
            if a_knw is None or not isinstance(a_knw, (tuple, KnW)):
                raise ValueError(f"Variable {var} does not have a proper corresponding kernel")
            temp.append(a_knw)
        knw = temp
    else:
        raise ValueError("knw needs to be a KnW object, or list/dictionaries containing that ")

    if isinstance(prefetched, (np.ndarray, tuple)) or prefetched is None:
        prefetched = [prefetched for _ in range(num_var)]
    elif isinstance(prefetched, list):
        if len(prefetched) != num_var:
            raise ValueError("Mismatch between the number of prefetched arrays and variables")
    elif isinstance(prefetched, dict):
        prefetched = [prefetched.get(var) for var in var_name]
    else:
        raise ValueError(
            "prefetched needs to be a numpy array/tuple pair of numpy array, or list/dictionaries containing that"
        )
    if isinstance(prefetch_prefix, tuple):
        prefetch_prefix = [prefetch_prefix for _ in range(num_var)]
    elif prefetch_prefix is None:
        prefetch_prefix = [None for _ in range(num_var)]
    elif isinstance(prefetch_prefix, list):
        if len(prefetch_prefix) != num_var:
            raise ValueError("Mismatch between the number of prefetched arrays prefix prefetch_prefix and variables")
    elif isinstance(prefetch_prefix, dict):
        prefetch_prefix = [prefetch_prefix.get(var) for var in var_name]
    else:
        raise ValueError("prefetched prefix prefetch_prefix needs to be a tuple, or list/dictionaries containing that ")
    return single, var_name, knw, prefetched, prefetch_prefix


def _plan(od, has_face, var_name, knw, prefetched, prefetch_prefix):
    """What has to be launched for a request, in order: a list of ``(keys, kind, names, kernels, prefetched, prefix)``
    with kind ``"m"`` (several scalars that share the kernel and the grid position: one launch, up to four at a time),
    ``"s"`` (one scalar) or ``"v"`` (a horizontal vector on a grid with faces).  ``keys`` are the ``(variable, kernel)``
    pairs the launch answers; a pair asked for twice is computed once (the reference hashes likewise)."""
    jobs = []
    for i, var in enumerate(var_name):
        if isinstance(var, str):
            if not isinstance(knw[i], KnW):
                raise ValueError("a scalar variable needs a single KnW")
            jobs.append(((var, knw[i]), "s", var, knw[i], prefetched[i], prefetch_prefix[i]))
        elif isinstance(var, tuple):
            if not isinstance(knw[i], tuple) or len(knw[i]) != 2:
                raise ValueError(
                    "When knw is a tuple, we assume it to be kernels for a horizontal vector,"
                    "thus, it has to have 2 elements"
                )
            if not has_face:
                for j in range(len(var)):
                    pre = prefetched[i][j] if prefetched[i] is not None else None
                    jobs.append(((var[j], knw[i][j]), "s", var[j], knw[i][j], pre, prefetch_prefix[i]))
            else:
                jobs.append(((var, knw[i]), "v", var, knw[i], prefetched[i], prefetch_prefix[i]))
        elif var is None:
            pass
        else:
            raise ValueError("var_name needs to be string, tuple, or a list of the above.")
    launches, done = [], set()
    groups = {}
    for key, kind, names, kernels, pre, prefix in jobs:
        if kind == "s" and pre is None and prefix is None:
            gkey = (kernels, tuple(od[names].dims), str(od[names].dtype))
            if key not in groups.setdefault(gkey, []):
                groups[gkey].append(key)
    for (kernels, dims, _), keys in groups.items():
        for at in range(0, len(keys), MAX_MULTI_FIELDS):
            chunk = keys[at : at + MAX_MULTI_FIELDS]
            if len(chunk) > 1:
                launches.append((chunk, "m", [k[0] for k in chunk], kernels, None, None))
                done.update(chunk)
    for key, kind, names, kernels, pre, prefix in jobs:
        if key not in done:
            launches.append(([key], kind, names, kernels, pre, prefix))
            done.add(key)
    return launches


def _assemble(final, var_name, knw, has_face, single):
    """Results in the nesting of the request."""
    output = []
    for var, kk in zip(var_name, knw):
        if var is None:
            output.append(None)
        elif isinstance(var, tuple) and not has_face:
            output.append(tuple(final.get((var[i], kk[i])) for i in range(len(var))))
        else:
            output.append(final.get((var, kk)))
    return output[0] if single else output


def interpolate_device(pos, var_name, knw, vec_transform=True, prefetched=None, prefetch_prefix=None):
    """Like :func:`interpolate` but the results stay on the device (torch tensors)."""
    single, var_name, knw, prefetched, prefetch_prefix = _broadcast_inputs(pos, var_name, knw, prefetched, prefetch_prefix)
    od = pos.ocedata
    has_face = pos.rel.has("face") if "_st" not in pos.__dict__ else pos.__dict__["_st"].get("face") is not None
    pts = _DevicePoints(pos)
    final = {}
    for keys, kind, names, kernels, pre, prefix in _plan(od, has_face, var_name, knw, prefetched, prefetch_prefix):
        if kind == "m":
            for key, out in zip(keys, interp_scalars_device(pts, names, od[names[0]].dims, kernels)):
                final[key] = out
        elif kind == "s":
            final[keys[0]] = interp_scalar_device(pts, names, od[names].dims, kernels, pre, prefix)
        else:
            dims = (od[names[0]].dims, od[names[1]].dims)
            final[keys[0]] = interp_vector_device(pts, names, dims, kernels, vec_transform, pre, prefix)
    return _assemble(final, var_name, knw, has_face, single)


# ------------------------------------------------------------------------------------------------
# host arrays in, host arrays out: one call, chunks pipelined inside the library
# ------------------------------------------------------------------------------------------------
HOST_PATH_THRESHOLD = 200_000


def pinned(shape, dtype=np.float64):
    """A numpy array in page-locked host memory (from torch's caching host allocator, so repeated requests of a
    size reuse the same pages).  Coordinates handed to ``OceInterp`` / ``interpolate_host`` in such arrays are
    uploaded by DMA at PCIe speed; ordinary numpy arrays work too and go through the driver's staging copy."""
    import torch  # noqa: PLC0415

    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
           np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()


def _vertical_time_selectors(od, dims, knw, have_dep, have_time):
    """Which index family of ``sdk_locate`` a variable and its kernel read (same choices as :func:`_vertical_time`)."""
    zsel = tsel = SEL_NONE
    if "Z" in dims or "Zl" in dims:
        if knw.vkernel not in ("nearest", "linear", "dz"):
            raise ValueError("vkernel not supported")
        if not have_dep or not od.readiness["Z" if "Z" in dims else "Zl"]:
            raise ValueError("the variable has a vertical dimension but the points have no depth")
        near = knw.vkernel == "nearest"
        zsel = (SEL_Z if near else SEL_Z_LIN) if "Z" in dims else (SEL_ZL if near else SEL_ZL_LIN)
    if "time" in dims:
        if knw.tkernel not in ("nearest", "linear", "dt"):
            raise ValueError("tkernel not supported")
        if not have_time:
            raise ValueError("the variable has a time dimension but the points have no time")
        tsel = SEL_T if knw.tkernel == "nearest" else SEL_T_LIN
    return zsel, tsel


def interpolate_host(od, x, y, z, t, var_name, knw, vec_transform=True, sort=True, max_chunks=0):
    """``Position().from_latlon(x, y, z, t, data=od).interpolate(var_name, knw)`` for host arrays as ONE library call
    (``sdk_interp_host``): the points are cut into chunks whose upload, locating, interpolation and read-back overlap
    on three streams, and nothing but the requested values comes back.  Same values, same nesting of the result, same
    exceptions as the two-step route (``tests/test_gpu_interp.py``); the results are numpy arrays in pinned memory."""
    import torch  # noqa: PLC0415

    single, var_name, knw, prefetched, prefix = _broadcast_inputs(None, var_name, knw, None, None)
    n = len(x)
    have_dep = z is not None and (od.readiness["Z"] or od.readiness["Zl"])
    have_time = t is not None and bool(od.readiness["time"])
    has_face = bool(od.tp.has_face)
    launches = _plan(od, has_face, var_name, knw, prefetched, prefix)
    L = _bind()
    jobs = (InterpJob * max(len(launches), 1))()
    keep, final = [], {}
    for job, (keys, kind, names, kernels, _, _) in zip(jobs, launches):
        if kind == "v":
            uk, vk = kernels
            udims, vdims = od[names[0]].dims, od[names[1]].dims
            if uk.vkernel != vk.vkernel or uk.tkernel != vk.tkernel:
                raise NotImplementedError("the two kernels of a vector must share vkernel and tkernel")
            ignore = uk.ignore_mask or vk.ignore_mask
            umask, vmask = _mask_for(od, udims, ignore, "U"), _mask_for(od, vdims, ignore, "V")
            ignore = ignore or umask is None or vmask is None
            ku, keep_ku = kernel_desc(od, uk, "U", ignore, bottom_no_flux=uk.vkernel == "linear")
            kv, keep_kv = kernel_desc(od, vk, "V", ignore, bottom_no_flux=vk.vkernel == "linear")
            fu, keep_fu = _field(od, names[0], udims, None, None, umask)
            fv, keep_fv = _field(od, names[1], vdims, None, None, vmask)
            if vec_transform and (od.CS is None or od.SN is None):
                raise ValueError("vec_transform needs CS and SN in the dataset")
            job.kind, job.n_fields, job.vec_transform = JOB_VECTOR, 2, int(bool(vec_transform))
            job.fields[0], job.fields[1], job.knw[0], job.knw[1] = fu, fv, ku, kv
            job.zsel, job.tsel = _vertical_time_selectors(od, udims, uk, have_dep, have_time)
            keep += [keep_ku, keep_kv, keep_fu, keep_fv]
            outs = [pinned(n), pinned(n)]
            final[keys[0]] = (outs[0], outs[1])
        else:
            names = [names] if kind == "s" else names
            dims = od[names[0]].dims
            cuvwg = "G" if ("Xp1" in dims or "Yp1" in dims) else "C"
            mask = _mask_for(od, dims, kernels.ignore_mask)
            k, keep_k = kernel_desc(od, kernels, cuvwg, kernels.ignore_mask or mask is None, bottom_no_flux=False)
            job.kind, job.n_fields = JOB_SCALARS, len(names)
            job.knw[0] = k
            keep.append(keep_k)
            for j, name in enumerate(names):
                f, keep_f = _field(od, name, dims, None, None, mask)
                job.fields[j] = f
                keep.append(keep_f)
            job.zsel, job.tsel = _vertical_time_selectors(od, dims, kernels, have_dep, have_time)
            outs = [pinned(n) for _ in names]
            for key, out in zip(keys, outs):
                final[key] = out
        for j, out in enumerate(outs):
            job.out[j] = out.ctypes.data
    coords = [np.ascontiguousarray(c, np.float64) if c is not None else None for c in (x, y, z if have_dep else None, t if have_time else None)]
    io = InterpIO()
    io.n = n
    io.lon, io.lat, io.dep, io.t = (None if c is None else c.ctypes.data for c in coords)
    io.sort, io.max_chunks = int(bool(sort)), int(max_chunks)
    flags = np.zeros(1, np.uint32)
    io.out_flags = flags.ctypes.data
    n_out = sum(job.n_fields for job in jobs[: len(launches)])
    need = int(L.sdk_interp_host_scratch_bytes(n, n_out))
    scratch = od._interp_cache.get("host_scratch")
    if scratch is None or scratch.numel() < need:
        od._interp_cache.pop("host_scratch", None)
        del scratch
        scratch = od._interp_cache["host_scratch"] = torch.empty(need, dtype=torch.uint8, device=od._torch_device())
    ts_dev, n_t = None, 0
    if have_time:
        ts_dev, n_t = od.to_device("ts", od.ts, np.float64), len(od.ts)
    _lib.check(
        L.sdk_interp_host(od.grid_handle(), C.byref(io), _lib.ptr(ts_dev), n_t, jobs, len(launches), scratch.data_ptr(), _lib.current_stream()),
        "sdk_interp_host",
    )
    bits = int(flags[0])
    used_z = {job.zsel for job in jobs[: len(launches)]}
    used_t = {job.tsel for job in jobs[: len(launches)]}
    if bits & 16 or (bits & 32 and SEL_Z_LIN in used_z) or (bits & 64 and SEL_T_LIN in used_t):
        raise ValueError("Value out of bound.")  # reference find_ind, utils.py:301
    del keep
    return _assemble(final, var_name, knw, has_face, single)


def _to_host(x):
    if x is None:
        return None
    if isinstance(x, tuple):
        return tuple(_to_host(v) for v in x)
    if isinstance(x, list):
        return [_to_host(v) for v in x]
    # through pinned memory (torch's caching host allocator): DMA at PCIe speed instead of a pageable copy
    import torch  # noqa: PLC0415

    out = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    out.copy_(x)
    return out.numpy()


def interpolate(pos, var_name, knw, vec_transform=True, prefetched=None, prefetch_prefix=None):
    """``Position.interpolate`` (reference ``eulerian.py:1268``): numpy out, same nesting as the input."""
    return _to_host(interpolate_device(pos, copy.copy(var_name), knw, vec_transform, prefetched, prefetch_prefix))
