"""The regular-grid stencils of the Eulerian budget, on the device.

Drop-in versions of the numpy functions of the reference's ``seaduck/eulerian_budget.py`` that do the
array work of a tracer budget: the halo ("buffer") builders across LLC faces, periodic edges and the
vertical ends (:214-342) and the advection schemes that give the tracer concentration on cell walls
(second-order flux limiter :363-445, third-order upwind :448, third-order direct space-time :477-532).
Same names, argument order and results (bit for bit); the work runs in ``csrc/stencil.cu`` behind
``sdk_gather_plane`` / ``sdk_wall_stencil`` / ``sdk_upwind3_z``.

Inputs may be numpy arrays (copied to the device, result copied back as numpy) or CUDA tensors (result
stays a CUDA tensor, nothing touches the host).  The xgcm-based helpers of the reference module
(``create_ecco_grid``, ``hor_div`` ...) are thin wrappers over xgcm and are not part of this path.
"""

from __future__ import annotations

import numpy as np

from . import _lib

__all__ = [
    "buffer_x_withface", "buffer_y_withface", "buffer_x_periodic", "buffer_y_periodic", "buffer_z_nearest",
    "second_order_flux_limiter_x", "second_order_flux_limiter_y", "second_order_flux_limiter_z",
    "third_order_upwind_z", "third_order_DST_x", "third_order_DST_y",
]

_SCHEME_DST3, _SCHEME_LIMITER, _SCHEME_LIMITER_Z = 0, 1, 2
_LEFT, _RIGHT, _DOWN, _UP = 2, 3, 1, 0  # the moves of Topology.ind_tend

# ---------------------------------------------------------------------------------------------------------
# halo maps (host, cached): flat source index of every point of the buffered plane, -1 = stays zero
# ---------------------------------------------------------------------------------------------------------


def _margin_sources(tp, face, axis, side, width):
    """Flat indices ``[width][n]`` (x margin: ``[n][width]``) into a ``(nface, ny, nx)`` plane of the cells
    that continue ``face`` beyond one of its edges, or None where the reference leaves zeros (no neighbour).

    The neighbours of a whole edge come from walking its two end cells across (like the reference, so an
    edge whose *corner* walk fails stays empty) and filling in the rectangle between them; the
    rectangle is turned by a quarter when the neighbouring face is rotated against this one.
    """
    nface, ny, nx = tp.h_shape[-3:]
    move = {("x", 0): _LEFT, ("x", 1): _RIGHT, ("y", 0): _DOWN, ("y", 1): _UP}[(axis, side)]
    if axis == "x":
        col = 0 if side == 0 else nx - 1
        far, near = (face, ny - 1, col), (face, 0, col)
    else:
        row = 0 if side == 0 else ny - 1
        far, near = (face, row, nx - 1), (face, row, 0)
    try:
        f1, y1, x1 = tp.ind_moves(far, [move] * width)
        f2, y2, x2 = tp.ind_moves(near, [move])
    except IndexError:
        return None
    del f2
    plane = np.arange(nface * ny * nx, dtype=np.int64).reshape(nface, ny, nx)
    block = plane[f1, min(y1, y2) : max(y1, y2) + 1, min(x1, x2) : max(x1, x2) + 1]
    want = (ny, width) if axis == "x" else (width, nx)
    if block.shape != want:
        # a rotated neighbour: x margins turn clockwise, y margins anticlockwise (utils.py:899-904)
        block = np.swapaxes(block[:, ::-1], 0, 1) if axis == "x" else np.swapaxes(block[::-1, :], 0, 1)
    if block.shape != want:
        raise ValueError(f"cannot fit the neighbours of face {face} ({block.shape}) into a margin of shape {want}")
    return block


def _withface_map(tp, face, axis, lm, rm):
    key = ("withface", axis, int(face), int(lm), int(rm))
    cache = tp.__dict__.setdefault("_halo_maps", {})
    if key not in cache:
        if rm <= 0 or lm <= 0:
            # the reference slices with [lm:-rm]: a zero margin is not something it can do
            raise ValueError("buffer_*_withface needs lm >= 1 and rm >= 1")
        nface, ny, nx = tp.h_shape[-3:]
        plane = np.arange(nface * ny * nx, dtype=np.int64).reshape(nface, ny, nx)
        if axis == "x":
            out = np.full((ny, nx + lm + rm), -1, np.int64)
            out[:, lm : lm + nx] = plane[face]
            left, right = _margin_sources(tp, face, "x", 0, lm), _margin_sources(tp, face, "x", 1, rm)
            if left is not None:
                out[:, :lm] = left
            if right is not None:
                out[:, lm + nx :] = right
        else:
            out = np.full((ny + lm + rm, nx), -1, np.int64)
            out[lm : lm + ny] = plane[face]
            low, high = _margin_sources(tp, face, "y", 0, lm), _margin_sources(tp, face, "y", 1, rm)
            if low is not None:
                out[:lm] = low
            if high is not None:
                out[lm + ny :] = high
        cache[key] = out
    return cache[key]


def _wrap_map(shape, axis, lm, rm, mode):
    """Map for a periodic (``mode='wrap'``) or nearest-value (``'edge'``) margin along ``axis`` of a
    plane of shape ``shape`` (the trailing dimensions that take part)."""
    n = shape[axis]
    if mode == "wrap" and (lm > n or rm > n):
        raise ValueError("margin wider than the array")
    src = np.arange(-lm, n + rm)
    src = np.mod(src, n) if mode == "wrap" else np.clip(src, 0, n - 1)
    plane = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
    return np.take(plane, src, axis=axis)


# ---------------------------------------------------------------------------------------------------------
# device plumbing
# ---------------------------------------------------------------------------------------------------------

_map_cache = {}


def _device_map(host_map, device, key=None):
    import torch  # noqa: PLC0415

    if key is not None:
        hit = _map_cache.get((key, str(device)))
        if hit is not None:
            return hit
    t = torch.from_numpy(np.ascontiguousarray(host_map).reshape(-1)).to(device)
    if key is not None:
        if len(_map_cache) > 256:
            _map_cache.clear()
        _map_cache[(key, str(device))] = t
    return t


def _to_device(a):
    """(contiguous fp64 CUDA tensor, was_numpy)"""
    import torch  # noqa: PLC0415

    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise TypeError("tensors given to seaduck_b200.eulerian_budget must live on a CUDA device (or pass numpy arrays)")
        return a.to(torch.float64).contiguous(), False
    if not torch.cuda.is_available():
        raise RuntimeError("seaduck_b200.eulerian_budget runs on the GPU only; no CUDA device is visible")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda(), True


def _is_numpy(a):
    import torch  # noqa: PLC0415

    return not isinstance(a, torch.Tensor)


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _gather(s, host_map, plane_ndim, out_plane_shape, key=None, keep_on_device=False):
    """Apply a halo map to the trailing ``plane_ndim`` dimensions of ``s``."""
    import torch  # noqa: PLC0415

    d, was_numpy = _to_device(s)
    lead = tuple(d.shape[: d.ndim - plane_ndim])
    n_outer = int(np.prod(lead)) if lead else 1
    src_plane = int(np.prod(d.shape[d.ndim - plane_ndim :]))
    m = _device_map(host_map, d.device, key)
    out = torch.empty(lead + tuple(out_plane_shape), dtype=torch.float64, device=d.device)
    _lib.check(
        _lib.lib().sdk_gather_plane(d.data_ptr(), n_outer, src_plane, m.data_ptr(), m.numel(), out.data_ptr(), _lib.current_stream()),
        "sdk_gather_plane",
    )
    return out if keep_on_device else _back(out, was_numpy)


def _wall_stencil(scheme, buf, vel, axis, buf_was_numpy=None):
    """``axis`` counts from the end (-1 X, -2 Y, -3 Z)."""
    import torch  # noqa: PLC0415

    b, np_b = _to_device(buf)
    v, np_v = _to_device(vel)
    if buf_was_numpy is not None:
        np_b = buf_was_numpy
    if v.device != b.device:
        raise ValueError("field and velocity live on different devices")
    n_out = b.shape[axis] - 3
    want = list(b.shape)
    want[axis] = n_out
    if n_out < 0 or list(v.shape) != want:
        # same condition under which numpy cannot broadcast the slices of the reference expression
        raise ValueError(f"operands could not be broadcast together: buffer {tuple(b.shape)} needs a velocity of shape {tuple(want)}, got {tuple(v.shape)}")
    n_inner = int(np.prod(b.shape[b.ndim + axis + 1 :])) if axis != -1 else 1
    n_outer = int(np.prod(b.shape[: b.ndim + axis])) if b.ndim + axis > 0 else 1
    out = torch.empty_like(v)
    _lib.check(
        _lib.lib().sdk_wall_stencil(scheme, b.data_ptr(), v.data_ptr(), out.data_ptr(), n_outer, n_out, n_inner, _lib.current_stream()),
        "sdk_wall_stencil",
    )
    return _back(out, np_b and np_v)


# ---------------------------------------------------------------------------------------------------------
# the reference's functions
# ---------------------------------------------------------------------------------------------------------


def buffer_x_withface(s, face, lm, rm, tp):
    """``s[..., face, :, :]`` with ``lm`` / ``rm`` more columns taken from the neighbouring faces (zeros
    where there is none).  Reference ``buffer_x_withface`` (eulerian_budget.py:214)."""
    if tuple(s.shape[-3:]) != tuple(tp.h_shape[-3:]):
        raise ValueError(f"the last three dimensions {tuple(s.shape[-3:])} are not (face, Y, X) = {tuple(tp.h_shape[-3:])}")
    m = _withface_map(tp, face, "x", lm, rm)
    return _gather(s, m, 3, m.shape, key=("x", tp.typ, tuple(tp.h_shape[-3:]), int(face), int(lm), int(rm)))


def buffer_y_withface(s, face, lm, rm, tp):
    """Same in Y: ``lm`` rows below, ``rm`` above.  Reference ``buffer_y_withface`` (:262)."""
    if tuple(s.shape[-3:]) != tuple(tp.h_shape[-3:]):
        raise ValueError(f"the last three dimensions {tuple(s.shape[-3:])} are not (face, Y, X) = {tuple(tp.h_shape[-3:])}")
    m = _withface_map(tp, face, "y", lm, rm)
    return _gather(s, m, 3, m.shape, key=("y", tp.typ, tuple(tp.h_shape[-3:]), int(face), int(lm), int(rm)))


def buffer_x_periodic(s, lm, rm, _keep=False):
    """Reference ``buffer_x_periodic`` (:309)."""
    nx = s.shape[-1]
    return _gather(s, _wrap_map((nx,), 0, lm, rm, "wrap"), 1, (nx + lm + rm,), ("xp", nx, lm, rm), _keep)


def buffer_y_periodic(s, lm, rm, _keep=False):
    """Reference ``buffer_y_periodic`` (:321)."""
    ny, nx = s.shape[-2:]
    return _gather(s, _wrap_map((ny, nx), 0, lm, rm, "wrap"), 2, (ny + lm + rm, nx), ("yp", ny, nx, lm, rm), _keep)


def buffer_z_nearest(s, lm, rm, _keep=False):
    """Reference ``buffer_z_nearest`` (:333): the end levels repeated."""
    nz, ny, nx = s.shape[-3:]
    return _gather(s, _wrap_map((nz, ny, nx), 0, lm, rm, "edge"), 3, (nz + lm + rm, ny, nx), ("zn", nz, ny, nx, lm, rm), _keep)


def second_order_flux_limiter_x(s_center, u_cfl):
    """Reference ``second_order_flux_limiter_x`` (:363): superbee-limited wall concentration, X periodic."""
    return _wall_stencil(_SCHEME_LIMITER, buffer_x_periodic(s_center, 2, 2, _keep=True), u_cfl, -1, _is_numpy(s_center))


def second_order_flux_limiter_y(s_center, u_cfl):
    """Reference ``second_order_flux_limiter_y`` (:391)."""
    return _wall_stencil(_SCHEME_LIMITER, buffer_y_periodic(s_center, 2, 2, _keep=True), u_cfl, -2, _is_numpy(s_center))


def second_order_flux_limiter_z(s_center, u_cfl):
    """Reference ``second_order_flux_limiter_z`` (:419)."""
    return _wall_stencil(_SCHEME_LIMITER_Z, buffer_z_nearest(s_center, 2, 1, _keep=True), u_cfl, -3, _is_numpy(s_center))


def third_order_upwind_z(s, w):
    """Reference ``third_order_upwind_z`` (:448).  Like there, ``w[0]`` is set to zero in place."""
    import torch  # noqa: PLC0415

    sd, np_s = _to_device(s)
    wd, np_w = _to_device(w)
    if sd.shape != wd.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(sd.shape)} {tuple(wd.shape)}")
    nz = sd.shape[0]
    n_inner = int(np.prod(sd.shape[1:])) if sd.ndim > 1 else 1
    out = torch.empty_like(sd)
    _lib.check(_lib.lib().sdk_upwind3_z(sd.data_ptr(), wd.data_ptr(), out.data_ptr(), nz, n_inner, _lib.current_stream()), "sdk_upwind3_z")
    # the reference's side effect on the caller's array
    if np_w:
        if nz:
            w[0] = 0
    elif wd.data_ptr() != w.data_ptr() and nz:
        w[0] = 0
    return _back(out, np_s and np_w)


def third_order_DST_x(xbuffer, u_cfl):
    """Reference ``third_order_DST_x`` (:477): ``xbuffer`` carries two more columns before and one after."""
    return _wall_stencil(_SCHEME_DST3, xbuffer, u_cfl, -1)


def third_order_DST_y(ybuffer, u_cfl):
    """Reference ``third_order_DST_y`` (:506)."""
    return _wall_stencil(_SCHEME_DST3, ybuffer, u_cfl, -2)
