"""Wet / dry masks at cell centres and on the walls, and the two particle callbacks built on them.

Public surface of the reference's ``seaduck/get_masks.py``: ``get_mask_arrays`` (:110), ``get_masked`` (:169),
``which_not_stuck`` (:219), ``abandon_stuck`` (:231).  The wall masks themselves (``mask_u_node`` :12,
``mask_v_node`` :46, ``mask_w_node`` :79: a wall is wet when a cell on either side of it is) are built on the device
by ``sdk_build_masks`` -- the interpolation kernels read them there -- and fetched to the host only when one of
these functions is asked for numpy values.
"""

from __future__ import annotations

import warnings

import numpy as np

from .ocedata import RelCoord

_KINDS = ("C", "U", "V", "Wvel")


def _host_masks(od):
    """{"C" | "U" | "V" | "Wvel": numpy uint8 array} in the padded shapes the reference's ``get_masked`` uses, or {}."""
    hit = od._interp_cache.get("host_masks")
    if hit is None:
        from .interp import _device_masks  # noqa: PLC0415

        hit = {k: v.cpu().numpy() for k, v in _device_masks(od).items()}
        od._interp_cache["host_masks"] = hit
    return hit


def get_mask_arrays(od):
    """``(maskC, maskU, maskV, maskW)``; all ones (with a warning) for a dataset without ``maskC``."""
    masks = _host_masks(od)
    if not masks:
        warnings.warn("no maskC in the dataset, assuming nothing is masked.")
        ones = np.ones_like(np.asarray(od._ds["XC"]) + np.asarray(od._ds["Z"]).reshape((-1,) + (1,) * np.asarray(od._ds["XC"]).ndim))
        return ones, ones, ones, ones
    return tuple(masks[k] for k in _KINDS)


def get_masked(od, ind, cuvwg="C"):
    """Mask values at the grid points ``ind`` (tuple of index arrays, vertical index first)."""
    if cuvwg not in _KINDS:
        raise NotImplementedError("cuvwg(the kind of grid point) for mask not supported")
    masks = _host_masks(od)
    if not masks:
        warnings.warn("no maskC in the dataset, assuming nothing is masked.")
        return np.ones_like(ind[0])
    return masks[cuvwg][tuple(np.asarray(i) for i in ind)]


def which_not_stuck(p):
    """True for the particles whose cell is wet -- usable as ``Particle(callback=which_not_stuck)``: particles that
    ran aground stop being integrated."""
    ind = [] if p.izl_lin is None else [p.izl_lin - 1]
    if p.face is not None:
        ind.append(p.face)
    ind += [p.iy, p.ix]
    return get_masked(p.ocedata, tuple(ind)).astype(bool)


def abandon_stuck(p):
    """Drop the particles that sit in dry cells from a host-side particle set (a ``Position`` or a ``Particle.subset``):
    every per-particle array and rel-coordinate is cut down to the wet ones, in place."""
    wet = which_not_stuck(p)
    for name, item in list(vars(p).items()):
        if name.startswith("_"):
            continue
        if isinstance(item, np.ndarray) and item.ndim == 1:
            object.__setattr__(p, name, item[wet])
        elif isinstance(item, RelCoord):
            object.__setattr__(p, name, item.subset(wet))
    object.__setattr__(p, "N", int(wet.sum()))
