"""Interpolation kernels (``KnW``) and their weights.

Public surface of the reference's ``seaduck.kernel_weight`` (``KnW`` ``kernel_weight.py:645``,
``get_weight`` :772, ``kernel_weight`` :431, cascade helpers :481-584, ``_translate_to_tendency``
:46).  A ``KnW`` is a *description*: a horizontal stencil, the nested smaller stencils to fall back
on next to land ("inheritance"), and what to do in x/y, z and t (interpolate or differentiate).
The CUDA interpolation kernel evaluates the same weights on the device from the compact descriptor
built by :meth:`KnW.describe`; the numpy evaluation here (:meth:`KnW.get_weight`) is the host mirror
kept for API compatibility.

Weights are Lagrange-polynomial weights.  For a one-dimensional set of nodes ``xs`` the weight of
node ``x`` for the ``k``-th derivative is, as in the reference (:84-306),

    sum over all subsets S of the other nodes with |S| = len(xs)-1-k of prod_{o in S}(r - o)
    -----------------------------------------------------------------------------------------
                                prod_{o != x} (x - o)

(without the ``k!`` factor, except at the maximum order where the reference uses ``(k-1)!``).
"""

from __future__ import annotations

import logging
from itertools import combinations
from math import factorial

import numpy as np

DEFAULT_KERNEL = np.array([[0, 0], [0, 1], [0, 2], [0, -1], [0, -2], [-1, 0], [-2, 0], [1, 0], [2, 0]])
DEFAULT_INHERITANCE = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8],
    [0, 1, 2, 3, 5, 7, 8],
    [0, 1, 3, 5, 7],
    [0],
]
DEFAULT_KERNELS = [np.array([DEFAULT_KERNEL[i] for i in doll]) for doll in DEFAULT_INHERITANCE]

# stencil kinds understood by the device code
KIND_CROSS_INTERP, KIND_LINE_X, KIND_LINE_Y, KIND_RECT = 0, 1, 2, 3


def _translate_to_tendency(kernel):
    """Moves (0 up, 1 down, 2 left, 3 right) that lead from [0,0] to node ``[x, y]`` (ref. :46)."""
    x, y = int(kernel[0]), int(kernel[1])
    tend = [0] * y if y > 0 else [1] * (-y)
    tend += [2] * (-x) if x < 0 else [3] * x
    return tend


def _axis_weight(nodes, target, r, order):
    """Weight of node ``target`` among 1-D ``nodes`` at position(s) ``r`` for derivative ``order``."""
    others = [o for o in nodes if o != target]
    denom = 1.0
    for o in others:
        denom *= target - o
    max_order = len(nodes) - 1
    if order > max_order:
        raise ValueError("Kernel is too small for this derivative")
    if order == max_order:
        num = np.ones_like(r, dtype=float) * float(factorial(max(order - 1, 0)))
    else:
        num = np.zeros_like(r, dtype=float)
        for subset in combinations(others, len(others) - order):
            term = np.ones_like(r, dtype=float)
            for o in subset:
                term = term * (r - o)
            num = num + term
    return num / denom


class _Stencil:
    """One horizontal stencil (a level of the inheritance) and how to weigh it."""

    def __init__(self, kernel, hkernel="interp", h_order=0):
        kernel = np.asarray(kernel)
        self.kernel = kernel
        xs = sorted(set(int(v) for v in kernel[:, 0]))
        ys = sorted(set(int(v) for v in kernel[:, 1]))
        self.xs, self.ys = xs, ys
        m = len(kernel)
        if m == len(xs) + len(ys) - 1:
            typ = hkernel[1:] if "d" in hkernel else hkernel
            if len(xs) == 1:
                typ = "y"
            elif len(ys) == 1:
                typ = "x"
            if typ == "interp":
                self.kind, self.xorder, self.yorder = KIND_CROSS_INTERP, 0, 0
            elif typ == "x":
                self.kind, self.xorder, self.yorder = KIND_LINE_X, h_order, 0
                if h_order > len(xs) - 1:
                    raise ValueError("Kernel is too small for this derivative")
            elif typ == "y":
                self.kind, self.xorder, self.yorder = KIND_LINE_Y, 0, h_order
                if h_order > len(ys) - 1:
                    raise ValueError("Kernel is too small for this derivative")
            else:
                raise ValueError(f"kernel_type = {hkernel} not supported")
        elif m == len(xs) * len(ys):
            self.kind = KIND_RECT
            if hkernel == "interp":
                self.xorder, self.yorder = 0, 0
            elif hkernel == "dx":
                self.xorder, self.yorder = h_order, 0
            elif hkernel == "dy":
                self.xorder, self.yorder = 0, h_order
            else:
                raise ValueError(f"kernel_type = {hkernel} not supported")
            if self.xorder > len(xs) - 1 or self.yorder > len(ys) - 1:
                raise ValueError("Kernel is too small for this derivative")
        else:
            raise NotImplementedError("The shape of the kernel is neither cross or square")

    def __call__(self, rx, ry):
        rx = np.asarray(rx, float)
        ry = np.asarray(ry, float)
        n = len(rx)
        w = np.zeros((n, len(self.kernel)))
        for i, (x, y) in enumerate(self.kernel):
            x, y = int(x), int(y)
            if self.kind == KIND_CROSS_INTERP:
                if x != 0:
                    w[:, i] = _axis_weight(self.xs, x, rx, 0)
                elif y != 0:
                    w[:, i] = _axis_weight(self.ys, y, ry, 0)
                else:
                    w[:, i] = _axis_weight(self.xs, 0, rx, 0) + _axis_weight(self.ys, 0, ry, 0) - 1
            elif self.kind == KIND_LINE_X:
                if y == 0:
                    w[:, i] = _axis_weight(self.xs, x, rx, self.xorder)
            elif self.kind == KIND_LINE_Y:
                if x == 0:
                    w[:, i] = _axis_weight(self.ys, y, ry, self.yorder)
            else:
                w[:, i] = _axis_weight(self.xs, x, rx, self.xorder) * _axis_weight(self.ys, y, ry, self.yorder)
        return w


def kernel_weight(kernel, kernel_type="interp", order=0):
    """Function ``f(rx, ry) -> (N, M)`` weights of a cross or rectangular kernel (ref. :431)."""
    return _Stencil(kernel, kernel_type, order)


def kernel_weight_x(kernel, kernel_type="interp", order=0):
    """Cross-shaped kernels (reference :84)."""
    hk = {"x": "dx", "y": "dy"}.get(kernel_type, kernel_type)
    return _Stencil(kernel, hk, order)


def kernel_weight_s(kernel, xorder=0, yorder=0):
    """Rectangular kernels (reference :313)."""
    st = _Stencil(kernel, "interp", 0)
    if st.kind != KIND_RECT:
        # a line is both a cross and a rectangle; force the rectangular rule
        st.kind = KIND_RECT
    if xorder > len(st.xs) - 1 or yorder > len(st.ys) - 1:
        raise ValueError("Kernel is too small for this derivative")
    st.xorder, st.yorder = xorder, yorder
    return st


default_interp_funcs = [kernel_weight_x(a_kernel) for a_kernel in DEFAULT_KERNELS]


def find_which_points_for_each_kernel(masked, inheritance="default"):
    """For every stencil of the inheritance, the points that use it (reference :481)."""
    if isinstance(inheritance, str) and inheritance == "default":
        inheritance = DEFAULT_INHERITANCE
    masked = np.asarray(masked)
    chosen = np.full(masked.shape[0], -1)
    for level, doll in enumerate(inheritance):
        wet = masked[:, np.array(doll)].all(axis=1)
        chosen[(chosen < 0) & wet] = level
    return [list(np.where(chosen == level)[0]) for level in range(len(inheritance))]


def get_weight_cascade(rx, ry, pk, kernel_large=DEFAULT_KERNEL, inheritance=DEFAULT_INHERITANCE, funcs=default_interp_funcs):
    """Weights with the stencil chosen per point; NaN in column 0 where no stencil fits (ref. :525)."""
    weight = np.zeros((len(rx), len(kernel_large)))
    weight[:, 0] = np.nan
    for i in range(len(pk)):
        if len(pk[i]) == 0:
            continue
        sub_weight = np.zeros((len(pk[i]), len(kernel_large)))
        sub_weight[:, np.array(inheritance[i])] = funcs[i](rx[pk[i]], ry[pk[i]])
        weight[pk[i]] = sub_weight
    return weight


def _find_pk_4d(masked, inheritance=DEFAULT_INHERITANCE):
    """Stencil choice for every vertical / temporal node (reference :571)."""
    maskedT = masked.T
    pk4d = []
    for i in range(maskedT.shape[0]):
        z = []
        for j in range(maskedT.shape[1]):
            z.append(find_which_points_for_each_kernel(maskedT[i, j].T, inheritance))
        pk4d.append(z)
    return pk4d


# Nodes at the same distance from the centre are ordered by a slightly skewed distance, so that +x comes before +y
# before -x / -y exactly as in the reference's node order (kernel_weight.py:623-642, :675-680): the skew is the
# reference's and part of the observable behaviour (it fixes the column order of every weight matrix).
_SKEW = np.array([0.01, 0.00618])
_LEVELS_PER_MODE = {"dz": 2, "linear": 2, "dt": 2, "nearest": 1}


def _node_key(kernel):
    return tuple((int(i), int(j)) for (i, j) in kernel)


def _auto_inheritance(kernel, hkernel="interp"):
    """The nested stencils a kernel falls back on next to land, largest first (reference :623): all nodes (for a
    derivative: the nodes on the axis it acts along), then the nodes within Chebyshev distance R - 1, R - 2 ... 0 of the
    centre, every level in skewed-distance order and listed only when it differs from the one before."""
    kernel = np.asarray(kernel)
    on_axis = {"interp": np.ones(len(kernel), bool), "dx": kernel[:, 1] == 0, "dy": kernel[:, 0] == 0}[hkernel]
    reach = np.abs(kernel).max(axis=1)
    skewed = np.abs(kernel + _SKEW).max(axis=1)
    members = np.flatnonzero(on_axis)
    members = members[np.argsort(skewed[members], kind="stable")].tolist()
    levels = [members]
    for radius in range(int(round(reach[members].max())) - 1, -1, -1):
        inner = [i for i in levels[-1] if reach[i] <= radius]
        if inner != levels[-1]:
            levels.append(inner)
    return levels


class KnW:
    """Kernel object: everything about the interpolation / derivative kernel (reference :645)."""

    def __init__(
        self,
        kernel=DEFAULT_KERNEL,
        inheritance="auto",
        hkernel="interp",
        vkernel="nearest",
        tkernel="nearest",
        h_order=0,
        ignore_mask=False,
    ):
        given = np.asarray(kernel)
        if inheritance is not None and ignore_mask:
            logging.info("Overwriting the inheritance object to None, because we ignore masking")
        if isinstance(inheritance, str):
            if inheritance != "auto":
                raise ValueError("Unknown type of inherirance")
            inheritance = _auto_inheritance(given, hkernel=hkernel)
        elif inheritance is None:
            inheritance = [list(range(len(given)))]
        elif not isinstance(inheritance, list):
            raise ValueError("Unknown type of inherirance")
        # canonical node order: by (skewed) L1 distance from the centre; the levels are re-expressed in it
        order = np.abs(given + _SKEW).sum(axis=1).argsort()
        position = np.empty(len(order), int)
        position[order] = np.arange(len(order))
        self.kernel = given[order]
        self.inheritance = [sorted(int(position[i]) for i in level) for level in inheritance]
        self.hkernel, self.vkernel, self.tkernel = hkernel, vkernel, tkernel
        self.h_order, self.ignore_mask = h_order, ignore_mask
        self.kernels = [self.kernel[np.array(level, int)] for level in self.inheritance]
        self.hfuncs = [_Stencil(nodes, hkernel, h_order) for nodes in self.kernels]

    def same_hsize(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("the argument is not a KnW object")
        try:
            return np.allclose(self.kernel, other.kernel)
        except (ValueError, AttributeError):
            return False

    def same_size(self, other):
        return self.same_hsize(other) and self.nz == other.nz and self.nt == other.nt

    def _modes(self):
        return (self.hkernel, self.vkernel, self.tkernel)

    def __eq__(self, other):
        # (like the reference, h_order and ignore_mask take part in the hash but not in the comparison)
        if not isinstance(other, type(self)):
            return False
        return self.same_hsize(other) and self.inheritance == other.inheritance and self._modes() == other._modes()

    def __hash__(self):
        levels = tuple(tuple(level) for level in self.inheritance)
        return hash((_node_key(self.kernel), levels, self.ignore_mask, self.h_order) + self._modes())

    def size_hash(self):
        return hash((_node_key(self.kernel), self.nz, self.nt))

    @property
    def nz(self):
        return _LEVELS_PER_MODE[self.vkernel]

    @property
    def nt(self):
        return _LEVELS_PER_MODE[self.tkernel]

    @staticmethod
    def _line_factors(mode, r, n, interp, deriv):
        """Factors of the (up to two) nodes along z or t as an ``(n, levels)`` array: ``[1 - r, r]`` to interpolate,
        ``[-1, 1]`` to differentiate, ``[1]`` for the nearest node."""
        if mode == interp:
            r = np.full(n, float(r)) if np.ndim(r) == 0 else np.array(r, float)
            return np.stack([1 - r, r], axis=1)
        if mode == deriv:
            return np.tile(np.array([-1.0, 1.0]), (n, 1))
        return np.ones((n, 1))

    def get_weight(self, rx, ry, rz=0, rt=0, pk4d=None, bottom_scheme="no flux"):
        """Weights ``(N, M, nz, nt)`` given the rel-coords (reference :772).  Host mirror of what the CUDA interpolation
        kernel evaluates per point: horizontal Lagrange weights of the stencil each (z, t) node can use
        (``pk4d`` from :func:`_find_pk_4d`; the full stencil without it), times the vertical and the temporal factor."""
        rx, ry = np.asarray(rx, float), np.asarray(ry, float)
        n, nz, nt = len(rx), self.nz, self.nt
        weight = np.zeros((n, len(self.kernel), nz, nt))
        if pk4d is None:
            weight[:, self.inheritance[0]] = self.hfuncs[0](rx, ry)[:, :, None, None]
        else:
            if nt != len(pk4d) or nz != len(pk4d[0]):
                raise ValueError("The kernel and the input pk4d does not match")
            for jt in range(nt):
                for jz in range(nz):
                    weight[:, :, jz, jt] = get_weight_cascade(
                        rx, ry, pk4d[jt][jz], kernel_large=self.kernel, inheritance=self.inheritance, funcs=self.hfuncs
                    )
        along_z = self._line_factors(self.vkernel, rz, n, "linear", "dz")
        along_t = self._line_factors(self.tkernel, rt, n, "linear", "dt")
        ghost_below = self.vkernel == "linear" and bottom_scheme == "no flux"
        for jt in range(nt):
            if ghost_below:
                # No-flux bottom: where the deeper of the two levels (index 0) has no usable stencil it is treated as a
                # ghost point carrying the value of the level above -- all the weight goes to level 1 -- unless the
                # point lies in the lower half of the layer (rz < 1/2), which is inside the bottom: no weight at all.
                # (The factor of level 1 stays 1 for the time nodes that follow, as in the reference.)
                dry_below = np.isnan(weight[:, :, 0, jt]).any(axis=1)
                weight[dry_below, :, 0, jt] = 0
                weight[dry_below & (np.asarray(rz) < 0.5), :, 1, jt] = 0
                along_z[dry_below, 1] = 1
            weight[:, :, :, jt] = weight[:, :, :, jt] * along_z[:, None, :] * along_t[:, None, None, jt]
        return weight

    def describe(self):
        """Compact description for the device: nodes, inheritance levels, 1-D node sets per level."""
        levels = []
        for doll, st in zip(self.inheritance, self.hfuncs):
            levels.append(
                dict(nodes=list(doll), kind=st.kind, xs=list(st.xs), ys=list(st.ys), xorder=st.xorder, yorder=st.yorder)
            )
        vmode = {"nearest": 0, "linear": 1, "dz": 2}[self.vkernel]
        tmode = {"nearest": 0, "linear": 1, "dt": 2}[self.tkernel]
        return dict(kernel=self.kernel.astype(int), levels=levels, vmode=vmode, tmode=tmode, ignore_mask=bool(self.ignore_mask))


def get_func(kernel, hkernel="interp", h_order=0, **kwargs):
    """Weight function of a kernel (reference ``kernel_weight.py:608``; no caching needed here)."""
    return _Stencil(np.asarray(kernel), hkernel, h_order)
