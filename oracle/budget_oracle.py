"""TEST INFRASTRUCTURE -- numpy restatement of the crossing-log post-processing of the Lagrangian budget
(reference ``seaduck/lagrangian_budget.py``: ``find_ind_frac_tres`` :245 with ``by_type=True``,
``first_last_neither`` :114, ``which_wall`` :85, ``wall_index`` :192, ``ind_tend_uv`` :163,
``pseudo_motion`` :97, ``redo_index`` :223, ``residence_time`` :139, ``tres_update`` :144;
``_time2wall`` / ``_which_early`` ``utils.py:859,875``).

Parity status: PINNED -- ``tests/test_budget_oracle.py`` feeds it the crossing logs of the Lagrangian
goldens and compares with ``tests/golden/budget_*.npz`` (both made by the unmodified reference).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as ora

MOVE = {0: (0, 0), 1: (0, 1), 2: (0, 0), 3: (1, 0), 4: (0, 0), 5: (0, 0)}  # delta_y, delta_x


def first_last_neither(shapes):
    acc = np.cumsum(shapes)
    last = acc - 1
    first = np.roll(acc, 1)
    first[0] = 0
    neither = np.array([first[i] + j for i, n in enumerate(shapes) for j in range(1, n - 1)], dtype=int)
    return first, last, neither


def stationary_time(u, du, x0):
    """utils.py:821"""
    with np.errstate(all="ignore"):
        tl = np.log(1 - du / u * (0.5 + x0)) / du
        tr = np.log(1 + du / u * (0.5 - x0)) / du
        flat = np.abs(du) < 1e-12
        tl[flat] = (-x0[flat] - 0.5) / u[flat]
        tr[flat] = (0.5 - x0[flat]) / u[flat]
    return tl, tr


def time2wall(xs, us, dus, tf):
    """utils.py:859"""
    ts = []
    sign = np.sign(tf)
    for x0, u, du in zip(xs, us, dus):
        tl, tr = stationary_time(u, du, x0)
        ul = u - (x0 + 0.5) * du
        ur = u + (0.5 - x0) * du
        tl[-ul * sign <= 1e-12] = -sign[-ul * sign <= 1e-12]
        tr[ur * sign <= 1e-12] = -sign[ur * sign <= 1e-12]
        ts += [tl, tr]
    return ts


def which_early(tf, ts):
    """utils.py:875"""
    ts = ts + [np.ones(len(ts[0])) * tf]
    with np.errstate(all="ignore"):
        directed = np.array(ts) * np.sign(tf)
    directed[np.isnan(directed)] = np.inf
    directed[directed < 0] = np.inf
    tend = directed.argmin(axis=0)
    return tend, np.array(ts)[tend, np.arange(len(tend))]


def wall_index(og, izl, face, iy, ix, iwall):
    """lagrangian_budget.py:192"""
    lib = ora.lib()
    iw = iwall // 2
    iz = izl.copy()
    ny = iy + np.array([MOVE[int(k)][0] for k in iwall], dtype=int)
    nx = ix + np.array([MOVE[int(k)][1] for k in iwall], dtype=int)
    iz[iwall == 4] += 1
    iz -= 1
    if og.typ != "LLC":
        return np.array([iw, iz, ny, nx]).astype(int)
    nf = face.copy()
    gny, gnx = og.g_shape[-2:]
    for j in np.where((ny > gny - 1) | (nx > gnx - 1))[0]:
        ind = (C.c_int32 * 3)(int(face[j]), int(iy[j]), int(ix[j]))
        out = (C.c_int32 * 3)()
        new_iw = C.c_int(0)
        err = lib.ora_ind_tend_uv(C.byref(og.c), int(iw[j]), ind, C.byref(new_iw), out)
        if err:
            raise IndexError(f"no wall next to cell {tuple(ind)}")
        iw[j], nf[j], ny[j], nx[j] = new_iw.value, out[0], out[1], out[2]
    return np.array([iw, iz, nf, ny, nx]).astype(int)


def redo_index(og, rec, sel):
    """lagrangian_budget.py:223 for the records ``sel``"""
    g = lambda k: np.asarray(rec[k])[sel]  # noqa: E731
    xs = [g("rx"), g("ry"), g("rzl") - 1 / 2]
    us, dus = [g("uu"), g("vv"), g("ww")], [g("du"), g("dv"), g("dw")]
    n = len(xs[0])
    tendf, tf = which_early(1e80, time2wall(xs, us, dus, 1e80 * np.ones(n)))
    tendb, tb = which_early(-1e80, time2wall(xs, us, dus, -1e80 * np.ones(n)))
    tendf[tendf == 6] = 0
    tendb[tendb == 6] = 0
    face = g("fc").astype(int) if og.has_face else np.zeros(n, int)
    inds = (g("izl").astype(int), face, g("iy").astype(int), g("ix").astype(int))
    vf = wall_index(og, *inds, tendf)
    vb = wall_index(og, *inds, tendb)
    with np.errstate(all="ignore"):
        frac = -tb / (tf - tb)
    return vf, vb, np.nan_to_num(frac, nan=1)


def find_ind_frac_tres(og, rec, offset):
    """lagrangian_budget.py:245; ``rec`` = dict of flat log arrays, ``offset`` = per-particle starts (n+1)."""
    shapes = np.diff(offset)
    n = int(offset[-1])
    first, last, neither = first_last_neither(shapes)
    rows = 5 if og.has_face else 4
    ind1 = np.zeros((rows, n), "int16")
    ind2 = np.ones((rows, n), "int16")
    frac = np.ones(n)
    g = lambda k, sel: np.asarray(rec[k])[sel]  # noqa: E731
    if len(neither) > 0:
        rx, ry, rzl = g("rx", neither), g("ry", neither), g("rzl", neither)
        dist = np.vstack([np.abs(0.5 + rx), np.abs(0.5 - rx), np.abs(0.5 + ry), np.abs(0.5 - ry), np.abs(rzl), np.abs(1 - rzl)])
        iwall = np.argmin(dist, axis=0)  # which_wall :85
        face = g("fc", neither).astype(int) if og.has_face else np.zeros(len(neither), int)
        ind1[:, neither] = wall_index(og, g("izl", neither).astype(int), face, g("iy", neither).astype(int), g("ix", neither).astype(int), iwall)
    ind1[:, first], ind2[:, first], frac[first] = redo_index(og, rec, first)
    ind1[:, last], ind2[:, last], frac[last] = redo_index(og, rec, last)
    # residence_time :139, tres_update :144
    u, du, x = np.asarray(rec["uu"]), np.asarray(rec["du"]), np.asarray(rec["rx"])
    v, dv, y = np.asarray(rec["vv"]), np.asarray(rec["dv"]), np.asarray(rec["ry"])
    w, dw, z = np.asarray(rec["ww"]), np.asarray(rec["dw"]), np.asarray(rec["rzl"]) - 0.5
    ulist = np.array([u - (x + 0.5) * du, u + (0.5 - x) * du, v - (y + 0.5) * dv, v + (0.5 - y) * dv, w - (z + 0.5) * dw, w + (0.5 - z) * dw])
    with np.errstate(all="ignore"):
        tres0 = 2 / np.sum(np.abs(ulist), axis=0)
    fracs = np.ones(n)
    fracs[first + 1] = frac[first]
    fracs[last] -= frac[last]
    with np.errstate(all="ignore"):
        tres = tres0 * fracs
    tres[np.asarray(rec["vs"]) > 6] = 0.0
    return ind1, ind2, frac, tres, last, first


def wall_concentration_array(fields, xname="sxprime", yname="syprime", zname="szprime"):
    """lagrangian_budget.py:548 (``prefetch_vector``): the three wall fields stacked, padded to a common shape."""
    parts = [np.asarray(fields[k], float) for k in (xname, yname, zname)]
    shape = tuple(max(p.shape[d] for p in parts) for d in range(parts[0].ndim))
    out = np.full((3,) + shape, np.nan)
    for k, p in enumerate(parts):
        out[(k,) + tuple(slice(0, m) for m in p.shape)] = p
    return out


def calculate_budget(rec, offset, ind1, ind2, frac, tres, fields, rhs_list, lhs_name="lhs", dir_of_time=-1, wall_names=None):
    """lagrangian_budget.py:603 with ``lhs_contribution`` :575 and ``contr_p_relaxed`` :583 (p = 1), record by
    record.  ``rec`` = flat log arrays (tt, izl, fc, iy, ix), ``fields`` = name -> array of one model step."""
    walls = wall_concentration_array(fields, *(wall_names or ()))
    n = int(offset[-1])
    is_last = np.zeros(n, bool)
    is_last[np.asarray(offset[1:]) - 1] = True
    conc = np.empty(n)
    for i in range(n):
        s1 = walls[tuple(int(v) for v in ind1[:, i])]
        s2 = walls[tuple(int(v) for v in ind2[:, i])]
        conc[i] = s1 * frac[i] + (1 - frac[i]) * s2
    names = list(rhs_list)
    out = {k: np.empty(n - 1) for k in names + ["error", lhs_name]}
    tt = np.asarray(rec["tt"], float)
    with np.errstate(all="ignore"):
        for i in range(n - 1):
            cell = (int(rec["izl"][i]) - 1, int(rec["fc"][i]), int(rec["iy"][i]), int(rec["ix"][i]))
            if is_last[i]:
                delta = span = dt = 0.0
            else:
                delta = float(np.nan_to_num(conc[i + 1] - conc[i]))
                span = dir_of_time * tres[i + 1]
                dt = float(np.nan_to_num(tt[i + 1] - tt[i]))
            lhs = dt * fields[lhs_name][cell]
            target = delta - lhs
            x = [np.float64(fields[k][cell]) for k in names]
            deno, total = np.float64(0.0), np.float64(0.0)
            for v in x:
                deno = deno + v * v
                total = total + v
            gap = target - total * span
            acc = np.float64(0.0)
            for k, v in zip(names, x):
                c = v * span + (v * v) / deno * gap
                out[k][i] = c
                acc = acc + c
            out["error"][i] = 0.0 + (target - acc)
            out[lhs_name][i] = lhs
    return out, conc


def read_wall_list(og, rec, prefetch, scalar=True):
    """lagrangian_budget.py:406 for grids with faces: the values of three wall fields (x walls, y walls, z walls) on
    the six walls of the cell of every record -> (6, n): left, right, bottom-y, top-y, deep, shallow."""
    if not og.has_face:
        raise NotImplementedError("the oracle restates the branch for grids with faces")
    ua, va, wa = (np.asarray(p, float) for p in prefetch)
    fc, iy, ix = (np.asarray(rec[k]).astype(int) for k in ("fc", "iy", "ix"))
    izl = np.asarray(rec["izl"]).astype(int)
    n = len(fc)
    right, up = np.empty((3, n), int), np.empty((3, n), int)
    for j in range(n):
        here = (fc[j], iy[j], ix[j])
        for move, out in ((3, right), (0, up)):
            err, new = og.ind_tend(here, move)  # ind_tend_vec :526 keeps the old index where there is no neighbour
            out[:, j] = here if err else new
    # rotation of the neighbouring face, cos / sin as they come (topology.py:248-253: not rounded here)
    lib = ora.lib()
    rot_r, rot_u = np.zeros((n, 4)), np.zeros((n, 4))
    m4 = (C.c_double * 4)()
    for j in range(n):
        for nf, out in ((right[0, j], rot_r), (up[0, j], rot_u)):
            if lib.ora_uv_rotation_raw(C.byref(og.c), int(fc[j]), int(nf), m4):
                raise ValueError("faces not connected")
            out[j] = m4[:]
    if scalar:
        rot_r, rot_u = np.abs(rot_r), np.abs(rot_u)
    lev = izl - 1
    at = lambda a, k, idx: np.nan_to_num(a[k, idx[0], idx[1], idx[2]])  # noqa: E731
    ur = at(ua, lev, right) * rot_r[:, 0] + at(va, lev, right) * rot_r[:, 1]
    vr = at(ua, lev, up) * rot_u[:, 2] + at(va, lev, up) * rot_u[:, 3]
    own = (fc, iy, ix)
    return np.array([np.nan_to_num(v) for v in (ua[(lev,) + own], ur, va[(lev,) + own], vr, wa[(izl,) + own], wa[(lev,) + own])])


def particle_data_compat(og, rec, fields, wall_names=("sx", "sy", "sz"), conv_name="divus"):
    """lagrangian_budget.py:470 (``check_particle_data_compat`` with ``debug=True``): wall velocities from (u, du, r)
    (``_uleftright_from_udu`` utils.py:852), wall values, the convergence of their products (``crude_convergence``
    :464) and the Eulerian convergence at the records' cells."""
    c_list = read_wall_list(og, rec, [fields[k] for k in wall_names])
    g = lambda k: np.asarray(rec[k], float)  # noqa: E731
    walls = []
    for u, du, r in ((g("uu"), g("du"), g("rx")), (g("vv"), g("dv"), g("ry")), (g("ww"), g("dw"), g("rzl") - 0.5)):
        walls += [u - (r + 0.5) * du, u + (0.5 - r) * du]
    u_list = np.array(walls)
    flux = c_list * u_list
    conv = np.zeros(flux.shape[1])
    for k, sign in enumerate((1, -1, 1, -1, 1, -1)):
        conv = conv + flux[k] * sign
    cell = (np.asarray(rec["izl"]).astype(int) - 1,) + tuple(np.asarray(rec[k]).astype(int) for k in ("fc", "iy", "ix"))
    return u_list, c_list, conv, np.asarray(fields[conv_name], float)[cell]
