"""Grid topology: what is connected to what (host side).

API mirror of the reference's ``seaduck.topology.Topology`` (``seaduck/topology.py:257``):
``ind_tend``, ``ind_moves``, ``ind_tend_vec``, ``check_illegal``, ``get_the_other_edge``,
``mutual_direction``, ``get_uv_mask_from_face``, ``four_matrix_for_uv`` keep their names, argument
meaning and error behaviour.  The implementation is different: every operation is a *vectorised*
table lookup over whole index arrays (the reference walks one point at a time through numba
functions), because the same routines also build the lookup tables that the CUDA kernels read:

* the 4-entry-per-face neighbour tables (``nbr_face``/``nbr_edge``) that live in ``__constant__``
  memory and drive the closed-form cell move across a face edge (reference ``_llc_ind_tend``,
  ``topology.py:149``),
* the corner (G point) tables of the cells on the top row / right column of every face
  (reference ``find_px_py`` ``utils.py:495`` -> ``_ind_tend_G`` ``topology.py:649``),
* the per-kernel "rim" tables used by the interpolation kernel for cells whose stencil leaves the
  face (reference ``Position._fatten_h`` ``eulerian.py:386`` + ``four_matrix_for_uv``
  ``topology.py:689``).

Directions are ``0,1,2,3 = up, down, left, right`` (reference ``topology.py:30``).
Status codes returned by the vectorised routines stand for the exceptions the reference raises:
``OK`` / ``EDGE`` (``IndexError``: no neighbouring face, or "no wall between a cell and itself")
/ ``ABNORMAL`` (``ValueError``: faces not connected in a normal way).
"""

from __future__ import annotations

import logging

import numpy as np

LLC_FACE_CONNECT = np.array(
    [
        # up, down, left, right ; 999 = nothing there
        [1, 999, 12, 3],
        [2, 0, 11, 4],
        [6, 1, 10, 5],
        [4, 999, 0, 9],
        [5, 3, 1, 8],
        [6, 4, 2, 7],
        [10, 5, 2, 7],
        [10, 5, 6, 8],
        [11, 4, 7, 9],
        [12, 3, 8, 999],
        [2, 7, 6, 11],
        [1, 8, 10, 12],
        [0, 9, 11, 999],
    ]
)
ASTE_FACE_CONNECT = np.array([[1, 999, 3, 999], [3, 999, 0, 2], [3, 999, 1, 999], [0, 2, 1, 999]])
DIRECTIONS = np.array([np.pi / 2, -np.pi / 2, np.pi, 0])
NO_FACE = 999

OK, EDGE, ABNORMAL = 0, 1, 2

_DY = np.array([1, -1, 0, 0])
_DX = np.array([0, 0, -1, 1])
# after entering a face that is rotated by 90 / 270 degrees the remaining moves are re-labelled
# (reference topology.py:475-482)
_REMAP = np.array([[0, 1, 2, 3], [2, 3, 1, 0], [0, 1, 2, 3], [3, 2, 0, 1]])  # index: quarter turns


def _quarter_turns(edge, new_edge):
    """rot = pi - DIR[edge] + DIR[new_edge] in units of 90 degrees (0..3)."""
    rot = (np.pi - DIRECTIONS[edge] + DIRECTIONS[new_edge]) % (2 * np.pi)
    q = int(np.round(rot / (np.pi / 2))) % 4
    return q


class Topology:
    """Remembers the structure of the grid and moves indices around on it."""

    def __init__(self, od=None, typ=None, *, h_shape=None, g_shape=None, itmax=0, izmax=0, lon_lr=None):
        if od is not None:
            try:
                h_shape = tuple(int(i) for i in od["XC"].shape)
            except KeyError:
                try:
                    h_shape = (int(od["lat"].shape[0]), int(od["lon"].shape[0]))
                except KeyError as exc:
                    raise KeyError(
                        "Either XC or lat/lon is needed to create the Topology object"
                    ) from exc
            if "XG" in od.variables:
                g_shape = tuple(int(i) for i in od["XG"].shape)
            try:
                itmax = len(od["time"]) - 1
            except (KeyError, TypeError):
                itmax = 0
            try:
                izmax = len(od["Z"]) - 1
            except (KeyError, TypeError):
                izmax = 0
        self.h_shape = tuple(h_shape)
        self.g_shape = tuple(g_shape) if g_shape is not None else self.h_shape
        self.itmax = itmax
        self.izmax = izmax
        self.face_connect = None
        if typ:
            self.typ = typ
            if len(self.h_shape) == 3:
                self.num_face, ny, nx = self.h_shape
            else:
                ny, nx = self.h_shape
            self.iymax, self.ixmax = ny - 1, nx - 1
        elif len(self.h_shape) == 3:
            self.num_face, ny, nx = self.h_shape
            self.iymax, self.ixmax = ny - 1, nx - 1
            if self.num_face == 13:
                self.typ = "LLC"
                self.face_connect = LLC_FACE_CONNECT
            elif self.num_face == 4:
                self.typ = "LLC"
                self.face_connect = ASTE_FACE_CONNECT
            else:
                raise NotImplementedError(
                    "Other complex topology is not implemented yet."
                    "If you want to work with such data, please contact the authors."
                )
        elif len(self.h_shape) == 2:
            ny, nx = self.h_shape
            self.iymax, self.ixmax = ny - 1, nx - 1
            if lon_lr is None:
                try:
                    lon_right = float(od["XC"][0, self.ixmax])
                    lon_left = float(od["XC"][0, 0])
                except KeyError:
                    lon_right = float(od["lon"][self.ixmax])
                    lon_left = float(od["lon"][0])
            else:
                lon_left, lon_right = lon_lr
            left_to_right = (lon_right - lon_left) % 360
            right_to_left = (lon_left - lon_right) % 360
            self.typ = "x_periodic" if left_to_right > 50 * right_to_left else "box"
        else:
            raise ValueError(f"unsupported horizontal shape {self.h_shape}")
        if self.typ == "LLC" and self.face_connect is None:
            self.face_connect = LLC_FACE_CONNECT if self.h_shape[0] == 13 else ASTE_FACE_CONNECT
        self.has_face = len(self.h_shape) == 3
        if self.typ == "LLC":
            self._build_face_tables()

    @classmethod
    def from_shape(cls, h_shape, g_shape=None, typ=None, lon_lr=(0.0, 10.0)):
        """Topology from the grid shape only (what the reference's shape-only tests need)."""
        return cls(None, typ, h_shape=h_shape, g_shape=g_shape, lon_lr=lon_lr)

    # ------------------------------------------------------------------------------------------
    # face tables
    # ------------------------------------------------------------------------------------------
    def _build_face_tables(self):
        fc = self.face_connect
        nf = fc.shape[0]
        self.nbr_face = np.full((nf, 4), -1, np.int32)
        self.nbr_edge = np.full((nf, 4), -1, np.int32)
        for f in range(nf):
            for e in range(4):
                g = int(fc[f, e])
                if g == NO_FACE:
                    continue
                back = np.where(fc[g] == f)[0]
                self.nbr_face[f, e] = g
                self.nbr_edge[f, e] = int(back[0])
        # relative orientation of every ordered pair of faces, allowing one face in between
        # (reference _llc_mutual_direction, topology.py:34, transitive=True)
        self._md_edge = np.full((nf, nf), -1, np.int32)
        self._md_new_edge = np.full((nf, nf), -1, np.int32)
        self._md_direct = np.zeros((nf, nf), bool)
        self._md_status = np.full((nf, nf), ABNORMAL, np.int32)
        for f in range(nf):
            for g in range(nf):
                if f == g:
                    continue
                res = self._mutual_direction_scalar(f, g, transitive=True)
                if res is None:
                    continue
                self._md_edge[f, g], self._md_new_edge[f, g], self._md_direct[f, g] = res
                self._md_status[f, g] = OK
        self._md_turns = np.zeros((nf, nf), np.int32)
        for f in range(nf):
            for g in range(nf):
                if self._md_status[f, g] == OK:
                    self._md_turns[f, g] = _quarter_turns(self._md_edge[f, g], self._md_new_edge[f, g])

    def _mutual_direction_scalar(self, face, neighbor_face, transitive=False):
        fc = self.face_connect
        to_face = np.where(fc[face] == neighbor_face)[0]
        to_neigh = np.where(fc[neighbor_face] == face)[0]
        if len(to_face) > 0 and len(to_neigh) > 0:
            return int(to_face[0]), int(to_neigh[0]), True
        if not transitive:
            return None
        common = -1
        for cand in fc[face]:
            if cand in fc[neighbor_face] and cand != NO_FACE:
                common = int(cand)
                break
        if common < 0:
            return None
        edge_1 = int(np.where(fc[face] == common)[0][0])
        new_edge_1 = int(np.where(fc[common] == face)[0][0])
        edge_2 = int(np.where(fc[common] == neighbor_face)[0][0])
        new_edge_2 = int(np.where(fc[neighbor_face] == common)[0][0])
        vert = (0, 1)
        if (edge_1 in vert) == (new_edge_1 in vert):
            return edge_2, new_edge_2, False
        if (edge_2 in vert) == (new_edge_2 in vert):
            return edge_1, new_edge_1, False
        return None

    # ------------------------------------------------------------------------------------------
    # reference-compatible scalar API
    # ------------------------------------------------------------------------------------------
    def _need_faces(self):
        if self.typ in ("x_periodic", "box"):
            raise ValueError(
                "It makes no sense to tinker with face_connection when there is only one face"
            )
        if self.typ != "LLC":
            raise NotImplementedError

    def get_the_other_edge(self, face, edge):
        """``(new_face, new_edge)``: what touches ``face`` along ``edge`` (ref. topology.py:340)."""
        self._need_faces()
        if self.nbr_face[face, edge] < 0:
            raise IndexError("Reaching the edge where the face is not connected to any other face")
        return int(self.nbr_face[face, edge]), int(self.nbr_edge[face, edge])

    def mutual_direction(self, face, new_face, transitive=False):
        """Relative orientation of two faces (ref. topology.py:370)."""
        self._need_faces()
        face, new_face = int(face), int(new_face)
        if self._md_status[face, new_face] == OK and (transitive or self._md_direct[face, new_face]):
            return int(self._md_edge[face, new_face]), int(self._md_new_edge[face, new_face])
        if transitive:
            raise ValueError("The two faces does not share common face, transitive did not help")
        raise ValueError("The two faces are not connected (transitive = False)")

    def ind_tend(self, ind, tend, cuvwg="C", **kwarg):
        """Move one index one step in direction ``tend`` (ref. topology.py:389)."""
        if -1 in ind:
            return tuple(-1 for _ in ind)
        arrs = tuple(np.array([int(i)]) for i in ind)
        tend = np.array([int(tend)])
        if self.typ == "LLC":
            if cuvwg == "C":
                out, status = self._cmove(arrs, tend)
            elif cuvwg == "U":
                out, _, status = self._umove(arrs, tend)
            elif cuvwg == "V":
                out, _, status = self._vmove(arrs, tend)
            elif cuvwg == "G":
                out, status = self._gmove(arrs, tend)
            else:
                raise ValueError("The type of grid point should be among C,U,V,G")
            self._raise(status[0], ind)
        elif self.typ in ("x_periodic", "box"):
            out, _ = self._cmove(arrs, tend)
        else:
            raise NotImplementedError
        return tuple(int(o[0]) for o in out)

    def ind_moves(self, ind, moves, **kwarg):
        """Walk from ``ind`` along ``moves`` (ref. topology.py:436)."""
        if self.check_illegal(tuple(int(i) for i in ind)):
            return tuple(-1 for _ in ind)
        if not set(moves).issubset({0, 1, 2, 3}):
            raise ValueError("Illegal move. Must be 0,1,2,3")
        arrs = tuple(np.array([int(i)]) for i in ind)
        out, status = self._walk(arrs, list(moves), kwarg.get("cuvwg", "C"))
        self._raise(status[0], ind)
        return tuple(int(o[0]) for o in out)

    @staticmethod
    def _raise(status, ind):
        if status == EDGE:
            raise IndexError(f"no neighbouring face / no wall for index {tuple(ind)}")
        if status == ABNORMAL:
            raise ValueError(f"The faces around index {tuple(ind)} are not connected in a normal way")

    def check_illegal(self, ind, cuvwg="C"):
        """True where an index is outside the grid (ref. topology.py:491)."""
        shape = self.h_shape if cuvwg == "C" else self.g_shape
        if isinstance(ind[0], (int, np.integer)):
            return any(not 0 <= int(z) <= shape[i] - 1 for i, z in enumerate(ind))
        bad = False
        for i, z in enumerate(ind):
            z = np.asarray(z)
            bad = np.logical_or(np.logical_or(z < 0, z > shape[i] - 1), bad)
        return bad

    def ind_tend_vec(self, inds, tend, cuvwg="C", swallow=None, return_status=False, **kwarg):
        """Move many indices, one step each (ref. topology.py:526).

        Like the reference, an ``IndexError`` at a point ("on the edge") keeps that point's old
        index and logs a warning; a ``ValueError`` propagates.
        """
        inds = [np.array(i, dtype=np.int64) for i in inds]
        tend = np.asarray(tend).astype(np.int64)
        if tend.ndim == 0:
            tend = np.full(inds[0].shape, int(tend))
        naive = [i.copy() for i in inds]
        naive[-2] = naive[-2] + _DY[tend]
        naive[-1] = naive[-1] + _DX[tend]
        illegal = np.asarray(self.check_illegal(tuple(naive), cuvwg=cuvwg))
        status = np.zeros(inds[0].shape, np.int32)
        if illegal.any():
            sub = tuple(i[illegal] for i in inds)
            has_neg = np.zeros(sub[0].shape, bool)
            for s in sub:
                has_neg |= s == -1
            if self.typ == "LLC":
                if cuvwg == "C":
                    out, st = self._cmove(sub, tend[illegal])
                elif cuvwg == "U":
                    out, _, st = self._umove(sub, tend[illegal])
                elif cuvwg == "V":
                    out, _, st = self._vmove(sub, tend[illegal])
                elif cuvwg == "G":
                    out, st = self._gmove(sub, tend[illegal])
                else:
                    raise ValueError("The type of grid point should be among C,U,V,G")
            else:
                out, st = self._cmove(sub, tend[illegal])
            if (st == ABNORMAL).any() and not swallow:
                bad = np.where(st == ABNORMAL)[0][0]
                self._raise(ABNORMAL, tuple(int(s[bad]) for s in sub))
            keep = st != OK
            if keep.any():
                logging.warning("Some points are on the edge")
            for k in range(len(inds)):
                val = np.where(keep, sub[k], out[k])
                val = np.where(has_neg, -1, val)
                naive[k][illegal] = val
            status[illegal] = st
        if return_status:
            return np.array(naive), status
        return np.array(naive)

    def get_uv_mask_from_face(self, faces):
        """Rotation coefficients of every kernel node relative to node 0 (ref. topology.py:665)."""
        self._need_faces()
        faces = np.asarray(faces).astype(int)
        out = self.four_matrix_for_uv(faces[None, :])
        return tuple(o[0] for o in out)

    def four_matrix_for_uv(self, fface):
        """(UfromUvel, UfromVvel, VfromUvel, VfromVvel) for an (N, M) array of faces (ref. :689)."""
        self._need_faces()
        fface = np.asarray(fface).astype(int)
        f0 = np.broadcast_to(fface[:, :1], fface.shape)
        same = fface == f0
        st = np.where(same, OK, self._md_status[f0, np.where(same, 0, fface)])
        if (st != OK).any():
            raise ValueError("The two faces does not share common face, transitive did not help")
        edge = self._md_edge[f0, fface]
        new_edge = self._md_new_edge[f0, fface]
        rot = np.pi - DIRECTIONS[edge] + DIRECTIONS[new_edge]
        cos, sin = np.cos(rot), np.sin(rot)
        ufu = np.where(same, 1.0, cos)
        ufv = np.where(same, 0.0, sin)
        vfu = np.where(same, 0.0, -sin)
        vfv = np.where(same, 1.0, cos)
        return ufu, ufv, vfu, vfv

    # ------------------------------------------------------------------------------------------
    # vectorised core
    # ------------------------------------------------------------------------------------------
    def _cmove(self, ind, tend):
        """One C-grid step for arrays. Returns ``(new_ind, status)``."""
        tend = np.asarray(tend)
        if self.typ == "LLC":
            f, iy, ix = (np.asarray(a).astype(np.int64) for a in ind)
            iymax, ixmax = self.iymax, self.ixmax
            ny_, nx_ = iy + _DY[tend], ix + _DX[tend]
            off = (ny_ < 0) | (ny_ > iymax) | (nx_ < 0) | (nx_ > ixmax)
            nf = np.where(off, self.nbr_face[f, tend], f)
            ne = self.nbr_edge[f, tend]
            status = np.where(off & (nf < 0), EDGE, OK).astype(np.int32)
            s = np.where(tend >= 2, iy, ix)  # coordinate along the edge being crossed
            flip = (tend % 2 == 0) == (ne % 2 == 0)  # both in {up,left} or both in {down,right}
            dest_is_x = ne <= 1  # arriving through an up/down edge: along-edge coordinate is ix
            s_new = np.where(flip, np.where(dest_is_x, ixmax, iymax) - s, s)
            cy = np.where(ne == 0, iymax, np.where(ne == 1, 0, s_new))
            cx = np.where(ne == 2, 0, np.where(ne == 3, ixmax, s_new))
            good = off & (nf >= 0)
            oy = np.where(good, cy, np.where(off, iy, ny_))
            ox = np.where(good, cx, np.where(off, ix, nx_))
            of = np.where(good, nf, f)
            return (of, oy, ox), status
        iy, ix = (np.asarray(a).astype(np.int64) for a in ind)
        ny_, nx_ = iy + _DY[tend], ix + _DX[tend]
        bad = (ny_ > self.iymax) | (ny_ < 0)
        if self.typ == "x_periodic":
            # NB reproduces the reference's wrap (topology.py:141-144): ix > ixmax -> ix - ixmax
            nx_ = np.where(nx_ > self.ixmax, nx_ - self.ixmax, nx_)
            nx_ = np.where(nx_ < 0, self.ixmax + nx_ + 1, nx_)
        elif self.typ == "box":
            bad |= (nx_ > self.ixmax) | (nx_ < 0)
        else:
            raise NotImplementedError
        ny_ = np.where(bad, -1, ny_)
        nx_ = np.where(bad, -1, nx_)
        return (ny_, nx_), np.zeros(iy.shape, np.int32)

    def _walk(self, ind, moves, cuvwg="C"):
        """Vectorised ``ind_moves``: the same list of moves for every starting index."""
        if self.typ != "LLC":
            status = np.zeros(np.asarray(ind[0]).shape, np.int32)
            cur = tuple(np.asarray(a).astype(np.int64) for a in ind)
            for m in moves:
                neg = cur[0] == -1
                nxt, _ = self._cmove(cur, np.full(cur[0].shape, m))
                cur = tuple(np.where(neg, -1, n) for n in nxt)
            return cur, status
        cur = tuple(np.asarray(a).astype(np.int64) for a in ind)
        n = cur[0].shape
        status = np.zeros(n, np.int32)
        face = cur[0].copy()
        pending = [np.full(n, m, np.int64) for m in moves]
        for k in range(len(pending)):
            mv = pending[k]
            live = (status == OK) & (cur[0] != -1)
            if cuvwg == "C":
                nxt, st = self._cmove(cur, mv)
            elif cuvwg == "U":
                nxt, _, st = self._umove(cur, mv)
            elif cuvwg == "V":
                nxt, _, st = self._vmove(cur, mv)
            elif cuvwg == "G":
                nxt, st = self._gmove(cur, mv)
            else:
                raise ValueError("The type of grid point should be among C,U,V,G")
            status = np.where(live, st, status)
            cur = tuple(np.where(live & (st == OK), a, b) for a, b in zip(nxt, cur))
            changed = live & (st == OK) & (cur[0] != face)
            if changed.any():
                fo = np.where(changed, face, 0)
                fn = np.where(changed, cur[0], 0)
                md_bad = changed & (self._md_status[fo, fn] != OK)
                status = np.where(md_bad, ABNORMAL, status)
                turns = np.where(changed & ~md_bad, self._md_turns[fo, fn], 0)
                for later in range(k + 1, len(pending)):
                    pending[later] = _REMAP[turns, pending[later]]
                face = np.where(changed, cur[0], face)
        return cur, status

    # reference-compatible scalar faces of the vectorised routines ---------------------------------
    def _find_wall_between(self, ind1, ind2):
        """``("U"|"V", index)`` of the wall between two adjacent cells (ref. topology.py:569)."""
        self._need_faces()
        a = tuple(np.array([int(i)]) for i in ind1)
        b = tuple(np.array([int(i)]) for i in ind2)
        out, is_v, st = self._wall_between_vec(a, b)
        if st[0] == EDGE:
            raise IndexError(f"there is no wall between a cell and itself,{ind1},{ind2}")
        if st[0] == ABNORMAL:
            raise ValueError(f"The two face connecting the indexes {ind1},{ind2} are not connected in a normal way")
        return ("V" if is_v[0] else "U"), tuple(int(o[0]) for o in out)

    def _ind_tend_U(self, ind, tend):
        out, is_v, st = self._umove(tuple(np.array([int(i)]) for i in ind), np.array([int(tend)]))
        self._raise(st[0], ind)
        return ("V" if is_v[0] else "U"), tuple(int(o[0]) for o in out)

    def _ind_tend_V(self, ind, tend):
        out, is_v, st = self._vmove(tuple(np.array([int(i)]) for i in ind), np.array([int(tend)]))
        self._raise(st[0], ind)
        return ("V" if is_v[0] else "U"), tuple(int(o[0]) for o in out)

    def _ind_tend_G(self, ind, tend):
        if tend not in (0, 1, 2, 3):
            raise ValueError(f"tend {tend} not supported")
        out, st = self._gmove(tuple(np.array([int(i)]) for i in ind), np.array([int(tend)]))
        self._raise(st[0], ind)
        return tuple(int(o[0]) for o in out)

    def _wall_between_vec(self, ind1, ind2):
        """Vectorised wall lookup between adjacent cells (ref. topology.py:569).

        Returns ``(index, is_v, status)``; ``is_v`` tells whether the wall is a V wall.
        """
        f1, y1, x1 = ind1
        f2, y2, x2 = ind2
        same = f1 == f2
        # same face: the wall is stored with ceil(mean) of the two indices (the larger one for
        # adjacent cells); U wall if that differs from "the other" cell in x, V wall if in y
        ry = (y1 + y2 + 1) // 2
        rx = (x1 + x2 + 1) // 2
        second_is_ret = (y2 == ry) & (x2 == rx)
        othery = np.where(second_is_ret, y1, y2)
        otherx = np.where(second_is_ret, x1, x2)
        same_is_u = rx > otherx
        same_is_v = ~same_is_u & (ry > othery)
        st_same = np.where(same_is_u | same_is_v, OK, EDGE)
        # different faces
        fa = np.where(same, 0, f1)
        fb = np.where(same, 0, f2)
        direct = self._md_direct[fa, fb] & ~same
        d12 = self._md_edge[fa, fb]
        d21 = self._md_new_edge[fa, fb]
        use2 = np.isin(d12, (0, 3))  # face 2 lies up/right of face 1 -> wall stored with cell 2
        use1 = ~use2 & np.isin(d21, (0, 3))
        key = np.where(use2, d21, d12)
        diff_is_v = key == 1
        diff_is_u = key == 2
        st_diff = np.where(direct & (use1 | use2) & (diff_is_u | diff_is_v), OK, ABNORMAL)
        pick2 = ~same & use2
        of = np.where(same, f1, np.where(pick2, f2, f1))
        oy_ = np.where(same, ry, np.where(pick2, y2, y1))
        ox_ = np.where(same, rx, np.where(pick2, x2, x1))
        is_v = np.where(same, same_is_v, diff_is_v)
        status = np.where(same, st_same, st_diff).astype(np.int32)
        return (of, oy_, ox_), is_v, status

    def _umove(self, ind, tend):
        """Move U-point indices (ref. ``_ind_tend_U`` topology.py:618). Returns (ind, is_v, status)."""
        first, st = self._cmove(ind, tend)
        stay = (first[0] == ind[0]) | (st != OK)
        is_v = np.zeros(first[0].shape, bool)
        if stay.all():
            return first, is_v, st
        go = ~stay
        sub = tuple(a[go] for a in ind)
        alter, st_a = self._walk_pair(sub, 2, tend[go])
        f_sub = tuple(a[go] for a in first)
        equal = (alter[0] == f_sub[0]) & (alter[1] == f_sub[1]) & (alter[2] == f_sub[2]) & (st_a == OK)
        wall, wv, st_w = self._wall_between_vec(f_sub, alter)
        st_w = np.where(st_a != OK, st_a, st_w)
        out = tuple(np.where(equal, a, w) for a, w in zip(alter, wall))
        out_v = np.where(equal, False, wv)
        out_st = np.where(equal, OK, st_w)
        res = tuple(a.copy() for a in first)
        for k in range(3):
            res[k][go] = out[k]
        is_v[go] = out_v
        st = st.copy()
        st[go] = out_st
        return res, is_v, st

    def _vmove(self, ind, tend):
        """Move V-point indices (ref. ``_ind_tend_V`` topology.py:635)."""
        first, st = self._cmove(ind, tend)
        stay = (first[0] == ind[0]) | (st != OK)
        is_v = np.ones(first[0].shape, bool)
        if stay.all():
            return first, is_v, st
        go = ~stay
        sub = tuple(a[go] for a in ind)
        alter, st_a = self._walk_pair(sub, 1, tend[go])
        f_sub = tuple(a[go] for a in first)
        wall, wv, st_w = self._wall_between_vec(f_sub, alter)
        st_w = np.where(st_a != OK, st_a, st_w)
        res = tuple(a.copy() for a in first)
        for k in range(3):
            res[k][go] = wall[k]
        is_v[go] = wv
        st = st.copy()
        st[go] = st_w
        return res, is_v, st

    def _gmove(self, ind, tend):
        """Move corner-point indices (ref. ``_ind_tend_G`` topology.py:649)."""
        tend = np.asarray(tend)
        vert = tend <= 1
        res = tuple(np.asarray(a).astype(np.int64).copy() for a in ind)
        st = np.zeros(res[0].shape, np.int32)
        if vert.any():
            out, _, s = self._umove(tuple(a[vert] for a in res), tend[vert])
            for k in range(3):
                res[k][vert] = out[k]
            st[vert] = s
        hor = ~vert
        if hor.any():
            src = tuple(np.asarray(a).astype(np.int64)[hor] for a in ind)
            out, _, s = self._vmove(src, tend[hor])
            for k in range(3):
                res[k][hor] = out[k]
            st[hor] = s
        return res, st

    def _walk_pair(self, ind, first_move, second_moves):
        """``ind_moves(ind, [first_move, second])`` with a per-element second move (C grid)."""
        second_moves = np.asarray(second_moves)
        out = tuple(np.asarray(a).astype(np.int64).copy() for a in ind)
        st = np.zeros(out[0].shape, np.int32)
        for m in np.unique(second_moves):
            sel = second_moves == m
            o, s = self._walk(tuple(a[sel] for a in ind), [first_move, int(m)], "C")
            for k in range(3):
                out[k][sel] = o[k]
            st[sel] = s
        return out, st

    # ------------------------------------------------------------------------------------------
    # stencil ("fatten") helpers shared by the host API and the device table builders
    # ------------------------------------------------------------------------------------------
    def fatten_node(self, ind, node, cuvwg="C"):
        """Index of kernel node ``(dx, dy)`` for every starting index (ref. eulerian.py:386-449).

        Returns ``(new_ind, status)``.  Naive offset where that stays on the (C or G shaped) grid,
        a topology walk elsewhere -- exactly like the reference's redo loop.
        """
        ind = tuple(np.asarray(a).astype(np.int64) for a in ind)
        dx, dy = int(node[0]), int(node[1])
        naive = [a.copy() for a in ind]
        naive[-2] = naive[-2] + dy
        naive[-1] = naive[-1] + dx
        status = np.zeros(ind[0].shape, np.int32)
        illegal = np.asarray(self.check_illegal(tuple(naive), cuvwg=cuvwg))
        if illegal.any():
            moves = [0] * dy if dy > 0 else [1] * (-dy)
            moves += [2] * (-dx) if dx < 0 else [3] * dx
            sub = tuple(a[illegal] for a in ind)
            out, st = self._walk(sub, moves, cuvwg)
            for k in range(len(ind)):
                naive[k][illegal] = np.where(st == OK, out[k], sub[k])
            status[illegal] = st
        return tuple(naive), status

    def wall_source(self, ind, which):
        """Index the advection kernel reads for the ``"right"`` / ``"top"`` wall of each cell."""
        node, cuvwg = ((1, 0), "U") if which == "right" else ((0, 1), "V")
        if not self.has_face:
            out, _ = self.fatten_node(ind, node, "G")
            return out
        out, _ = self.fatten_node(ind, node, cuvwg)
        return out

    def wall_source_missing(self, ind, which):
        node, cuvwg = ((1, 0), "U") if which == "right" else ((0, 1), "V")
        _, st = self.fatten_node(ind, node, cuvwg if self.has_face else "G")
        return st != OK

    def wall_source_is_other(self, ind, which):
        """True where the wall value is the *other* velocity component on the neighbouring face."""
        ind = tuple(np.asarray(a) for a in ind)
        if not self.has_face:
            return np.zeros(ind[0].shape, bool)
        src = self.wall_source(ind, which)
        f0, f1 = ind[0], src[0]
        same = f0 == f1
        turns = np.where(same, 0, self._md_turns[f0, np.where(same, 0, f1)])
        return turns % 2 == 1

    def corner_indices(self, ind):
        """Corner (G) indices LL, LR, UR, UL of every cell (ref. ``find_px_py`` utils.py:495)."""
        ind = tuple(np.asarray(a).astype(np.int64) for a in ind)
        n = ind[0].shape
        ind1 = tuple(self.ind_tend_vec(ind, np.full(n, 3), cuvwg="G", swallow=False))
        ind2 = tuple(self.ind_tend_vec(ind1, np.full(n, 0), cuvwg="G", swallow=False))
        ind3 = tuple(self.ind_tend_vec(ind, np.full(n, 0), cuvwg="G", swallow=False))
        return ind, ind1, ind2, ind3
