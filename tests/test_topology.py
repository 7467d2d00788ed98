"""Host Topology: known answers of the reference's own tests (tests/test_topology.py of the
reference, shape-only), the C oracle restatement, and -- when the reference checkout is present --
an exhaustive comparison over every edge cell."""

import itertools

import numpy as np
import pytest

from oracle.refload import reference_available
from seaduck_b200.topology import Topology


@pytest.fixture(scope="module")
def tp():
    return Topology.from_shape((13, 90, 90))


@pytest.fixture(scope="module")
def tp_aste():
    return Topology.from_shape((4, 270, 270))


# reference tests/test_topology.py:73-89
@pytest.mark.parametrize(
    "ind,tend,result",
    [
        ((1, 45, 45), 0, (1, 46, 45)),
        ((1, 45, 45), 1, (1, 44, 45)),
        ((1, 45, 45), 2, (1, 45, 44)),
        ((1, 45, 45), 3, (1, 45, 46)),
        ((5, 89, 89), 3, (7, 0, 0)),
        ((5, 89, 89), 0, (6, 0, 89)),
        ((6, 89, 89), 0, (10, 0, 0)),
        ((6, 0, 0), 2, (2, 89, 89)),
    ],
)
def test_llc_ind_tend(tp, ind, tend, result):
    assert tp.ind_tend(ind, tend) == result


# reference tests/test_topology.py:92-108
@pytest.mark.parametrize(
    "ind,tend,result",
    [
        ((0, 269, 269), 0, (1, 0, 0)),
        ((1, 0, 0), 2, (0, 269, 269)),
        ((1, 269, 0), 0, (3, 269, 0)),
        ((1, 269, 269), 3, (2, 269, 0)),
    ],
)
def test_llc_ind_tend_aste(tp_aste, ind, tend, result):
    assert tp_aste.ind_tend(ind, tend) == result


@pytest.mark.parametrize("face,edge", [(0, 1), (3, 1), (9, 3), (12, 3)])
def test_antarctica_error(tp, face, edge):
    with pytest.raises(IndexError):
        tp.get_the_other_edge(face, edge)


# reference tests/test_topology.py:202-230
@pytest.mark.parametrize(
    "ind,tend,ans",
    [((1, 45, 45), 0, (1, 46, 45)), ((1, 45, 45), 2, (1, 45, 44)), ((1, 0, 0), 2, (12, 89, 0)),
     ((1, 0, 0), 1, (0, 89, 0)), ((4, 45, 89), 3, (8, 0, 45))],
)
def test_ind_tend_v(tp, ind, tend, ans):
    assert tp.ind_tend(ind, tend, cuvwg="V") == ans


@pytest.mark.parametrize(
    "ans,tend,ind",
    [((1, 45, 45), 1, (1, 46, 45)), ((1, 45, 44), 2, (1, 45, 45)), ((1, 0, 0), 0, (12, 89, 0)), ((4, 45, 89), 1, (8, 0, 45))],
)
def test_ind_tend_u(tp, ind, tend, ans):
    assert tp.ind_tend(ind, tend, cuvwg="U") == ans


def test_wall_between(tp):
    assert tp._find_wall_between((11, 0, 45), (8, 89, 45)) == ("V", (11, 0, 45))
    assert tp._find_wall_between((8, 45, 0), (7, 45, 89)) == ("U", (8, 45, 0))
    with pytest.raises(IndexError):
        tp._find_wall_between((1, 14, 14), (1, 14, 14))


@pytest.mark.parametrize("ind, tend, exp", [((10, 89, 45), 0, ("U", (2, 44, 0))), ((7, 0, 45), 1, ("U", (5, 44, 89)))])
def test_ind_tend_v_detailed(tp, ind, tend, exp):
    assert tp._ind_tend_V(ind, tend) == exp


@pytest.mark.parametrize("ind, tend, exp", [((2, 44, 0), 2, ("V", (10, 89, 45))), ((5, 44, 89), 3, ("V", (7, 0, 45)))])
def test_ind_tend_u_detailed(tp, ind, tend, exp):
    assert tp._ind_tend_U(ind, tend) == exp


@pytest.mark.parametrize("face,nface,rot", [(4, 9, True), (3, 8, True), (9, 4, True), (8, 3, True), (0, 4, False), (1, 3, False)])
def test_transitive_mutual_direction(tp, face, nface, rot):
    e1, e2 = tp.mutual_direction(face, nface, transitive=True)
    assert (e1 // 2 + e2 // 2) % 2 == rot


mundane = np.array([[[1.0, 1.0]], [[0.0, 0.0]], [[0.0, -0.0]], [[1.0, 1.0]]])


@pytest.mark.parametrize("fface,cis", [([[1, 1]], True), ([[1, 2]], True), ([[10, 2]], False), ([[6, 10]], False)])
def test_4_matrix(tp, fface, cis):
    ans = np.array(tp.four_matrix_for_uv(np.array(fface)))
    assert np.allclose(ans, mundane) == cis


def test_other_errors(tp):
    with pytest.raises(ValueError):
        tp.ind_tend((1, 45, 45), 0, cuvwg="other")
    with pytest.raises(ValueError):
        tp.ind_moves((1, 45, 45), ["left", "left"])
    assert tp.ind_moves((1, -1, 89), [0, 0]) == (-1, -1, -1)


def test_box_and_xperiodic():
    box = Topology.from_shape((10, 10), typ="box")
    assert box.ind_tend((9, 9), 0) == (-1, -1)
    assert box.ind_tend((9, 9), 3) == (-1, -1)
    assert box.ind_tend((-1, -1), 0) == (-1, -1)
    xp = Topology.from_shape((10, 10), typ="x_periodic")
    assert xp.ind_tend((3, 9), 3) == (3, 1)  # sic: the reference wraps to column 1 (topology.py:141)
    assert xp.ind_tend((3, 0), 2) == (3, 9)
    assert xp.ind_tend((9, 3), 0) == (-1, -1)
    auto = Topology.from_shape((10, 360), lon_lr=(-179.5, 179.5))
    assert auto.typ == "x_periodic"
    assert Topology.from_shape((10, 20), lon_lr=(0.0, 19.0)).typ == "box"


def _edge_cells(nf, n):
    return [(f, iy, ix) for f in range(nf) for iy in range(n) for ix in range(n) if min(iy, ix) <= 1 or max(iy, ix) >= n - 2]


@pytest.mark.parametrize("nf", [13, 4])
def test_oracle_topology_agrees_with_product(nf, oracle_lib):
    """Two independent implementations (vectorised tables vs scalar C restatement) on every edge cell."""
    n = 6

    class _Od:
        readiness = {"h": "oceanparcel", "Z": False, "Zl": False, "time": False}

    tp = Topology.from_shape((nf, n, n))
    od = _Od()
    od.tp = tp
    od.XC = od.YC = od.XG = od.YG = np.zeros((nf, n, n), np.float32)
    od.dX = od.dY = od.CS = od.SN = None
    g = oracle_lib.OracleGrid(od)
    errs = {0: None, 1: IndexError, 2: ValueError}
    for cu, c, t in itertools.product("CUVG", _edge_cells(nf, n), range(4)):
        err, out = g.ind_tend(c, t, cu)
        if errs[err] is None:
            assert tp.ind_tend(c, t, cuvwg=cu) == out, (cu, c, t)
        else:
            with pytest.raises(errs[err]):
                tp.ind_tend(c, t, cuvwg=cu)
    for cu, c in itertools.product("CUVG", _edge_cells(nf, n)):
        for mv in ([0, 0], [3, 3], [2, 1], [1, 3], [0, 2], [3, 0], [2, 2]):
            err, out = g.ind_moves(c, mv, cu)
            if errs[err] is None:
                assert tp.ind_moves(c, mv, cuvwg=cu) == out, (cu, c, mv)
            else:
                with pytest.raises(errs[err]):
                    tp.ind_moves(c, mv, cuvwg=cu)


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("nf", [13, 4])
def test_product_topology_agrees_with_reference(nf):
    from oracle.refload import load_reference

    load_reference()
    import xarray as xr
    from seaduck.topology import Topology as RefTopology

    n = 6
    ds = xr.Dataset(coords=dict(XC=(["face", "Y", "X"], np.zeros((nf, n, n))), XG=(["face", "Yp1", "Xp1"], np.zeros((nf, n, n)))))
    ref, mine = RefTopology(ds), Topology(ds)

    def call(obj, name, *args, **kw):
        try:
            return getattr(obj, name)(*args, **kw)
        except IndexError:
            return "IndexError"
        except ValueError:
            return "ValueError"

    for cu, c, t in itertools.product("CUVG", _edge_cells(nf, n), range(4)):
        assert call(ref, "ind_tend", c, t, cuvwg=cu) == call(mine, "ind_tend", c, t, cuvwg=cu), (cu, c, t)
    nodes = [(dx, dy) for dx in range(-2, 3) for dy in range(-2, 3)]
    from seaduck.kernel_weight import _translate_to_tendency

    for cu, c, nd in itertools.product("CUVG", _edge_cells(nf, n), nodes):
        mv = _translate_to_tendency(np.array(nd))
        assert call(ref, "ind_moves", c, mv, cuvwg=cu) == call(mine, "ind_moves", c, mv, cuvwg=cu), (cu, c, nd)
    cells = np.array(_edge_cells(nf, n)).T
    for t in range(4):
        for cu in "CG":
            a = ref.ind_tend_vec(tuple(cells), np.full(cells.shape[1], t), cuvwg=cu)
            b = mine.ind_tend_vec(tuple(cells), np.full(cells.shape[1], t), cuvwg=cu)
            np.testing.assert_array_equal(a, b)
