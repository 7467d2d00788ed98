"""KnW host mirror: the reference's own properties (tests/test_weight.py of the reference) and,
when the reference checkout is present, value-by-value agreement with the reference's weights."""

import numpy as np
import pytest

import seaduck_b200.kernel_weight as kw
from oracle.refload import reference_available
from seaduck_b200.lagrangian import uknw

K9 = np.array([[0, 0], [0, 1], [0, 2], [0, -1], [0, -2], [-1, 0], [-2, 0], [1, 0], [2, 0]])
K5 = np.array([[0, 0], [0, 1], [0, -1], [-1, 0], [1, 0]])
K4 = np.array([[0, 0], [0, -1], [-1, 0], [1, 0]])
K3X = np.array([[0, 0], [-1, 0], [1, 0]])
K3Y = np.array([[0, 0], [0, 1], [0, -1]])
K1 = np.array([[0, 0]])
SQ = np.array([[0, 0], [0, 1], [1, 0], [1, 1]])
POINTS = [(np.array([0.0]), np.array([0.0])), (np.array([0.5]), np.array([0.08]))]


@pytest.fixture
def aknw():
    return kw.KnW(inheritance=[[0, 1, 2, 3, 4, 5, 6, 7, 8]], hkernel="interp", vkernel="linear", tkernel="dt")


def test_same_size_and_equal(aknw):
    assert not aknw.same_size(uknw)
    assert not aknw == 1
    assert kw.KnW() == kw.KnW()
    assert hash(kw.KnW()) == hash(kw.KnW())


def test_default_knw_has_three_levels():
    k = kw.KnW()
    assert [len(h) for h in k.inheritance] == [9, 5, 1]


@pytest.mark.parametrize(
    "masked,ans",
    [
        (np.array([[1, 1, 1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 1, 1, 1, 1], [1, 1, 0, 1, 0, 1, 1, 1, 1],
                   [1, 0, 1, 1, 0, 1, 1, 1, 1], [0, 1, 1, 1, 0, 1, 1, 1, 1]]), [[0], [1], [2], [3]]),
        (np.array([[0, 1, 1, 1, 0, 1, 1, 1, 1]]), [[] for i in range(4)]),
    ],
)
def test_cascade(masked, ans):
    assert kw.find_which_points_for_each_kernel(masked) == ans


@pytest.mark.parametrize("rx,ry", POINTS)
@pytest.mark.parametrize("pk,expected", [([[], [], [], []], np.nan), ([[0], [], [], []], 1), ([[], [], [0], []], 1)])
def test_cascade_weight(rx, ry, pk, expected):
    np.testing.assert_allclose(np.sum(kw.get_weight_cascade(rx, ry, pk)), expected)


@pytest.mark.parametrize("rx,ry", POINTS)
@pytest.mark.parametrize("kernel", [K9, K5, K4, K3X, K3Y, K1, SQ])
def test_interp_weights_sum_to_one(kernel, rx, ry):
    weight = kw.get_func(kernel)(rx, ry)
    assert weight.shape == (1, len(kernel))
    np.testing.assert_allclose(np.sum(weight), 1)


@pytest.mark.parametrize("rx,ry", POINTS)
@pytest.mark.parametrize("kernel_type", ["dx", "dy"])
@pytest.mark.parametrize("horder", [0, 1])
def test_square_kernel(kernel_type, horder, rx, ry):
    weight = kw.get_func(SQ, hkernel=kernel_type, h_order=horder)(rx, ry)
    np.testing.assert_allclose(np.sum(weight), 1 if horder == 0 else 0, atol=1e-14)


@pytest.mark.parametrize("rx,ry", POINTS)
@pytest.mark.parametrize("kernel,hk", [(K9, "dx"), (K5, "dx"), (K4, "dx"), (K3X, "dx"), (K9, "dy"), (K5, "dy"), (K3Y, "dy")])
@pytest.mark.parametrize("order", [1, 2])
def test_derivative_weights_sum_to_zero(kernel, hk, rx, ry, order):
    weight = kw.get_func(kernel, hkernel=hk, h_order=order)(rx, ry)
    assert np.allclose(0, np.sum(weight), atol=1e-14)


@pytest.mark.parametrize("kernel_type", ["dx", "dy"])
@pytest.mark.parametrize("kernel,order", [(K9, 5), (K5, 3), (K1, 1), (SQ, 2)])
def test_order_too_high_error(kernel, order, kernel_type):
    with pytest.raises(Exception):
        kw.kernel_weight(kernel, kernel_type=kernel_type, order=order)


def test_dt_and_bottom_scheme(aknw):
    weight = aknw.get_weight(np.array([0.618]), np.array([0.0618]))
    assert weight.shape[0] == 1
    assert np.allclose(0, np.sum(weight))


@pytest.mark.parametrize("which", ["dx", "dy"])
def test_maxorder_dxdy_square(which):
    kkk = np.array([[0, 0], [0, 1], [1, 0], [1, 1], [1, 2], [0, 2], [2, 2], [2, 0], [2, 1]])
    func = kw.kernel_weight_s(kkk, 2 if which == "dx" else 0, 2 if which == "dy" else 0)
    weight = func(np.array([0.0]), np.array([0.0]))
    assert weight.shape == (1, len(kkk))
    assert np.allclose(0, np.sum(weight), atol=1e-14)


@pytest.mark.parametrize("hkernel", ["dx", "dy"])
def test_auto_inheritance(hkernel):
    inheritance = kw._auto_inheritance(SQ, hkernel=hkernel)
    assert len(inheritance) > 1 and isinstance(inheritance[0], list)


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference checkout not present (GPU box)")
def test_weights_agree_with_reference():
    from oracle.refload import load_reference

    load_reference()
    import seaduck.kernel_weight as rkw

    rng = np.random.RandomState(0)
    rx, ry = rng.uniform(-0.5, 0.5, 50), rng.uniform(-0.5, 0.5, 50)
    rz, rt = rng.uniform(0, 1, 50), rng.uniform(0, 1, 50)
    configs = [
        dict(), dict(kernel=K5), dict(kernel=SQ, vkernel="linear"), dict(kernel=K9, hkernel="dx", h_order=1),
        dict(kernel=K9, hkernel="dx", h_order=1, inheritance=None), dict(kernel=K9, hkernel="dy", h_order=2, inheritance=None),
        dict(kernel=K5, hkernel="dx", h_order=2, tkernel="linear", inheritance=[[0, 1, 2, 3, 4], [0, 3, 4]]),
        dict(kernel=SQ, hkernel="dx", h_order=1, vkernel="dz", inheritance=None), dict(kernel=K3X, tkernel="dt"),
        dict(kernel=np.array([[0, 0], [1, 0], [0, 1]]), inheritance=[[0, 1]], ignore_mask=True),
        dict(kernel=np.array([[0, 0], [1, 0], [0, 1]]), inheritance=[[0, 2]], hkernel="dy", h_order=1, ignore_mask=True),
    ]
    built = 0
    for cfg in configs:
        try:
            ref = rkw.KnW(**cfg)
        except ValueError:
            with pytest.raises(ValueError):  # e.g. the 1-node level of an auto inheritance cannot do d/dx
                kw.KnW(**cfg)
            continue
        mine = kw.KnW(**cfg)
        built += 1
        np.testing.assert_array_equal(mine.kernel, ref.kernel)
        assert mine.inheritance == [list(map(int, h)) for h in ref.inheritance]
        np.testing.assert_allclose(mine.get_weight(rx, ry, rz, rt), ref.get_weight(rx, ry, rz, rt), rtol=1e-12, atol=1e-13)
        # masked cascade: random wet/dry patterns for every vertical / temporal node
        m = len(mine.kernel)
        masked = (rng.uniform(size=(50, m, mine.nz, mine.nt)) > 0.25).astype(float)
        a = mine.get_weight(rx, ry, rz, rt, pk4d=kw._find_pk_4d(masked, mine.inheritance), bottom_scheme="no flux")
        b = ref.get_weight(rx, ry, rz, rt, pk4d=rkw._find_pk_4d(masked, ref.inheritance), bottom_scheme="no flux")
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13)
    assert built >= 8
