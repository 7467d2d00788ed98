"""Host side of the Eulerian-budget stencils: the halo maps that `sdk_gather_plane` applies, checked
against what the unmodified reference produced (tests/golden/eulerian_budget.npz, made by
oracle/make_golden_eulerian_budget.py).  No GPU needed: a map is applied with numpy fancy indexing."""

import os

import numpy as np
import pytest

from seaduck_b200 import eulerian_budget as eb
from seaduck_b200.topology import Topology

GOLD = os.path.join(os.path.dirname(__file__), "golden", "eulerian_budget.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def apply_map(m, a, plane_ndim):
    flat = a.reshape(a.shape[: a.ndim - plane_ndim] + (-1,))
    idx = m.reshape(-1)
    out = np.where(idx >= 0, flat[..., np.maximum(idx, 0)], 0.0)
    return out.reshape(a.shape[: a.ndim - plane_ndim] + m.shape)


def llc_topology(n):
    return Topology.from_shape((13, n, n), typ="LLC")


@pytest.mark.parametrize("face", range(13))
def test_withface_maps_match_reference(gold, face):
    s = gold["s"]
    tp = llc_topology(s.shape[-1])
    for axis in "xy":
        got = apply_map(eb._withface_map(tp, face, axis, 2, 1), s, 3)
        np.testing.assert_array_equal(got, gold[f"{axis}buf_{face}"])


def test_wide_margins(gold):
    s = gold["s"]
    tp = llc_topology(s.shape[-1])
    np.testing.assert_array_equal(apply_map(eb._withface_map(tp, 5, "x", 3, 2), s, 3), gold["xbuf_wide_5"])
    np.testing.assert_array_equal(apply_map(eb._withface_map(tp, 10, "y", 1, 3), s, 3), gold["ybuf_wide_10"])


def test_maps_are_cached_on_the_topology(gold):
    tp = llc_topology(8)
    assert eb._withface_map(tp, 3, "x", 2, 1) is eb._withface_map(tp, 3, "x", 2, 1)


def test_every_halo_cell_is_a_neighbour():
    """Size-independent property: next to the interior, a halo cell is the `ind_tend` neighbour of the edge cell."""
    n = 12
    tp = llc_topology(n)
    for face in range(13):
        m = eb._withface_map(tp, face, "x", 1, 1)
        for iy in (0, 5, n - 1):
            for col, start, move in ((0, 0, 2), (n + 1, n - 1, 3)):
                try:
                    f, y, x = tp.ind_tend((face, iy, start), move)
                except IndexError:
                    assert m[iy, col] == -1
                    continue
                assert m[iy, col] == (f * n + y) * n + x
        m = eb._withface_map(tp, face, "y", 1, 1)
        for ix in (0, 4, n - 1):
            for row, start, move in ((0, 0, 1), (n + 1, n - 1, 0)):
                try:
                    f, y, x = tp.ind_tend((face, start, ix), move)
                except IndexError:
                    assert m[row, ix] == -1
                    continue
                assert m[row, ix] == (f * n + y) * n + x


def test_periodic_and_nearest_maps(gold):
    sp = gold["sp"]
    np.testing.assert_array_equal(apply_map(eb._wrap_map(sp.shape[-1:], 0, 2, 2, "wrap"), sp, 1), gold["xper"])
    np.testing.assert_array_equal(apply_map(eb._wrap_map(sp.shape[-2:], 0, 1, 3, "wrap"), sp, 2), gold["yper"])
    np.testing.assert_array_equal(apply_map(eb._wrap_map(sp.shape[-3:], 0, 2, 1, "edge"), sp, 3), gold["znear"])


def test_zero_margin_is_refused():
    with pytest.raises(ValueError):
        eb._withface_map(llc_topology(8), 0, "x", 2, 0)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="GPU only"):
        eb.buffer_x_periodic(np.zeros((3, 4)), 1, 1)
