"""CPU: the oracle against the known answers of the reference's own unit tests.

* the analytic in-cell step -- ``tests/test_particle.py:75-209`` of the reference (issue-55 regressions,
  under-flowing ``u`` / ``du``, closed walls, impossible inflow, both directions of time), with the outputs
  of the reference's ``_time2wall`` / ``_which_early`` / ``_increment`` stored by
  ``oracle/make_golden_known_answers.py`` in ``tests/golden/analytic_step.npz``;
* the 1-D index / relative-coordinate lookups -- the tables of ``tests/test_utils.py:108-161``.
"""

import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "analytic_step.npz")


@pytest.fixture(scope="module")
def analytic(oracle_lib):
    L = oracle_lib.lib()
    d3 = C.c_double * 3
    L.ora_analytic.restype = None
    L.ora_analytic.argtypes = [d3, d3, d3, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double), d3, d3]

    def step(x0, u, du, tf):
        tend, tev = C.c_int(-1), C.c_double(0.0)
        nx, nu = d3(), d3()
        L.ora_analytic(d3(*x0), d3(*u), d3(*du), float(tf), C.byref(tend), C.byref(tev), nx, nu)
        return tend.value, tev.value, np.array(nx[:]), np.array(nu[:])

    return step


def test_analytic_step_matches_reference(analytic):
    g = np.load(GOLD)
    n = len(g["tf"])
    assert n >= 140 and set(np.unique(g["tend"])) == set(range(7))  # every wall and the stop are in the sample
    for i in range(n):
        tend, tev, nx, nu = analytic(g["x0"][i], g["u"][i], g["du"][i], g["tf"][i])
        label = f"{g['name'][i]} tf={g['tf'][i]:g}"
        assert tend == g["tend"][i], label
        # libm vs numpy log / exp may differ in the last bit: 1e-12 relative (north star: 1e-9)
        np.testing.assert_allclose(tev, g["t_event"][i], rtol=1e-12, atol=0, err_msg=label)
        np.testing.assert_allclose(nx, g["newx"][i], rtol=1e-12, atol=1e-15, err_msg=label)
        np.testing.assert_allclose(nu, g["newu"][i], rtol=1e-12, atol=1e-25, err_msg=label)


def test_analytic_step_stays_in_the_cell(analytic):
    """What the reference asserts on these vectors (test_particle.py:19-41): with no stop in reach the particle ends
    on a wall (never `tend == 6` for the regression cases) and not beyond it."""
    g = np.load(GOLD)
    far = np.abs(g["tf"]) > 1e79
    for i in np.where(far)[0]:
        tend, _, nx, _ = analytic(g["x0"][i], g["u"][i], g["du"][i], g["tf"][i])
        name = str(g["name"][i])
        if tend != 6:
            assert (np.abs(nx) <= 0.5 + 1e-4).all(), name
        if g["tf"][i] > 0 and name.startswith(("issue55", "corner", "underflow")):
            assert tend != 6, name


ascend = np.array([1, 11, 111, 1111])
dscend = -ascend
darray = np.array([1, 10, 100, 1000])


@pytest.mark.parametrize(
    "value,array,ascending,above,peri,ans",
    [  # reference tests/test_utils.py:110-118
        (10, ascend, 1, True, None, 0),
        (10, ascend, 1, False, None, 1),
        (0, ascend, 1, False, None, 0),
        (-2, dscend, -1, True, None, 1),
        (-2, dscend, -1, False, None, 0),
        (0, ascend, 1, False, 110.5, 2),
        (0, ascend, 1, False, 1111.5, 3),
    ],
)
def test_find_ind_known_answers(oracle_lib, value, array, ascending, above, peri, ans):
    ix, _, _, bx = oracle_lib.find_rel(np.array([float(value)]), array, None, ascending, above, peri)
    assert ix[0] == ans and bx[0] == array[ans]
    assert np.issubdtype(ix.dtype, np.integer)


@pytest.mark.parametrize("value,array,ascending", [(0.5, ascend, 1), (-2000, dscend, -1)])  # test_utils.py:129
def test_find_ind_out_of_bound(oracle_lib, value, array, ascending):
    with pytest.raises(ValueError):
        oracle_lib.find_rel(np.array([float(value)]), array, None, ascending, True)


@pytest.mark.parametrize(
    "value,array,darray,ascending,above,peri,dx_right,ans",
    [  # reference tests/test_utils.py:139-151
        (10, ascend, darray, 1, True, None, False, 0.9),
        (10, ascend, 10 * darray, 1, True, None, True, 0.9),
        (10, ascend, None, 1, True, None, True, 0.9),
        (10, ascend, darray, 1, False, None, False, -0.1),
        (0, ascend, darray, 1, False, None, False, -1),
        (-2, dscend, darray, -1, True, None, False, 0.9),
        (-2, dscend, darray, -1, False, None, False, -0.1),
        (-2, dscend, 10 * darray, -1, False, None, True, -0.1),
        (0, dscend, darray, -1, True, None, False, 1),
        (0, ascend, darray, 1, False, 110.5, False, -0.5 / 100),
        (0, ascend, 10 * darray, 1, True, 1111.5, True, 0.5 / 10000),
    ],
)
def test_find_rel_known_answers(oracle_lib, value, array, darray, ascending, above, peri, dx_right, ans):
    ixs, rxs, dxs, bxs = oracle_lib.find_rel(np.array([float(value)]), array, darray, ascending, above, peri, dx_right)
    assert rxs[0] == ans
    assert np.issubdtype(ixs.dtype, np.integer)


def test_to_180_known_answers(oracle_lib):
    """reference tests/test_utils.py:26-31"""
    assert oracle_lib.to_180(365) == 5
    assert abs(oracle_lib.to_180(-1e-20)) < 1e-10
    # the wrap the particle loop relies on: a tiny negative difference rounds to 360 in the first modulo and
    # the second one brings it to zero (utils.py:199)
    assert oracle_lib.to_180(np.array([-1e-20, 180.0, -180.0, 540.0]))[0] == 0.0
    np.testing.assert_array_equal(oracle_lib.to_180(np.array([180.0, -180.0, 540.0, 190.0])), [-180.0, -180.0, -180.0, -170.0])


@pytest.mark.parametrize("lat,lon", [(0, 0), (45, 45), (50, 180)])  # reference tests/test_utils.py:41-51
def test_spherical2cartesian_radius(oracle_lib, lat, lon):
    x, y, z = oracle_lib.spherical2cartesian(lon, lat)
    assert np.isclose(x**2 + y**2 + z**2, 6371.0**2)
