"""Host helpers that keep the reference's names (reference ``seaduck/utils.py:206, 688, 704``)."""

import numpy as np
import pytest

from seaduck_b200 import utils


def test_convert_time_known_answers():
    assert utils.convert_time("1970-01-02") == 86400.0
    assert utils.convert_time(np.datetime64("1992-01-01")) == 694224000.0
    assert utils.convert_time("1992-02-01T12:00") == 694224000.0 + 31.5 * 86400
    with pytest.raises(ValueError, match="Unsupported time format"):
        utils.convert_time(694224000)


def test_easy_3d_cube_order_is_level_row_column():
    # the reference builds the cube with nested meshgrids: longitude runs fastest, then latitude, then depth
    x, y, z, t = utils.easy_3d_cube((-20.0, 30.0, 3), (10.0, 60.0, 2), (-5.0, -500.0, 4), "1992-02-01")
    assert x.shape == y.shape == z.shape == t.shape == (24,)
    assert x[:3].tolist() == [-20.0, 5.0, 30.0] and np.array_equal(x[:6], x[6:12])
    assert y[:6].tolist() == [10.0, 10.0, 10.0, 60.0, 60.0, 60.0]
    assert z[:6].tolist() == [-5.0] * 6 and z[-1] == -500.0 and np.array_equal(np.unique(z), np.linspace(-500.0, -5.0, 4))
    assert np.all(t == utils.convert_time("1992-02-01")) and t.dtype == np.float64


def test_easy_3d_cube_matches_nested_meshgrids():
    lon, lat, dep = (-180.0, 180.0, 13), (-80.0, 80.0, 5), (-10.0, -3000.0, 7)
    x, y, z, _ = utils.easy_3d_cube(lon, lat, dep, "1993-01-01")
    xs, ys, zs = (np.linspace(*a) for a in (lon, lat, dep))
    zz, yy, xx = np.meshgrid(zs, ys, xs, indexing="ij")
    assert np.array_equal(x, xx.ravel()) and np.array_equal(y, yy.ravel()) and np.array_equal(z, zz.ravel())


def test_local_to_latlon_and_general_len():
    uu, vv = utils.local_to_latlon(np.array([1.0, 0.0]), np.array([0.0, 2.0]), np.array([0.0, 0.6]), np.array([1.0, 0.8]))
    assert np.allclose(uu, [0.0, -1.6]) and np.allclose(vv, [1.0, 1.2])
    assert utils._general_len(3.0) == 1 and utils._general_len([1, 2, 3]) == 3 and utils._general_len(np.zeros(5)) == 5
