"""GPU: the reference's own acceptance tests for the particle loop (tests/test_streamfunction.py:100-143)
at BASELINE.json configs[0] size: in a non-divergent flow built from a corner stream function the
analytic trajectories conserve the stream function (atol 1e-7, as in the reference), and the whole run
agrees with the CPU oracle."""

import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _gknw(sd, vkernel):
    return sd.KnW(np.array([[0, 0], [0, 1], [1, 0], [1, 1]]), vkernel=vkernel, tkernel="nearest")


def test_horizontal_streamfunction_conservation_c1(oracle_lib):
    """configs[0]: double gyre on a 100 x 100 C-grid, 10^4 particles, Particle.to_list_of_time."""
    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    ds = synthetic.gyre(100, 100)
    n = 10_000
    rng = np.random.RandomState(2023)
    x = rng.random_sample(n) * 2 - 1
    y = rng.random_sample(n) * 2 - 1
    od = sd.OceData(ds)
    pt = sd.Particle(x=x, y=y, z=None, t=np.zeros(n), data=od, wname=None, transport=True)
    op = oracle_lib.oracle_particle_from_od(od, x, y, None, np.zeros(n), wname=None, transport=True, nthreads=8)
    knw = _gknw(sd, "nearest")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        before = pt.interpolate("streamfunc", knw)
        stops, snaps = pt.to_list_of_time(normal_stops=np.linspace(0, 1500, 15), update_stops=[])
        after = snaps[-1].interpolate("streamfunc", knw)
    assert len(snaps) == 15
    assert np.allclose(after, before, atol=1e-7), np.abs(after - before).max()
    total = 0
    for ts in stops:
        total += op.to_next_stop(float(ts))
    assert pt.crossings == total and total > 3e5
    np.testing.assert_array_equal(pt.iy, op.iy)
    np.testing.assert_array_equal(pt.ix, op.ix)
    cases.assert_close("lon", pt.lon, op.lon)
    cases.assert_close("lat", pt.lat, op.lat)


def test_vertical_streamfunction_conservation():
    """x-z overturning cell (reference test_vertical_streamfunc_conservation), backwards in time."""
    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    ds = synthetic.overturning_cell(100, 50)
    n = 2000
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, n)
    z = rng.uniform(-0.1, -1000, n)
    od = sd.OceData(ds)
    pt = sd.Particle(x=x, y=np.zeros(n), z=z, t=np.zeros(n), data=od, transport=True)
    knw = _gknw(sd, "linear")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        before = pt.interpolate("streamfunc", knw)
        pt.to_list_of_time(normal_stops=np.linspace(0, -8000, 40), update_stops=[])
        after = pt.interpolate("streamfunc", knw)
    assert np.allclose(after, before, atol=1e-7), np.nanmax(np.abs(after - before))
