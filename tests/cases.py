"""Seeded workloads shared by the golden generator, the oracle tests and the GPU parity tests."""

import numpy as np

from seaduck_b200 import synthetic

DAY = 86400.0


def case_gyre(n=120):
    ds = synthetic.gyre(50, 50)
    rng = np.random.RandomState(2023)
    x = rng.random_sample(n) * 2 - 1
    y = rng.random_sample(n) * 2 - 1
    return dict(
        name="gyre",
        ds=ds,
        x=x, y=y, z=np.full(n, -1.0), t=np.zeros(n),
        stops=[300.0, 700.0, 1500.0],
        kw=dict(wname=None, transport=True),
    )


def case_overturning(n=200):
    ds = synthetic.overturning_cell(100, 50)
    rng = np.random.RandomState(7)
    x = rng.uniform(-1, 1, n)
    z = rng.uniform(-0.1, -1000, n)
    return dict(
        name="overturning",
        ds=ds,
        x=x, y=np.zeros(n), z=z, t=np.zeros(n),
        stops=[-2000.0, -5000.0],
        kw=dict(transport=True),
    )


def case_llc_transport(n=300, ngrid=30, nz=10):
    ds = synthetic.llc(n=ngrid, nz=nz)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=1, depth_range=(5, 300))
    return dict(
        name="llc_transport",
        ds=ds,
        x=lon, y=lat, z=dep, t=np.zeros(n),
        stops=[20 * DAY, 60 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True),
    )


def case_llc_velocity(n=300, ngrid=30, nz=10):
    ds = synthetic.llc(n=ngrid, nz=nz, dtype=np.float32)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=2, depth_range=(5, 300))
    return dict(
        name="llc_velocity_f32",
        ds=ds,
        x=lon, y=lat, z=dep, t=np.zeros(n),
        stops=[15 * DAY, -10 * DAY],
        kw=dict(transport=False),
    )


def case_llc_time(n=200, ngrid=24, nz=6):
    ds = synthetic.llc(n=ngrid, nz=nz, nt=4)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=3, depth_range=(5, 120))
    return dict(
        name="llc_time",
        ds=ds,
        x=lon, y=lat, z=dep, t=np.full(n, 1.0 * DAY),
        stops=[4 * DAY, 12 * DAY, 17 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True),
        list_of_time=True,
    )


def case_llc_surface2d(n=300, ngrid=48):
    ds = synthetic.llc(n=ngrid, nz=0)
    lon, lat, _ = synthetic.llc_seed_particles(ds, n, seed=4)
    return dict(
        name="llc_surface2d",
        ds=ds,
        x=lon, y=lat, z=np.full(n, -1.0), t=np.zeros(n),
        stops=[10 * DAY, 30 * DAY],
        kw=dict(wname=None, transport=False),
    )


def case_llc_surface_time(n=300, ngrid=40):
    """The shape of BASELINE.json configs[3] (LLC4320 surface, two time snapshots used, reference
    docs/sciserver_notebooks/LLC4320.md:91-131) at test size: 2-D transports with a time axis, stops every few
    hours, the run crosses time_midp[0] so the velocity slab is swapped once."""
    ds = synthetic.llc(n=ngrid, nz=0, nt=3, u0=6.0)
    lon, lat, _ = synthetic.llc_seed_particles(ds, n, seed=14)
    return dict(
        name="llc_surface_time",
        ds=ds,
        x=lon, y=lat, z=np.full(n, -1.0), t=np.full(n, 4.5 * DAY),
        stops=[4.625 * DAY, 4.875 * DAY, 5.25 * DAY, 6.0 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname=None, transport=True),
        list_of_time=True,
    )


def case_llc_bool_array(ngrid=24, nz=8):
    """Particles seeded by ``from_bool_array`` (reference eulerian.py:219): volume-weighted random
    positions inside a masked region, then advected."""
    ds = synthetic.llc(n=ngrid, nz=nz)
    yc = np.asarray(ds["YC"])
    xc = np.asarray(ds["XC"])
    region = (np.abs(yc - 20.0) < 14.0) & (np.abs(xc + 10.0) < 16.0)
    bool_array = np.zeros((nz, 13, ngrid, ngrid), bool)
    bool_array[1:4] = region[None] & (np.asarray(ds["maskC"])[1:4] > 0)
    return dict(
        name="llc_bool_array",
        ds=ds,
        seeding=dict(bool_array=bool_array, num=600, random_seed=2024, t=0.0),
        stops=[15 * DAY, 40 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True),
    )


def case_channel_kickback(n=250):
    """Zonally periodic channel, velocities in m/s (transport=False), free_surface="kick_back"."""
    ds = synthetic.channel()
    rng = np.random.default_rng(12)
    x = rng.uniform(0.0, 360.0, n)
    y = rng.uniform(-46.0, 46.0, n)
    z = -rng.uniform(2.0, 350.0, n)
    return dict(
        name="channel_kickback",
        ds=ds,
        x=x, y=y, z=z, t=np.zeros(n),
        stops=[25 * DAY, 70 * DAY, 40 * DAY],
        kw=dict(transport=False, free_surface="kick_back"),
    )


def _stay_south(p):
    """Particle callback: keep integrating only while south of 31.3N (not a grid line: a threshold on a cell wall is decided by the last bit) (reference lagrangian.py:757,787)."""
    return p.lat < 31.3


def case_llc_callback(n=200, ngrid=24, nz=6):
    ds = synthetic.llc(n=ngrid, nz=nz)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=6, lat_range=(-40.0, 45.0), depth_range=(5, 150))
    return dict(
        name="llc_callback",
        ds=ds,
        x=lon, y=lat, z=dep, t=np.zeros(n),
        stops=[25 * DAY, 50 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True, callback=_stay_south, max_iteration=40),
    )


def case_rectilinear(n=300):
    """1-D lon / lat grid (altimetry-like), velocities in m/s, x-periodic: the `rectilinear` horizontal scheme
    (reference ocedata.py:255, utils.py:602, lagrangian.py:667-672, :693-703)."""
    ds = synthetic.rectilinear(nlat=40, nlon=80)
    rng = np.random.default_rng(11)
    return dict(
        name="rectilinear",
        ds=ds,
        x=rng.uniform(-179.0, 179.0, n), y=rng.uniform(-65.0, 65.0, n), z=np.full(n, -1.0), t=np.zeros(n),
        stops=[80 * DAY, -50 * DAY],
        kw=dict(uname="u", vname="v", wname=None, transport=False),
    )


def case_local_cartesian(n=300, ngrid=24, nz=6):
    """The LLC grid without its corner coordinates (XG, YG): the `local_cartesian` horizontal scheme
    (reference ocedata.py:252, utils.py:577, lagrangian.py:647-659, :693-703)."""
    from seaduck_b200 import minixr  # noqa: PLC0415

    full = synthetic.llc(n=ngrid, nz=nz)
    keep = {k: (tuple(full[k].dims), np.asarray(full[k])) for k in full.variables if k not in ("XG", "YG")}
    dims = {d for dd, _ in keep.values() for d in dd}
    ds = minixr.Dataset(coords={k: v for k, v in keep.items() if k in dims}, data_vars={k: v for k, v in keep.items() if k not in dims})
    lon, lat, dep = synthetic.llc_seed_particles(full, n, seed=8, depth_range=(5, 300))
    return dict(
        name="local_cartesian",
        ds=ds,
        x=lon, y=lat, z=dep, t=np.zeros(n),
        stops=[25 * DAY, -15 * DAY],
        kw=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True),
    )


# cases only the CPU oracle takes (none: the device runs all three horizontal schemes)
ORACLE_ONLY_CASES = []

ALL_CASES = [case_gyre, case_overturning, case_llc_transport, case_llc_velocity, case_llc_time, case_llc_surface2d, case_llc_surface_time, case_llc_bool_array, case_llc_callback, case_channel_kickback, case_rectilinear, case_local_cartesian]

LOG_INT = ["fc", "it", "izl", "iy", "ix", "vs"]
LOG_FLT = ["rzl", "rx", "ry", "tt", "uu", "vv", "ww", "du", "dv", "dw", "xx", "yy", "zz"]
FINAL_INT = ["face", "iy", "ix", "izl_lin", "it"]
FINAL_FLT = ["lon", "lat", "dep", "t", "rx", "ry", "rzl_lin", "u", "v", "w", "du", "dv", "dw"]


def assert_close(name, got, ref, rtol=1e-9):
    """|got - ref| <= rtol * max(|ref|, field scale): the 1e-9 relative bound of the north star."""
    got = np.asarray(got, float)
    ref = np.asarray(ref, float)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    scale = max(float(np.nanmax(np.abs(ref))), 1e-300)
    if name in ("rx", "ry", "rzl", "rzl_lin", "lon", "lat", "xx", "yy"):
        scale = max(scale, 1.0)
    err = np.abs(got - ref)
    bad = err > rtol * np.maximum(np.abs(ref), scale)
    assert not bad.any(), f"{name}: {int(bad.sum())} of {ref.size} beyond {rtol} (max err {err.max():.3e}, scale {scale:.3e})"


# ---- Lagrangian budget: terms and wall concentrations of one model time step ------------------------------

RHS_TERMS = ["adv", "dif", "forc"]


def budget_fields(shape, seed):
    """Budget terms + wall concentrations of one time step: name -> array (a stand-in for the model output)."""
    rng = np.random.default_rng(seed)
    nz = shape[0]
    out = {k: rng.normal(size=shape) * 1e-7 for k in RHS_TERMS + ["lhs"]}
    dead = rng.random(shape) < 0.03  # cells where every term vanishes: the relaxation divides 0 by 0
    for k in RHS_TERMS:
        out[k][dead] = 0.0
    out["forc"][rng.random(shape) < 0.3] = 0.0
    for k in ("sxprime", "syprime"):
        out[k] = rng.normal(size=shape)
    out["szprime"] = rng.normal(size=(nz + 1,) + tuple(shape[1:]))
    out["szprime_same"] = out["szprime"][:nz].copy()
    for k in ("sxprime", "syprime", "szprime"):
        out[k][rng.random(out[k].shape) < 0.02] = np.nan
    return out
