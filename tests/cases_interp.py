"""Seeded Eulerian-interpolation workloads shared by the golden generator (unmodified reference),
the numpy oracle tests and the GPU parity tests.

A *kernel spec* is a dict of ``KnW`` keyword arguments (``kernel`` may be the name of a stencil in
``STENCILS``), so the same case builds reference ``seaduck.KnW`` objects or the product's.
A *job* is ``(label, var_name, knw_spec, vec_transform)`` where ``var_name`` / ``knw_spec`` follow
the nesting ``Position.interpolate`` accepts (str | tuple | list).
"""

import numpy as np

from seaduck_b200 import synthetic

DAY = 86400.0

STENCILS = {
    "cross9": np.array([[0, 0], [0, 1], [0, 2], [0, -1], [0, -2], [-1, 0], [-2, 0], [1, 0], [2, 0]]),
    "cross5": np.array([[0, 0], [0, 1], [0, -1], [-1, 0], [1, 0]]),
    "square9": np.array([[i, j] for i in (-1, 0, 1) for j in (-1, 0, 1)]),
    "uv": np.array([[0, 0], [1, 0], [0, 1]]),
    "u2": np.array([[0, 0], [1, 0]]),
    "v2": np.array([[0, 0], [0, 1]]),
    "one": np.array([[0, 0]]),
    "xline5": np.array([[0, 0], [-1, 0], [1, 0], [-2, 0], [2, 0]]),
}

ALL9 = list(range(9))

# the kernels Particle uses (reference lagrangian.py:25-53)
UKNW = dict(kernel="uv", inheritance=[[0, 1]], ignore_mask=True)
VKNW = dict(kernel="uv", inheritance=[[0, 2]], ignore_mask=True)
DUKNW = dict(kernel="uv", inheritance=[[0, 1]], hkernel="dx", h_order=1, ignore_mask=True)
DVKNW = dict(kernel="uv", inheritance=[[0, 2]], hkernel="dy", h_order=1, ignore_mask=True)
WKNW = dict(kernel="one", inheritance=None, vkernel="linear", ignore_mask=True)
DWKNW = dict(kernel="one", inheritance=None, vkernel="dz", ignore_mask=True)
UKNW_S = dict(kernel="u2", inheritance=None, ignore_mask=True)
VKNW_S = dict(kernel="v2", inheritance=None, ignore_mask=True)
DUKNW_S = dict(kernel="u2", inheritance=None, hkernel="dx", h_order=1, ignore_mask=True)


def make_knw(KnW, spec):
    """Build KnW objects (of either package) from a spec with the same nesting."""
    if isinstance(spec, tuple):
        return tuple(make_knw(KnW, s) for s in spec)
    if isinstance(spec, list):
        return [make_knw(KnW, s) for s in spec]
    kw = dict(spec)
    if isinstance(kw.get("kernel"), str):
        kw["kernel"] = STENCILS[kw["kernel"]].copy()
    return KnW(**kw)


def _bathymetry(ds, seed=5):
    """Depth-dependent land mask + NaN in the tracers at dry cells (exercises the vertical cascade,
    the bottom ghost point and nan_to_num)."""
    mask = np.array(ds["maskC"]).astype(np.int8)
    nz = mask.shape[0]
    xc, yc = np.deg2rad(np.array(ds["XC"])), np.deg2rad(np.array(ds["YC"]))
    depth_frac = 0.55 + 0.45 * np.sin(3 * xc + 1.0) * np.cos(2 * yc)  # fraction of levels that are wet
    nwet = np.clip(np.round(depth_frac * nz), 1, nz).astype(int)
    lev = np.arange(nz)[:, None, None, None]
    mask = mask * (lev < nwet[None]).astype(np.int8)
    ds["maskC"].values[...] = mask
    ds["hFacC"].values[...] = mask
    rng = np.random.default_rng(seed)
    for name in ("T", "S"):
        arr = ds[name].values
        dry = np.broadcast_to(mask == 0, arr.shape)
        arr[dry & (rng.random(arr.shape) < 0.5)] = np.nan
    return ds


def _sphere_points(n, seed, lat_range=(-72.0, 88.0), depth=(5.0, 1500.0), t_range=(0.0, 19.0 * DAY)):
    lon, lat, _ = synthetic.llc_seed_particles(None, n, seed=seed, lat_range=lat_range)
    rng = np.random.default_rng(seed + 100)
    dep = -rng.uniform(depth[0], depth[1], n)
    t = rng.uniform(t_range[0], t_range[1], n)
    return lon, lat, dep, t


def case_llc3d(n=1500, ngrid=24, nz=8):
    ds = _bathymetry(synthetic.llc(n=ngrid, nz=nz, nt=3, land_blobs=True, tracers=True))
    lon, lat, dep, t = _sphere_points(n, seed=21, depth=(5.0, 1450.0))
    lin = dict(vkernel="linear", tkernel="linear")
    jobs = [
        ("T_default", "T", {}, True),
        ("T_linear", "T", lin, True),
        ("T_dz_dt", "T", dict(vkernel="dz", tkernel="dt"), True),
        ("S_dx", "S", dict(hkernel="dx", h_order=1, inheritance=[ALL9, [0, 1, 3, 5, 7]]), True),
        ("S_dy2", "S", dict(hkernel="dy", h_order=2, inheritance=[ALL9], tkernel="linear"), True),
        ("T_cross5", "T", dict(kernel="cross5", vkernel="linear"), True),
        ("T_ignore", "T", dict(ignore_mask=True, vkernel="linear"), True),
        ("S_xline_dx2", "S", dict(kernel="xline5", hkernel="dx", h_order=2, inheritance=[[0, 1, 2, 3, 4], [0, 1, 2]]), True),
        ("uv_particle", ("UVELMASS", "VVELMASS"), (UKNW, VKNW), False),
        ("uv_particle_rot", ("UVELMASS", "VVELMASS"), (UKNW, VKNW), True),
        ("duv_particle", ("UVELMASS", "VVELMASS"), (DUKNW, DVKNW), False),
        ("uv_masked", ("UVELMASS", "VVELMASS"), ({}, {}), True),
        ("uv_masked_lin", ("UVELMASS", "VVELMASS"), (lin, lin), True),
        ("uv_ignore_lin", ("utrans", "vtrans"), (dict(ignore_mask=True, **lin), dict(ignore_mask=True, **lin)), False),
        ("w_particle", "WVELMASS", WKNW, True),
        ("dw_particle", "WVELMASS", DWKNW, True),
        ("w_masked", "WVELMASS", dict(vkernel="linear", tkernel="linear"), True),
        ("w_nearest", "WVELMASS", {}, True),
        ("list_mixed", ["T", ("UVELMASS", "VVELMASS"), "S"], [lin, ({}, {}), dict(hkernel="dx", h_order=1, inheritance=[ALL9])], True),
    ]
    return dict(name="llc3d", ds=ds, x=lon, y=lat, z=dep, t=t, jobs=jobs, reach=2)


def case_llc3d_square(n=800, ngrid=20, nz=6):
    """Rectangular stencils (diagonal moves); static fields."""
    ds = _bathymetry(synthetic.llc(n=ngrid, nz=nz, land_blobs=True, tracers=True), seed=6)
    lon, lat, dep, _ = _sphere_points(n, seed=22, depth=(5.0, 900.0))
    jobs = [
        ("T_square", "T", dict(kernel="square9"), True),
        ("T_square_dx", "T", dict(kernel="square9", hkernel="dx", h_order=1, inheritance=[ALL9], vkernel="linear"), True),
        ("S_square_dy", "S", dict(kernel="square9", hkernel="dy", h_order=1, inheritance=[ALL9]), True),
        ("uv_square", ("UVELMASS", "VVELMASS"), (dict(kernel="square9"), dict(kernel="square9")), True),
    ]
    return dict(name="llc3d_square", ds=ds, x=lon, y=lat, z=dep, t=None, jobs=jobs, reach=2)


def case_llc2d(n=800, ngrid=30):
    """Surface-only LLC (no Z, no maskC): every kernel effectively unmasked."""
    ds = synthetic.llc(n=ngrid, nz=0, tracers=True)
    lon, lat, _, _ = _sphere_points(n, seed=23)
    jobs = [
        ("T_ignore", "T", dict(ignore_mask=True), True),
        ("uv_particle", ("UVELMASS", "VVELMASS"), (UKNW, VKNW), True),
        ("uv_ignore9", ("UVELMASS", "VVELMASS"), (dict(ignore_mask=True), dict(ignore_mask=True)), True),
    ]
    return dict(name="llc2d", ds=ds, x=lon, y=lat, z=None, t=None, jobs=jobs, reach=2)


def case_gyre(n=600):
    """Closed 2-D box without faces: vectors are split into scalars on staggered grids."""
    ds = synthetic.gyre(40, 50)
    rng = np.random.RandomState(31)
    x = rng.uniform(-0.999, 0.999, n)
    y = rng.uniform(-0.999, 0.999, n)
    jobs = [
        ("uv_scalar", ("UVELMASS", "VVELMASS"), (UKNW_S, VKNW_S), False),
        ("du_scalar", "UVELMASS", DUKNW_S, True),
        ("psi_cross", "streamfunc", dict(ignore_mask=True), True),
        ("psi_square_dy", "streamfunc", dict(kernel="square9", hkernel="dy", h_order=1, inheritance=[ALL9], ignore_mask=True), True),
    ]
    return dict(name="gyre", ds=ds, x=x, y=y, z=None, t=None, jobs=jobs, reach=0)


def case_overturning(n=500):
    """x-periodic-looking x-z slab with Z and Zl variables, no land mask."""
    ds = synthetic.overturning_cell(60, 20)
    rng = np.random.RandomState(32)
    x = rng.uniform(-0.99, 0.99, n)
    z = rng.uniform(-30.0, -970.0, n)
    jobs = [
        ("w_lin", "WVELMASS", WKNW, True),
        ("u_scalar", "UVELMASS", dict(kernel="u2", inheritance=None, ignore_mask=True, vkernel="linear"), True),
        ("psi_dz", "streamfunc", dict(kernel="xline5", ignore_mask=True, vkernel="dz"), True),
    ]
    return dict(name="overturning", ds=ds, x=x, y=np.zeros(n), z=z, t=None, jobs=jobs, reach=0)


def case_rectilinear(n=500):
    """1-D lon / lat grid (altimetry-like, x-periodic), unstaggered u, v in m/s: the `rectilinear` horizontal scheme."""
    ds = synthetic.rectilinear(nlat=40, nlon=80, nt=3)
    rng = np.random.RandomState(33)
    x = rng.uniform(-179.9, 179.9, n)
    y = rng.uniform(-70.0, 70.0, n)
    t = rng.uniform(0.0, 9.5 * DAY, n)
    jobs = [
        ("u_cross", "u", dict(ignore_mask=True), True),
        ("u_lin_t", "u", dict(ignore_mask=True, tkernel="linear"), True),
        ("v_dx", "v", dict(kernel="cross5", inheritance=[[0, 1, 2, 3, 4]], hkernel="dx", h_order=1, ignore_mask=True), True),
        ("uv_vector", ("u", "v"), (dict(ignore_mask=True), dict(ignore_mask=True)), True),
    ]
    return dict(name="rectilinear", ds=ds, x=x, y=y, z=None, t=t, jobs=jobs, reach=2)


def case_local_cartesian(n=500, ngrid=24, nz=6):
    """The LLC grid without XG, YG: the `local_cartesian` horizontal scheme for locating; the stencils are the same."""
    from seaduck_b200 import minixr  # noqa: PLC0415

    full = synthetic.llc(n=ngrid, nz=nz, tracers=True)
    keep = {k: (tuple(full[k].dims), np.asarray(full[k])) for k in full.variables if k not in ("XG", "YG")}
    dims = {d for dd, _ in keep.values() for d in dd}
    ds = minixr.Dataset(coords={k: v for k, v in keep.items() if k in dims}, data_vars={k: v for k, v in keep.items() if k not in dims})
    lon, lat, _ = synthetic.llc_seed_particles(full, n, seed=34, lat_range=(-70.0, 85.0))
    z = -np.random.RandomState(35).uniform(5.0, 600.0, n)
    jobs = [
        ("T_masked", "T", {}, True),
        ("uv_particle", ("UVELMASS", "VVELMASS"), (UKNW, VKNW), True),
    ]
    return dict(name="local_cartesian", ds=ds, x=lon, y=lat, z=z, t=None, jobs=jobs, reach=2)


ALL_CASES = [case_llc3d, case_llc3d_square, case_llc2d, case_gyre, case_overturning, case_rectilinear, case_local_cartesian]
# cases only the CPU oracle takes (none: the device runs all three horizontal schemes)
ORACLE_ONLY_CASES = []


def flatten_result(res):
    """Interpolation output (array | tuple | list) -> list of arrays in a fixed order."""
    if res is None:
        return []
    if isinstance(res, (tuple, list)):
        out = []
        for r in res:
            out += flatten_result(r)
        return out
    return [np.asarray(res, float)]


def assert_values_close(label, got, ref, rtol=1e-9):
    """NaN pattern identical; |got - ref| <= rtol * max(|ref|, scale of the field)."""
    got = np.asarray(got, float)
    ref = np.asarray(ref, float)
    assert got.shape == ref.shape, f"{label}: shape {got.shape} vs {ref.shape}"
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert (nan_g == nan_r).all(), f"{label}: NaN pattern differs at {int((nan_g != nan_r).sum())} of {ref.size} points"
    ok = ~nan_r
    if not ok.any():
        return
    scale = float(np.abs(ref[ok]).max())
    err = np.abs(got[ok] - ref[ok])
    bad = err > rtol * np.maximum(np.abs(ref[ok]), scale) + 1e-300
    assert not bad.any(), f"{label}: {int(bad.sum())} of {int(ok.sum())} beyond {rtol} (max err {err.max():.3e}, scale {scale:.3e})"


def case_oceinterp(n=150, ngrid=24, nz=8):
    """``OceInterp`` (reference oceinterp.py:14): Eulerian call and Lagrangian call on one dataset."""
    ds = synthetic.llc(n=ngrid, nz=nz, nt=4, tracers=True)
    lon, lat, dep, _ = _sphere_points(n, seed=41, lat_range=(-55.0, 60.0), depth=(5.0, 400.0))
    return dict(
        name="oceinterp", ds=ds, x=lon, y=lat, z=dep,
        t_euler=12.5 * DAY,
        t_lagrange=np.array([1.0, 4.0, 12.0, 17.0]) * DAY,
        var_euler=["T", ("UVELMASS", "VVELMASS"), "WVELMASS"],
        var_lagrange=["T", "__particle.lon", "__particle.lat", "__particle.dep", "__particle.ix", ("UVELMASS", "VVELMASS")],
        lagrange_kwarg=dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True),
    )
