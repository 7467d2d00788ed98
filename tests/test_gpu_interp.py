"""GPU: ``Position.from_latlon`` + ``Position.interpolate`` (CUDA, through the C ABI) against the
golden vectors produced by the unmodified reference (``oracle/make_golden_interp.py``).
Indices bit-exact, NaN pattern identical, values within 1e-9 relative."""

import os
import re

import numpy as np
import pytest

import cases_interp as ci

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _position(case, gold):
    import seaduck_b200 as sd

    od = sd.OceData(case["ds"])
    get = lambda k: gold[f"in_{k}"] if f"in_{k}" in gold else None  # noqa: E731
    pos = sd.Position().from_latlon(x=get("x"), y=get("y"), z=get("z"), t=get("t"), data=od)
    return sd, od, pos


@pytest.mark.parametrize("make", ci.ALL_CASES, ids=lambda f: f.__name__)
def test_locate_matches_reference(make):
    case = make()
    gold = np.load(os.path.join(GOLD, f"interp_{case['name']}.npz"))
    _, _, pos = _position(case, gold)
    for key in ("face", "iy", "ix", "iz", "izl", "izl_lin", "it", "iz_lin", "it_lin"):
        if f"pos_{key}" in gold:
            np.testing.assert_array_equal(getattr(pos, key), gold[f"pos_{key}"], err_msg=key)
    for key in ("rx", "ry", "rz", "rzl", "rzl_lin", "rt", "rz_lin", "rt_lin"):
        if f"pos_{key}" in gold:
            ci.assert_values_close(key, getattr(pos, key), gold[f"pos_{key}"])


@pytest.mark.parametrize("make", ci.ALL_CASES, ids=lambda f: f.__name__)
def test_interpolate_matches_reference(make):
    case = make()
    gold = np.load(os.path.join(GOLD, f"interp_{case['name']}.npz"))
    sd, _, pos = _position(case, gold)
    failures = []
    for label, var, spec, vec_transform in case["jobs"]:
        knw = ci.make_knw(sd.KnW, spec)
        res = pos.interpolate(var, knw, vec_transform=vec_transform)
        flat = ci.flatten_result(res)
        for k, arr in enumerate(flat):
            try:
                ci.assert_values_close(f"{label}[{k}]", arr, gold[f"job_{label}_{k}"])
            except AssertionError as exc:
                failures.append(str(exc))
        assert len(flat) == len([k for k in gold.files if re.fullmatch(rf"job_{label}_\d+", k)]), label
    assert not failures, "\n".join(failures)


def test_interpolate_output_nesting():
    """str -> array, tuple -> (u, v), list -> list; dict kernels; repeated entries."""
    case = ci.case_llc2d(n=64)
    import seaduck_b200 as sd

    od = sd.OceData(case["ds"])
    pos = sd.Position().from_latlon(x=case["x"], y=case["y"], data=od)
    k = sd.KnW(ignore_mask=True)
    a = pos.interpolate("T", k)
    assert isinstance(a, np.ndarray) and a.shape == (64,)
    uv = pos.interpolate(("UVELMASS", "VVELMASS"), (k, k))
    assert isinstance(uv, tuple) and len(uv) == 2
    out = pos.interpolate(["T", ("UVELMASS", "VVELMASS"), "T"], {"T": k, ("UVELMASS", "VVELMASS"): (k, k)})
    assert isinstance(out, list) and len(out) == 3
    np.testing.assert_array_equal(out[0], a)
    np.testing.assert_array_equal(out[2], a)
    np.testing.assert_array_equal(out[1][0], uv[0])
    with pytest.raises(ValueError):
        pos.interpolate(["T", "T"], [k])
    with pytest.raises(ValueError):
        pos.interpolate("T", (k, k, k))


def test_interpolate_after_editing_position():
    """Setting a rel-coord by hand invalidates the device copy made by from_latlon."""
    case = ci.case_llc2d(n=32)
    import seaduck_b200 as sd

    od = sd.OceData(case["ds"])
    pos = sd.Position().from_latlon(x=case["x"], y=case["y"], data=od)
    k = sd.KnW(ignore_mask=True, kernel=ci.STENCILS["one"], inheritance=None)
    a = pos.interpolate("T", k)
    pos.ix = (pos.ix + 1) % 30
    b = pos.interpolate("T", k)
    t = np.asarray(case["ds"]["T"])
    np.testing.assert_allclose(b, t[pos.face, pos.iy, pos.ix])
    assert not np.allclose(a, b)


def test_oceinterp_matches_reference():
    """``OceInterp`` in both modes (reference oceinterp.py:14) against the reference's own output."""
    import seaduck_b200 as sd

    case = ci.case_oceinterp()
    gold = np.load(os.path.join(GOLD, "oceinterp.npz"))
    od = sd.OceData(case["ds"])
    res = sd.OceInterp(od, case["var_euler"], case["x"], case["y"], case["z"], case["t_euler"])
    flat = ci.flatten_result(res)
    assert len(flat) == len([k for k in gold.files if k.startswith("euler_")])
    for k, arr in enumerate(flat):
        ci.assert_values_close(f"euler[{k}]", arr, gold[f"euler_{k}"])
    stops, res = sd.OceInterp(
        od, case["var_lagrange"], case["x"], case["y"], case["z"], case["t_lagrange"], lagrangian=True,
        lagrange_kwarg=case["lagrange_kwarg"],
    )
    np.testing.assert_allclose(stops, gold["stops"])
    flat = ci.flatten_result(res)
    assert len(flat) == len([k for k in gold.files if k.startswith("lagrange_")])
    for k, arr in enumerate(flat):
        ci.assert_values_close(f"lagrange[{k}]", arr, gold[f"lagrange_{k}"])
    with pytest.raises(ValueError):
        sd.OceInterp(od, "T", case["x"], case["y"], case["z"], 5)  # int time: "time format not supported"
    with pytest.raises(AttributeError):
        sd.OceInterp(od, ["__particle.lon"], case["x"], case["y"], case["z"], case["t_euler"])


def test_sorted_execution_equals_plain(monkeypatch):
    """Large located Positions are interpolated in cell-sorted order and scattered back: same numbers."""
    import seaduck_b200 as sd
    from seaduck_b200 import interp

    case = ci.case_llc3d(n=3000)
    od = sd.OceData(case["ds"])
    knw = sd.KnW(vkernel="linear", tkernel="linear")
    var = ["T", ("UVELMASS", "VVELMASS")]
    kernels = [knw, (knw, knw)]
    pos = sd.Position().from_latlon(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od)
    plain = ci.flatten_result(pos.interpolate(list(var), list(kernels)))
    monkeypatch.setattr(interp._DevicePoints, "SORT_THRESHOLD", 100)
    pos2 = sd.Position().from_latlon(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od)
    sorted_ = ci.flatten_result(pos2.interpolate(list(var), list(kernels)))
    assert "_perm" in pos2._located
    again = ci.flatten_result(pos2.interpolate(list(var), list(kernels)))  # cached permutation
    for a, b, c in zip(plain, sorted_, again):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)


def test_partition_of_unity_at_scale():
    """Size-independent property on 2 x 10^6 scattered points (LLC90 x 50): whatever stencil the land mask
    selects, interpolation weights sum to one and derivative weights to zero -- a constant field comes
    back as that constant (or NaN where even the centre cell is dry), its derivatives as zero."""
    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=50, land_blobs=True, tracers=True)
    ds["T"].values[...] = 7.25
    od = sd.OceData(ds)
    n = 2_000_000
    lon, lat, _ = synthetic.llc_seed_particles(ds, n, seed=31, lat_range=(-72.0, 88.0))
    dep = -np.random.default_rng(32).uniform(5.0, 4000.0, n)
    pos = sd.Position().from_latlon(x=lon, y=lat, z=dep, data=od)
    all9 = list(range(9))
    val, lin, ddx, ddz = pos.interpolate(
        ["T", "T", "T", "T"],
        [sd.KnW(), sd.KnW(vkernel="linear"), sd.KnW(hkernel="dx", h_order=1, inheritance=[all9, [0, 1, 3, 5, 7]]), sd.KnW(vkernel="dz")],
    )
    wet = ~np.isnan(val)
    assert 0.5 < wet.mean() < 0.95  # the land blobs and the bathymetry-free ocean
    np.testing.assert_allclose(val[wet], 7.25, rtol=0, atol=1e-12)
    ok = ~np.isnan(lin)
    np.testing.assert_allclose(lin[ok], 7.25, rtol=0, atol=1e-11)
    ok = ~np.isnan(ddx)
    assert ok.any()
    np.testing.assert_allclose(ddx[ok], 0.0, rtol=0, atol=1e-11)
    ok = ~np.isnan(ddz)
    np.testing.assert_allclose(ddz[ok], 0.0, rtol=0, atol=1e-11)


@pytest.mark.parametrize("shear", [0.0, 0.9, 1.8])
def test_locate_on_a_sheared_grid_matches_kdtree(shear, oracle_lib):
    """Nearest cell centre on a sheared curvilinear box with NaN (land) centres: the device search (bucket grid, greedy
    walk over the 8 neighbours, second-ring confirmation) against scipy's cKDTree on the same centres, the way the
    reference finds it (utils.py:225,309).  With shear the Voronoi neighbours of a cell are not all among its 8 index
    neighbours, which is what the second ring is for."""
    import seaduck_b200 as sd
    from seaduck_b200 import minixr

    ny, nx = 60, 80
    dlon, dlat = 0.25, 0.2
    j, i = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    xg = -30.0 + dlon * i + shear * dlon * (j - ny / 2) * 0.8
    yg = 10.0 + dlat * j + 0.02 * np.sin(i / 7.0)
    xc = 0.25 * (xg[:-1, :-1] + xg[1:, :-1] + xg[:-1, 1:] + xg[1:, 1:])
    yc = 0.25 * (yg[:-1, :-1] + yg[1:, :-1] + yg[:-1, 1:] + yg[1:, 1:])
    rng = np.random.default_rng(17)
    land = rng.random((ny, nx)) < 0.04
    xc_nan, yc_nan = xc.copy(), yc.copy()
    xc_nan[land] = np.nan
    yc_nan[land] = np.nan
    ds = minixr.Dataset(coords=dict(XC=(["Y", "X"], xc_nan), YC=(["Y", "X"], yc_nan), XG=(["Yp1", "Xp1"], xg), YG=(["Yp1", "Xp1"], yg)), data_vars={})
    od = sd.OceData(ds)
    n = 20000
    # points well inside the grid (the outer two rows / columns are left alone: past the edge "nearest" is all there is)
    jj = rng.uniform(3, ny - 3, n)
    ii = rng.uniform(3, nx - 3, n)
    lon = -30.0 + dlon * ii + shear * dlon * (jj - ny / 2) * 0.8
    lat = 10.0 + dlat * jj
    pos = sd.Position().from_latlon(x=lon, y=lat, data=od)
    og = oracle_lib.OracleGrid(od)
    ry, rx = og.find_ind_h(lon, lat)
    same = (pos.iy == ry) & (pos.ix == rx)
    if not same.all():
        # a query within rounding of a Voronoi edge may legitimately go either way: allow it only if both centres are
        # equally far to 1e-12 relative
        bad = np.flatnonzero(~same)
        from oracle.oracle import spherical2cartesian

        def d2(y_, x_):
            cx, cy, cz = spherical2cartesian(og.YC[y_, x_].astype(float), og.XC[y_, x_].astype(float))
            qx, qy, qz = spherical2cartesian(lat[bad], lon[bad])
            return (cx - qx) ** 2 + (cy - qy) ** 2 + (cz - qz) ** 2

        mine, ref = d2(pos.iy[bad], pos.ix[bad]), d2(ry[bad], rx[bad])
        worse = np.abs(mine - ref) > 1e-12 * ref
        off = np.stack([pos.iy[bad] - ry[bad], pos.ix[bad] - rx[bad]], axis=1)[worse][:8]
        assert not worse.any(), (
            f"{int(worse.sum())} of {n} points in a farther cell (shear {shear}); index offsets (dy, dx) of the first: {off.tolist()}, "
            f"distance ratios {np.sqrt(mine[worse] / ref[worse])[:8].round(4).tolist()}"
        )


@pytest.mark.parametrize("kind", ["llc", "x_periodic"])
def test_locate_next_to_face_edges_matches_kdtree(kind, oracle_lib):
    """Points scattered around the centres of the cells on the rim of every face (so that many of them belong to a cell
    across the edge, on another face or past the periodic seam): the device search, whose 8 neighbours of a rim cell
    come from the halo ring of the centre table, against scipy's cKDTree on all centres (reference utils.py:225,309)."""
    import seaduck_b200 as sd
    from oracle.oracle import spherical2cartesian
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=24, nz=4) if kind == "llc" else synthetic.channel(ny=30, nx=72, nz=4)
    od = sd.OceData(ds)
    og = oracle_lib.OracleGrid(od)
    xc, yc = np.asarray(og.XC, float), np.asarray(og.YC, float)
    rim = np.zeros(xc.shape, bool)
    rim[..., 0, :] = rim[..., -1, :] = rim[..., :, 0] = rim[..., :, -1] = True
    rim &= np.isfinite(xc) & np.isfinite(yc) & (np.abs(yc) < 88.0)
    idx = np.argwhere(rim)
    rng = np.random.default_rng(23)
    rep = 6
    lon0 = np.repeat(xc[rim], rep)
    lat0 = np.repeat(yc[rim], rep)
    # a jitter of about one cell width in every direction (the grid spacing, in degrees, from the neighbouring centre)
    step = np.abs(np.diff(yc, axis=-2)).max()
    lon = lon0 + rng.uniform(-1.2, 1.2, lon0.size) * step / np.maximum(np.cos(np.deg2rad(lat0)), 0.05)
    lat = np.clip(lat0 + rng.uniform(-1.2, 1.2, lon0.size) * step, -89.5, 89.5)
    lon = (lon + 180.0) % 360.0 - 180.0
    pos = sd.Position().from_latlon(x=lon, y=lat, data=od)
    ref = og.find_ind_h(lon, lat)
    mine = (pos.face, pos.iy, pos.ix) if len(ref) == 3 else (pos.iy, pos.ix)
    same = np.ones(lon.size, bool)
    for a, b in zip(mine, ref):
        same &= np.asarray(a) == np.asarray(b)
    moved = np.zeros(lon.size, bool)
    for a, col in zip(mine, np.repeat(idx, rep, axis=0).T):
        moved |= np.asarray(a) != col
    assert moved.mean() > 0.2, "the jitter is supposed to push many points into other cells"
    if not same.all():
        bad = np.flatnonzero(~same)

        def d2(ind):
            sel = tuple(np.asarray(a)[bad] for a in ind)
            cx, cy, cz = spherical2cartesian(yc[sel], xc[sel])
            qx, qy, qz = spherical2cartesian(lat[bad], lon[bad])
            return (cx - qx) ** 2 + (cy - qy) ** 2 + (cz - qz) ** 2

        got, want = d2(mine), d2(ref)
        worse = ~(np.abs(got - want) <= 1e-12 * want)
        assert not worse.any(), (
            f"{int(worse.sum())} of {lon.size} points in a farther cell than cKDTree's ({kind}); first: "
            f"{[tuple(int(np.asarray(a)[bad][worse][0]) for a in mine)]} vs {[tuple(int(np.asarray(a)[bad][worse][0]) for a in ref)]}, "
            f"distance ratios {np.sqrt(got[worse] / want[worse])[:6].round(4).tolist()}"
        )


def test_multi_field_launch_equals_single_launches():
    """Scalars that share a kernel go through one ``sdk_interp_scalar_multi`` launch: bit for bit what one launch per
    variable gives, for nearest / linear / derivative kernels in z and t, masked and unmasked."""
    import seaduck_b200 as sd

    case = ci.case_llc3d(n=4000)
    od = sd.OceData(case["ds"])
    pos = sd.Position().from_latlon(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od)
    for spec in ({}, dict(vkernel="linear", tkernel="linear"), dict(vkernel="dz", tkernel="dt"), dict(ignore_mask=True, vkernel="linear"),
                 dict(hkernel="dx", h_order=1, inheritance=[ci.ALL9, [0, 1, 3, 5, 7]]), dict(kernel="cross5", tkernel="linear")):
        knw = ci.make_knw(sd.KnW, spec)
        together = pos.interpolate(["T", "S", "T"], knw)
        alone = [pos.interpolate("T", knw), pos.interpolate("S", knw)]
        np.testing.assert_array_equal(together[0], alone[0], err_msg=str(spec))
        np.testing.assert_array_equal(together[1], alone[1], err_msg=str(spec))
        np.testing.assert_array_equal(together[2], alone[0], err_msg=str(spec))
        assert np.isnan(alone[0]).any() or spec.get("ignore_mask")  # the land / bottom rows are part of the comparison


def test_interpolate_a_million_points_matches_oracle(oracle_lib):
    """BASELINE.json configs[2] at test size: T, S (masked 9-point kernel, one multi-field launch) and (u, v) at 10^6
    scattered points on an LLC90-shaped grid with land, through the cell-sorted copy of the points -- against the numpy
    oracle on the same points: NaN pattern identical, values to 1e-9."""
    import seaduck_b200 as sd
    from oracle import interp_oracle as io
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=20, land_blobs=True, tracers=True)
    od = sd.OceData(ds)
    n = 1_000_000
    lon, lat, _ = synthetic.llc_seed_particles(None, n, seed=3, lat_range=(-72.0, 88.0))
    dep = -np.random.default_rng(4).uniform(5.0, 3000.0, n)
    var = ["T", "S", ("UVELMASS", "VVELMASS")]
    uknw, vknw = ci.make_knw(sd.KnW, (ci.UKNW, ci.VKNW))
    kernels = [sd.KnW(), sd.KnW(), (uknw, vknw)]
    pos = sd.Position().from_latlon(x=lon, y=lat, z=dep, data=od)
    got = ci.flatten_result(pos.interpolate(list(var), list(kernels)))
    opos = io.OraclePosition(od, x=lon, y=lat, z=dep)
    same_cell = (pos.face == opos.face) & (pos.iy == opos.iy) & (pos.ix == opos.ix)
    assert same_cell.mean() > 1 - 1e-5  # (a query within rounding of a Voronoi edge may go either way)
    ref = ci.flatten_result(io.InterpOracle(od, opos).interpolate(list(var), list(kernels)))
    assert len(got) == len(ref) == 4
    for k, (g, r) in enumerate(zip(got, ref)):
        ci.assert_values_close(f"variable {k}", g[same_cell], r[same_cell])
    assert np.isnan(ref[0]).mean() > 0.02  # land is part of the comparison


def _same(a, b, label=""):
    fa, fb = ci.flatten_result(a), ci.flatten_result(b)
    assert len(fa) == len(fb), label
    for k, (x, y) in enumerate(zip(fa, fb)):
        np.testing.assert_array_equal(x, y, err_msg=f"{label}[{k}]")


@pytest.mark.parametrize("sort", [True, False])
def test_host_pipeline_equals_position_interpolate(sort, monkeypatch):
    """``sdk_interp_host`` (coordinates in, values out, chunks pipelined over three streams inside the library) returns
    bit for bit what ``Position.from_latlon`` + ``Position.interpolate`` return: several chunks with a ragged last one,
    time-dependent 3-D variables with nearest / linear / derivative kernels, a masked multi-field group and a vector
    with face rotation; with and without the (level, bucket) sort of the points."""
    import seaduck_b200 as sd
    from seaduck_b200 import interp

    monkeypatch.setenv("SDK_INTERP_CHUNKS", "5")
    case = ci.case_llc3d(n=700_001, ngrid=24, nz=8)
    od = sd.OceData(case["ds"])
    x, y, z, t = case["x"], case["y"], case["z"], case["t"]
    uknw, vknw = ci.make_knw(sd.KnW, (ci.UKNW, ci.VKNW))
    lin = sd.KnW(vkernel="linear", tkernel="linear")
    requests = [
        (["T", "S", ("UVELMASS", "VVELMASS")], [sd.KnW(), sd.KnW(), (uknw, vknw)]),
        (["T", "S", "T"], [lin, lin, sd.KnW(vkernel="dz", tkernel="dt")]),
        ("T", sd.KnW(ignore_mask=True)),
        (("UVELMASS", "VVELMASS"), (ci.make_knw(sd.KnW, ci.DUKNW), ci.make_knw(sd.KnW, ci.DVKNW))),
        (["WVELMASS"], [ci.make_knw(sd.KnW, ci.WKNW)]),
    ]
    pos = sd.Position().from_latlon(x=x, y=y, z=z, t=t, data=od)
    for var, knw in requests:
        ref = pos.interpolate(var, knw)
        got = interp.interpolate_host(od, x, y, z, t, var, knw, sort=sort)
        assert type(got) is type(ref)
        _same(got, ref, str(var))
    # pinned coordinates take the same route
    xp, yp, zp, tp = (interp.pinned(len(x)) for _ in range(4))
    xp[:], yp[:], zp[:], tp[:] = x, y, z, t
    var, knw = requests[0]
    _same(interp.interpolate_host(od, xp, yp, zp, tp, var, knw, sort=sort), pos.interpolate(var, knw), "pinned")
    # vec_transform off
    _same(interp.interpolate_host(od, x, y, z, t, var[2], knw[2], vec_transform=False, sort=sort),
          pos.interpolate(var[2], knw[2], vec_transform=False), "no transform")


def test_host_pipeline_on_grids_without_faces():
    """The same on a rectilinear grid (no bucket locator, vectors are two scalars) and a 2-D curvilinear box."""
    import seaduck_b200 as sd
    from seaduck_b200 import interp

    for make, n in ((ci.case_rectilinear, 300_000), (ci.case_gyre, 280_000)):
        case = make(n=n)
        od = sd.OceData(case["ds"])
        get = lambda k: case.get(k)  # noqa: E731
        pos = sd.Position().from_latlon(x=get("x"), y=get("y"), z=get("z"), t=get("t"), data=od)
        for label, var, spec, vec_transform in case["jobs"]:
            knw = ci.make_knw(sd.KnW, spec)
            ref = pos.interpolate(var, knw, vec_transform=vec_transform)
            got = interp.interpolate_host(od, get("x"), get("y"), get("z"), get("t"), var, knw, vec_transform=vec_transform)
            _same(got, ref, f"{case['name']} {label}")


def test_oceinterp_takes_the_host_pipeline_for_large_batches(monkeypatch):
    """``OceInterp`` hands batches of 2 x 10^5 points or more to ``interpolate_host``; results equal the two-step route,
    and a depth below the last level raises the reference's "Value out of bound." from there too."""
    import seaduck_b200 as sd
    from seaduck_b200 import interp

    case = ci.case_llc3d(n=250_000, ngrid=24, nz=8)
    od = sd.OceData(case["ds"])
    x, y, z = case["x"], case["y"], case["z"]
    t0 = float(case["t"][0])
    calls = []
    real = interp.interpolate_host
    monkeypatch.setattr(interp, "interpolate_host", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    got = sd.OceInterp(od, ["T", ("UVELMASS", "VVELMASS")], x, y, z, t0)
    assert calls == [1]
    ref = sd.Position().from_latlon(x=x, y=y, z=z, t=t0, data=od).interpolate(
        ["T", ("UVELMASS", "VVELMASS")], [sd.KnW(), (sd.lagrangian.uknw, sd.lagrangian.vknw)])
    _same(got, ref, "OceInterp")
    small = sd.OceInterp(od, "T", x[:100], y[:100], z[:100], t0)
    assert calls == [1] and small[0].shape == (100,)
    deep = z.copy()
    deep[123_456] = -1e5
    with pytest.raises(ValueError, match="out of bound"):
        sd.OceInterp(od, "T", x, y, deep, t0)
    with pytest.raises(ValueError, match="out of bound"):
        sd.Position().from_latlon(x=x, y=y, z=deep, t=t0, data=od)
