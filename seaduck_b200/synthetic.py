"""Synthetic, grid-shaped datasets for tests, golden vectors and ``bench.py``.

No ECCO / LLC4320 data can be downloaded here, so every workload in ``BASELINE.json`` is run on
synthetic fields of the same *shape* (SURVEY.md §8d):

* :func:`gyre` / :func:`overturning_cell` -- the idealised double gyre and overturning cell of the
  reference's own ``tests/test_streamfunction.py:17-92`` (no face dimension).
* :func:`llc` -- a 13-face lat-lon-cap shaped grid whose geometry is consistent with the reference's
  ``LLC_FACE_CONNECT`` table (``seaduck/topology.py:8``): every shared edge has the same physical
  wall on both faces.  Velocities come from an analytic flow projected on the cell walls, so the
  wall values are single valued across faces.
* :func:`rectilinear` -- 1-D ``lon``/``lat`` dataset for the ``rectilinear`` readiness branch.

Everything is returned as a :class:`seaduck_b200.minixr.Dataset` (or a real ``xarray.Dataset`` when
``as_xarray=True`` and xarray is importable).
"""

from __future__ import annotations

import numpy as np

from . import minixr

R_EARTH = 6371e3

__all__ = ["gyre", "overturning_cell", "llc", "llc_surface_device", "rectilinear", "llc_seed_particles"]


def _dataset(coords, data_vars, as_xarray=False):
    if as_xarray:
        import xarray as xr  # noqa: PLC0415

        return xr.Dataset(coords=coords, data_vars=data_vars)
    return minixr.Dataset(coords=coords, data_vars=data_vars)


def _streamfunction(x, z):
    return np.cos(np.pi * x / 2) * np.cos(np.pi * z / 2)


def gyre(ny=50, nx=50, dtype=np.float64, as_xarray=False):
    """Closed 2-D double gyre on ``[-1,1]^2`` degrees (reference ``tests/test_streamfunction.py:62``).

    ``UVELMASS``/``VVELMASS`` are wall *transports* derived from a corner stream function, so the
    flow is exactly non-divergent and ``streamfunc`` is conserved along trajectories.
    """
    x = np.linspace(-1, 1, nx + 1)
    y = np.linspace(-1, 1, ny + 1)
    xg, yg = np.meshgrid(x, y)
    strmf = _streamfunction(xg, yg)
    xv = 0.5 * (xg[:, 1:] + xg[:, :-1])
    yv = 0.5 * (yg[:, 1:] + yg[:, :-1])
    xc = 0.5 * (xv[1:] + xv[:-1])
    yc = 0.5 * (yv[1:] + yv[:-1])
    u = np.diff(strmf, axis=0).astype(dtype)
    v = (-np.diff(strmf, axis=1)).astype(dtype)
    return _dataset(
        dict(
            XC=(["Y", "X"], xc),
            YC=(["Y", "X"], yc),
            XG=(["Yp1", "Xp1"], xg),
            YG=(["Yp1", "Xp1"], yg),
            rA=(["Y", "X"], np.ones_like(xc, float)),
        ),
        dict(
            UVELMASS=(["Y", "Xp1"], u),
            VVELMASS=(["Yp1", "X"], v),
            streamfunc=(["Yp1", "Xp1"], strmf),
        ),
        as_xarray,
    )


def overturning_cell(nx=100, nz=50, dtype=np.float64, as_xarray=False):
    """x-z overturning cell, one cell wide in y (reference ``tests/test_streamfunction.py:17``)."""
    x = np.linspace(-1, 1, nx + 1)
    y = np.linspace(-0.1, 0.1, 2)
    zl = np.linspace(0, -1000, nz)
    zp1 = np.append(zl, -1001)
    xg, yg = np.meshgrid(x, y)
    xv = 0.5 * (xg[:, 1:] + xg[:, :-1])
    yv = 0.5 * (yg[:, 1:] + yg[:, :-1])
    xc = 0.5 * (xv[1:] + xv[:-1])
    yc = 0.5 * (yv[1:] + yv[:-1])
    tempx, tempz = np.meshgrid(x, zl)
    strmf = _streamfunction(tempx, -tempz / 500 - 1).reshape(len(zl), 1, -1)
    z = 0.5 * (zp1[1:] + zp1[:-1])
    zl = zp1[:-1]
    drf = np.abs(np.diff(zp1))
    u = np.zeros((nz, 1, nx + 1), float)
    u[:-1] = np.diff(strmf, axis=0)
    w = np.diff(strmf, axis=-1)
    v = np.zeros((nz, 2, nx), float)
    stream = np.zeros((nz, 2, nx + 1))
    stream[:] = strmf
    return _dataset(
        dict(
            XC=(["Y", "X"], xc),
            YC=(["Y", "X"], yc),
            XG=(["Yp1", "Xp1"], xg),
            YG=(["Yp1", "Xp1"], yg),
            Zl=(["Zl"], zl),
            Z=(["Z"], z),
            drF=(["Z"], drf),
            rA=(["Y", "X"], np.ones_like(xc, float)),
        ),
        dict(
            UVELMASS=(["Z", "Y", "Xp1"], u.astype(dtype)),
            VVELMASS=(["Z", "Yp1", "X"], v.astype(dtype)),
            WVELMASS=(["Zl", "Y", "X"], w.astype(dtype)),
            streamfunc=(["Zl", "Yp1", "Xp1"], stream),
        ),
        as_xarray,
    )


def channel(ny=20, nx=72, nz=4, dtype=np.float64, as_xarray=False):
    """Zonally periodic curvilinear channel (auto-detected as ``x_periodic``, reference
    ``topology.py:325-337``): lon 0..360 in ``nx`` cells, lat -50..50, ``nz`` levels, velocities in m/s."""
    xg1 = np.linspace(0.0, 360.0, nx + 1)[:-1]
    yg1 = np.linspace(-50.0, 50.0, ny + 1)
    dlon = 360.0 / nx
    xg, yg = np.meshgrid(np.append(xg1, 360.0), yg1)
    xc = 0.25 * (xg[:-1, :-1] + xg[1:, :-1] + xg[:-1, 1:] + xg[1:, 1:])
    yc = 0.25 * (yg[:-1, :-1] + yg[1:, :-1] + yg[:-1, 1:] + yg[1:, 1:])
    phi = np.deg2rad(yc)
    lam = np.deg2rad(xc)
    dxg = np.broadcast_to(R_EARTH * np.cos(np.deg2rad(yg[:-1, :-1])) * np.deg2rad(dlon), (ny, nx)).copy()
    dyg = np.full((ny, nx), R_EARTH * np.deg2rad(100.0 / ny))
    drf = np.linspace(20.0, 200.0, nz)
    zp1 = np.concatenate([[0.0], -np.cumsum(drf)])
    z = 0.5 * (zp1[1:] + zp1[:-1])
    decay = np.exp(z / 300.0)[:, None, None]
    u = decay * (0.6 * np.cos(phi) * (1 + 0.4 * np.sin(2 * lam)))[None]
    v = decay * (0.15 * np.sin(3 * lam) * np.cos(phi) ** 2)[None]
    v[:, 0, :] = 0.0  # closed southern wall; the northern one is closed by the missing row
    w = np.zeros((nz, ny, nx))
    w[1:] = 1e-5 * np.sin(lam)[None] * np.cos(phi)[None] * (np.arange(1, nz)[:, None, None] / nz)
    return _dataset(
        dict(
            XC=(["Y", "X"], xc), YC=(["Y", "X"], yc), XG=(["Yp1", "Xp1"], xg[:, :]), YG=(["Yp1", "Xp1"], yg[:, :]),
            dxG=(["Yp1", "X"], np.vstack([dxg, dxg[-1:]])), dyG=(["Y", "Xp1"], np.hstack([dyg, dyg[:, -1:]])),
            rA=(["Y", "X"], dxg * dyg), Z=(["Z"], z), Zl=(["Zl"], zp1[:-1]), Zp1=(["Zp1"], zp1), drF=(["Z"], drf),
        ),
        dict(
            UVELMASS=(["Z", "Y", "X"], u.astype(dtype)), VVELMASS=(["Z", "Y", "X"], v.astype(dtype)),
            WVELMASS=(["Zl", "Y", "X"], w.astype(dtype)),
        ),
        as_xarray,
    )


def rectilinear(nlat=60, nlon=120, nt=1, dtype=np.float64, as_xarray=False):
    """Global 1-D lon/lat grid (``readiness['h'] == 'rectilinear'``, reference ``ocedata.py:255``): dimensions
    ``Y`` / ``X`` (the names ``Position.fatten`` knows, ``eulerian.py:520``) with the 1-D variables ``lon(X)``,
    ``lat(Y)`` and unstaggered velocities ``u, v`` in m/s -- the shape of an altimetry product."""
    lon = np.linspace(-180, 180, nlon, endpoint=False)
    lat = np.linspace(-75, 75, nlat)
    lam, phi = np.meshgrid(np.deg2rad(lon), np.deg2rad(lat))
    u = 0.4 * np.cos(phi) * (1 + 0.5 * np.sin(3 * lam))
    v = 0.12 * np.sin(2 * lam) * np.cos(phi) ** 2
    coords = dict(X=(["X"], np.arange(nlon)), Y=(["Y"], np.arange(nlat)))
    data = dict(lon=(["X"], lon), lat=(["Y"], lat))
    if nt > 1:
        time = np.arange(nt) * 86400.0 * 5
        coords["time"] = (["time"], time)
        scale = (1 + 0.25 * np.arange(nt)).reshape(nt, 1, 1)
        data.update(u=(["time", "Y", "X"], (u * scale).astype(dtype)), v=(["time", "Y", "X"], (v * scale).astype(dtype)))
    else:
        data.update(u=(["Y", "X"], u.astype(dtype)), v=(["Y", "X"], v.astype(dtype)))
    return _dataset(coords, data, as_xarray)


# ----------------------------------------------------------------------------------------------
# LLC-shaped grid
# ----------------------------------------------------------------------------------------------
_LAT_S, _LAT_N = -78.0, 66.0
_SECTOR_W = (-38.0, 52.0, 142.0, 232.0)  # west edge of the four 90 degree sectors
_CAP_LON0 = 97.0


def _llc_face_lonlat(face, fy, fx, n):
    """lon/lat (deg) of the point with real-valued index ``(fy, fx)`` on ``face``.

    Integer indices are G (corner) points, ``+0.5`` gives C (centre) points.  The layout follows
    SURVEY.md §8d and was checked edge by edge against ``LLC_FACE_CONNECT``.
    """
    fy = np.asarray(fy, float)
    fx = np.asarray(fx, float)
    dlat = (_LAT_N - _LAT_S) / (3 * n)
    dlon = 90.0 / n
    if face in (0, 1, 2, 3, 4, 5):
        sector = 0 if face < 3 else 1
        row = face % 3  # 0 = southern-most
        lon = _SECTOR_W[sector] + fx * dlon
        lat = _LAT_S + (row * n + fy) * dlat
    elif face in (7, 8, 9, 10, 11, 12):
        sector = 2 if face < 10 else 3
        row = (face - 7) % 3  # 0 = northern-most (touches the cap)
        lon = _SECTOR_W[sector] + fy * dlon
        lat = _LAT_N - (row * n + fx) * dlat
    elif face == 6:
        a = (fx - n / 2) / (n / 2)
        b = (fy - n / 2) / (n / 2)
        r = np.maximum(np.abs(a), np.abs(b))
        rs = np.where(r == 0, 1.0, r)
        bottom = (b <= -np.abs(a)) & (r > 0)
        right = (a >= np.abs(b)) & ~bottom & (r > 0)
        top = (b >= np.abs(a)) & ~bottom & ~right & (r > 0)
        angle = np.where(
            bottom,
            45 * a / rs,
            np.where(right, 90 + 45 * b / rs, np.where(top, 180 - 45 * a / rs, 270 - 45 * b / rs)),
        )
        angle = np.where(r == 0, 0.0, angle)
        lon = _CAP_LON0 + angle
        lat = 90.0 - (90.0 - _LAT_N) * r
    else:
        raise ValueError(face)
    lon = (lon + 180.0) % 360.0 - 180.0
    return lon, lat


def _sph2cart(lon, lat):
    lam = np.deg2rad(lon)
    phi = np.deg2rad(lat)
    return np.stack([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)], axis=-1)


def _flow_cart(p, amp_rot=1.0, amp_div=0.0):
    """Analytic horizontal velocity (m/s, cartesian components) at unit vectors ``p``.

    rotational part: ``ue = cos(phi)(1 + 0.5 sin 3 lam)``, ``vn = 0.3 sin(2 lam) cos^2(phi)``
    divergent part : ``ue = 0.5 cos(phi) cos(2 lam)``,     ``vn = 0.5 cos^2(phi) sin(3 phi)``
    both vanish at the poles.
    """
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    lam = np.arctan2(y, x)
    cphi = np.sqrt(np.maximum(x * x + y * y, 1e-30))
    phi = np.arctan2(z, cphi)
    ue = amp_rot * cphi * (1 + 0.5 * np.sin(3 * lam)) + amp_div * 0.5 * cphi * np.cos(2 * lam)
    vn = amp_rot * 0.3 * np.sin(2 * lam) * cphi**2 + amp_div * 0.5 * cphi**2 * np.sin(3 * phi)
    east = np.stack([-np.sin(lam), np.cos(lam), np.zeros_like(lam)], axis=-1)
    north = np.stack([-z * np.cos(lam), -z * np.sin(lam), cphi], axis=-1)
    return ue[..., None] * east + vn[..., None] * north


def _llc_land(n, land_blobs=False, xc=None, yc=None):
    """maskC for one level: 1 = wet.  Land where the reference cannot index."""
    k = max(2, n // 22)  # 4 for n = 90
    mask = np.ones((13, n, n), np.int8)
    mask[[0, 3], :2, :] = 0  # southern rows (no neighbour below faces 0 and 3)
    mask[[9, 12], :, n - 2 :] = 0  # right columns of faces 9, 12 (no neighbour)
    # the four corners of the cap and the touching corners of the neighbouring faces
    for sy in (slice(0, k), slice(n - k, n)):
        for sx in (slice(0, k), slice(n - k, n)):
            mask[6, sy, sx] = 0
    mask[2, n - k :, :k] = 0
    mask[2, n - k :, n - k :] = 0
    mask[5, n - k :, :k] = 0
    mask[5, n - k :, n - k :] = 0
    for f in (7, 10):
        mask[f, :k, :k] = 0
        mask[f, n - k :, :k] = 0
    c = n // 2
    mask[6, c - k // 2 - 1 : c + k // 2 + 1, c - k // 2 - 1 : c + k // 2 + 1] = 0  # pole
    if land_blobs:
        lam = np.deg2rad(xc)
        phi = np.deg2rad(yc)
        mask[np.sin(5 * lam) * np.cos(7 * phi) > 0.6] = 0
    return mask


def llc(
    n=90,
    nz=50,
    nt=0,
    u0=0.5,
    dtype=np.float64,
    land_blobs=False,
    tracers=False,
    as_xarray=False,
    seed=0,
):
    """LLC-shaped synthetic dataset: 13 faces x n x n, ``nz`` levels (0 = 2-D), ``nt`` snapshots (0 = static).

    Variables: ``XC YC XG YG`` (XG/YG have size ``n`` like ECCO, so ``g_shape == h_shape``),
    ``dxG dyG rA CS SN``, and for ``nz > 0`` ``Z Zl Zp1 drF drC hFacC maskC``;
    velocities ``UVELMASS VVELMASS WVELMASS`` (m/s on the walls) and transports
    ``utrans vtrans wtrans`` (m^3/s, with W from continuity so the 3-D flow is volume conserving
    away from land); ``T``/``S`` tracers if requested.
    """
    rng = np.random.default_rng(seed)
    idx = np.arange(n + 1, dtype=float)
    gy, gx = np.meshgrid(idx, idx, indexing="ij")  # (n+1, n+1) corner indices
    cy, cx = gy[:n, :n] + 0.5, gx[:n, :n] + 0.5

    XG = np.zeros((13, n, n))
    YG = np.zeros((13, n, n))
    XC = np.zeros((13, n, n))
    YC = np.zeros((13, n, n))
    dxG = np.zeros((13, n, n))
    dyG = np.zeros((13, n, n))
    rA = np.zeros((13, n, n))
    CS = np.zeros((13, n, n))
    SN = np.zeros((13, n, n))
    # unit tangent x wall-length products, kept to project the analytic flow on the walls
    u_nrm = np.zeros((13, n, n, 3))  # normal of the west wall (towards +ix), unit
    v_nrm = np.zeros((13, n, n, 3))  # normal of the south wall (towards +iy), unit
    u_mid = np.zeros((13, n, n, 3))
    v_mid = np.zeros((13, n, n, 3))
    for f in range(13):
        lon_g, lat_g = _llc_face_lonlat(f, gy, gx, n)
        lon_c, lat_c = _llc_face_lonlat(f, cy, cx, n)
        pg = _sph2cart(lon_g, lat_g)  # (n+1,n+1,3)
        XG[f], YG[f] = lon_g[:n, :n], lat_g[:n, :n]
        XC[f], YC[f] = lon_c, lat_c
        tx = pg[:n, 1:] - pg[:n, :n]  # south wall: G(iy,ix) -> G(iy,ix+1)
        ty = pg[1:, :n] - pg[:n, :n]  # west wall:  G(iy,ix) -> G(iy+1,ix)
        dxG[f] = np.linalg.norm(tx, axis=-1) * R_EARTH
        dyG[f] = np.linalg.norm(ty, axis=-1) * R_EARTH
        # cell area from the two triangles of the quadrilateral
        d1 = pg[1:, 1:] - pg[:n, :n]
        a1 = 0.5 * np.linalg.norm(np.cross(tx, d1), axis=-1)
        a2 = 0.5 * np.linalg.norm(np.cross(d1, ty), axis=-1)
        rA[f] = (a1 + a2) * R_EARTH**2
        mu = pg[:n, :n] + 0.5 * ty
        mu /= np.linalg.norm(mu, axis=-1, keepdims=True)
        mv = pg[:n, :n] + 0.5 * tx
        mv /= np.linalg.norm(mv, axis=-1, keepdims=True)
        nu = np.cross(ty, mu)  # y_local x up = x_local
        nv = np.cross(mv, tx)  # up x x_local = y_local
        with np.errstate(invalid="ignore", divide="ignore"):
            nu /= np.linalg.norm(nu, axis=-1, keepdims=True)
            nv /= np.linalg.norm(nv, axis=-1, keepdims=True)
        u_nrm[f], v_nrm[f] = np.nan_to_num(nu), np.nan_to_num(nv)
        u_mid[f], v_mid[f] = mu, mv
        # orientation of local x relative to east at the centre
        pc = _sph2cart(lon_c, lat_c)
        xdir = 0.5 * (pg[:n, 1:] + pg[1:, 1:]) - 0.5 * (pg[:n, :n] + pg[1:, :n])
        lam = np.deg2rad(lon_c)
        phi = np.deg2rad(lat_c)
        east = np.stack([-np.sin(lam), np.cos(lam), np.zeros_like(lam)], axis=-1)
        north = np.stack(
            [-np.sin(phi) * np.cos(lam), -np.sin(phi) * np.sin(lam), np.cos(phi)], axis=-1
        )
        ce = (xdir * east).sum(-1)
        cn = (xdir * north).sum(-1)
        nrm = np.hypot(ce, cn)
        nrm[nrm == 0] = 1.0
        CS[f], SN[f] = ce / nrm, cn / nrm
        del pc

    land2d = _llc_land(n, land_blobs, XC, YC)

    coords = dict(
        XC=(["face", "Y", "X"], XC),
        YC=(["face", "Y", "X"], YC),
        XG=(["face", "Yp1", "Xp1"], XG),
        YG=(["face", "Yp1", "Xp1"], YG),
        dxG=(["face", "Yp1", "X"], dxG),
        dyG=(["face", "Y", "Xp1"], dyG),
        rA=(["face", "Y", "X"], rA),
        CS=(["face", "Y", "X"], CS),
        SN=(["face", "Y", "X"], SN),
    )

    # wet-wall masks: a wall carries flow only if the cells on both sides are wet
    from .topology import Topology  # noqa: PLC0415  (host-side index logic of the product)

    tp = Topology.from_shape((13, n, n))
    f_i, y_i, x_i = np.meshgrid(np.arange(13), np.arange(n), np.arange(n), indexing="ij")
    flat = (f_i.ravel(), y_i.ravel(), x_i.ravel())
    lf, ly, lx = tp.ind_tend_vec(flat, np.full(f_i.size, 2), swallow=True)  # cell to the left
    df, dy_, dx_ = tp.ind_tend_vec(flat, np.full(f_i.size, 1), swallow=True)  # cell below
    wet_u = (land2d * land2d[lf, ly, lx].reshape(13, n, n)).astype(float)
    wet_v = (land2d * land2d[df, dy_, dx_].reshape(13, n, n)).astype(float)
    # cells without a left / lower neighbour keep their own index -> treat as closed wall
    no_left = (lf.reshape(13, n, n) == f_i) & (lx.reshape(13, n, n) == x_i)
    no_down = (df.reshape(13, n, n) == f_i) & (dy_.reshape(13, n, n) == y_i)
    wet_u[no_left] = 0.0
    wet_v[no_down] = 0.0

    vel_u_rot = (_flow_cart(u_mid, 1.0, 0.0) * u_nrm).sum(-1) * wet_u  # (13,n,n) per unit amplitude
    vel_v_rot = (_flow_cart(v_mid, 1.0, 0.0) * v_nrm).sum(-1) * wet_v
    vel_u_div = (_flow_cart(u_mid, 0.0, 1.0) * u_nrm).sum(-1) * wet_u
    vel_v_div = (_flow_cart(v_mid, 0.0, 1.0) * v_nrm).sum(-1) * wet_v

    data = {}
    if nz > 0:
        drF = np.linspace(10.0, 450.0, nz)
        Zp1 = np.concatenate([[0.0], -np.cumsum(drF)])
        Z = 0.5 * (Zp1[1:] + Zp1[:-1])
        Zl = Zp1[:-1]
        drC = np.concatenate([[Z[0] - 0.0], Z[:-1] - Z[1:]]) * -1.0
        drC = np.abs(drC)
        maskC = np.broadcast_to(land2d, (nz, 13, n, n)).copy()
        g_rot = np.exp(Z / 1000.0)
        g_div = np.cos(np.pi * (np.arange(nz) + 0.5) / nz)
        g_div = g_div - (g_div * drF).sum() / drF.sum()  # zero vertical integral -> W(surface) = 0
        g_div *= 0.3
        uvel = u0 * (g_rot[:, None, None, None] * vel_u_rot + g_div[:, None, None, None] * vel_u_div)
        vvel = u0 * (g_rot[:, None, None, None] * vel_v_rot + g_div[:, None, None, None] * vel_v_div)
        utr = uvel * drF[:, None, None, None] * dyG
        vtr = vvel * drF[:, None, None, None] * dxG
        # east / north wall transports of every cell (across faces) -> W from continuity
        rf, ry, rx = tp.wall_source(flat, "right")
        uf, uy, ux = tp.wall_source(flat, "top")
        r_is_v = tp.wall_source_is_other(flat, "right").reshape(13, n, n)
        t_is_u = tp.wall_source_is_other(flat, "top").reshape(13, n, n)
        east = np.where(r_is_v, vtr[:, rf, ry, rx].reshape(nz, 13, n, n), utr[:, rf, ry, rx].reshape(nz, 13, n, n))
        nrth = np.where(t_is_u, utr[:, uf, uy, ux].reshape(nz, 13, n, n), vtr[:, uf, uy, ux].reshape(nz, 13, n, n))
        closed_r = tp.wall_source_missing(flat, "right").reshape(13, n, n)
        closed_t = tp.wall_source_missing(flat, "top").reshape(13, n, n)
        east = np.where(closed_r, 0.0, east)
        nrth = np.where(closed_t, 0.0, nrth)
        divh = (east - utr) + (nrth - vtr)
        wtr = np.zeros((nz, 13, n, n))
        acc = np.zeros((13, n, n))
        for k in range(nz - 1, -1, -1):  # W(bottom) = 0, integrate upwards
            acc = acc - divh[k]
            wtr[k] = acc
        wtr *= maskC
        wvel = wtr / np.where(rA > 0, rA, 1.0)
        coords.update(
            Z=(["Z"], Z),
            Zl=(["Zl"], Zl),
            Zp1=(["Zp1"], Zp1),
            drF=(["Z"], drF),
            drC=(["Z"], drC),
            hFacC=(["Z", "face", "Y", "X"], maskC.astype(float)),
            maskC=(["Z", "face", "Y", "X"], maskC),
        )
        udims, vdims, wdims = (
            ["Z", "face", "Y", "Xp1"],
            ["Z", "face", "Yp1", "X"],
            ["Zl", "face", "Y", "X"],
        )
        cdims = ["Z", "face", "Y", "X"]
        fields = dict(
            UVELMASS=(udims, uvel),
            VVELMASS=(vdims, vvel),
            WVELMASS=(wdims, wvel),
            utrans=(udims, utr),
            vtrans=(vdims, vtr),
            wtrans=(wdims, wtr),
        )
        if tracers:
            lam = np.deg2rad(XC)
            phi = np.deg2rad(YC)
            zz = Z[:, None, None, None]
            fields["T"] = (cdims, 2.0 + 25.0 * np.cos(phi) ** 2 * np.exp(zz / 800.0) + np.sin(2 * lam))
            fields["S"] = (cdims, 34.5 + 0.8 * np.sin(phi) * np.cos(lam) + zz / 5000.0)
    else:
        uvel = u0 * (vel_u_rot + 0.0 * vel_u_div)
        vvel = u0 * (vel_v_rot + 0.0 * vel_v_div)
        udims, vdims = ["face", "Y", "Xp1"], ["face", "Yp1", "X"]
        fields = dict(
            UVELMASS=(udims, uvel),
            VVELMASS=(vdims, vvel),
            utrans=(udims, uvel * dyG),
            vtrans=(vdims, vvel * dxG),
        )
        if tracers:
            lam = np.deg2rad(XC)
            phi = np.deg2rad(YC)
            fields["T"] = (["face", "Y", "X"], 2.0 + 25.0 * np.cos(phi) ** 2 + np.sin(2 * lam))

    if nt > 0:
        time = np.arange(nt) * 86400.0 * 10
        coords["time"] = (["time"], time)
        scale = 1.0 + 0.3 * np.sin(np.arange(nt) * 1.3)
        for name, (dims, arr) in fields.items():
            wob = scale.reshape((nt,) + (1,) * arr.ndim)
            if name in ("T", "S"):
                wob = 1.0 + 0.01 * (wob - 1.0)
            data[name] = (["time"] + dims, (arr[None] * wob).astype(dtype))
    else:
        for name, (dims, arr) in fields.items():
            data[name] = (dims, arr.astype(dtype))
    del rng
    return _dataset(coords, data, as_xarray)


def llc_seed_particles(ds, num, seed=1, lat_range=(-70.0, 85.0), depth_range=(5.0, 1000.0), wet_only=True):
    """``num`` particle positions uniform on the sphere (normal-vector method), SURVEY.md §8d."""
    rng = np.random.default_rng(seed)
    lon = np.empty(0)
    lat = np.empty(0)
    while len(lon) < num:
        v = rng.normal(size=(2 * num + 16, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        la = np.rad2deg(np.arcsin(v[:, 2]))
        lo = np.rad2deg(np.arctan2(v[:, 1], v[:, 0]))
        ok = (la > lat_range[0]) & (la < lat_range[1])
        lon = np.concatenate([lon, lo[ok]])
        lat = np.concatenate([lat, la[ok]])
    lon, lat = lon[:num], lat[:num]
    dep = -rng.uniform(depth_range[0], depth_range[1], num)
    del wet_only, ds
    return lon, lat, dep


# ----------------------------------------------------------------------------------------------
# the same LLC surface dataset, built on the GPU (BASELINE.json configs[3]: LLC4320 x 2 snapshots)
# ----------------------------------------------------------------------------------------------
def _t_face_lonlat(torch, face, fy, fx, n):
    """torch restatement of :func:`_llc_face_lonlat` (same formulas, float64 on the device)."""
    dlat = (_LAT_N - _LAT_S) / (3 * n)
    dlon = 90.0 / n
    if face < 6:
        lon = _SECTOR_W[0 if face < 3 else 1] + fx * dlon
        lat = _LAT_S + ((face % 3) * n + fy) * dlat
    elif face > 6:
        lon = _SECTOR_W[2 if face < 10 else 3] + fy * dlon
        lat = _LAT_N - (((face - 7) % 3) * n + fx) * dlat
    else:
        a = (fx - n / 2) / (n / 2)
        b = (fy - n / 2) / (n / 2)
        r = torch.maximum(a.abs(), b.abs())
        rs = torch.where(r == 0, torch.ones_like(r), r)
        bottom = (b <= -a.abs()) & (r > 0)
        right = (a >= b.abs()) & ~bottom & (r > 0)
        top = (b >= a.abs()) & ~bottom & ~right & (r > 0)
        angle = torch.where(bottom, 45 * a / rs, torch.where(right, 90 + 45 * b / rs, torch.where(top, 180 - 45 * a / rs, 270 - 45 * b / rs)))
        angle = torch.where(r == 0, torch.zeros_like(angle), angle)
        lon = _CAP_LON0 + angle
        lat = 90.0 - (90.0 - _LAT_N) * r
    lon, lat = torch.broadcast_tensors(lon, lat)
    lon = torch.remainder(lon + 180.0, 360.0) - 180.0
    return lon, lat


def _t_sph2cart(torch, lon, lat):
    lam, phi = torch.deg2rad(lon), torch.deg2rad(lat)
    return torch.stack([torch.cos(phi) * torch.cos(lam), torch.cos(phi) * torch.sin(lam), torch.sin(phi)], dim=-1)


def _t_flow_rot(torch, p):
    """Rotational part of :func:`_flow_cart` at unit vectors ``p`` (unit amplitude)."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    lam = torch.atan2(y, x)
    cphi = torch.sqrt(torch.clamp(x * x + y * y, min=1e-30))
    ue = cphi * (1 + 0.5 * torch.sin(3 * lam))
    vn = 0.3 * torch.sin(2 * lam) * cphi**2
    east = torch.stack([-torch.sin(lam), torch.cos(lam), torch.zeros_like(lam)], dim=-1)
    north = torch.stack([-z * torch.cos(lam), -z * torch.sin(lam), cphi], dim=-1)
    return ue[..., None] * east + vn[..., None] * north


def llc_surface_device(n=4320, nt=3, u0=0.5, dtype=np.float64, device="cuda"):
    """:func:`llc` with ``nz = 0`` for grids too large to go through host memory in float64: the geometry and the
    analytic flow are evaluated face by face on the GPU.  Returns ``(ds, fields)``: ``ds`` holds the float32 grid
    (``XC YC XG YG dxG dyG rA CS SN`` -- the precision ``OceData`` keeps them in -- and ``time``), ``fields`` maps
    ``utrans`` / ``vtrans`` to ``(dims, device tensor [nt, 13, n, n])`` for ``OceData.stage_device_field``.
    Same formulas as :func:`llc` (checked against it at small ``n`` in ``tests/test_gpu_streamfunction.py``).
    """
    import torch  # noqa: PLC0415

    from .topology import Topology  # noqa: PLC0415

    dev = torch.device(device)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    idx = torch.arange(n + 1, dtype=torch.float64, device=dev)
    gy, gx = idx[:, None], idx[None, :]
    cy, cx = gy[:n] + 0.5, gx[:, :n] + 0.5
    names = ("XC", "YC", "XG", "YG", "dxG", "dyG", "rA", "CS", "SN")
    host = {k: np.empty((13, n, n), np.float32) for k in names}
    land = torch.from_numpy(_llc_land(n)).to(dev)
    # land flag of the cell to the left of column 0 / below row 0 of every face (across faces), from the host topology
    tp = Topology.from_shape((13, n, n))
    f_e = np.repeat(np.arange(13), n)
    k_e = np.tile(np.arange(n), 13)
    zeros = np.zeros_like(k_e)
    lf, ly, lx = tp.ind_tend_vec((f_e, k_e, zeros), np.full(f_e.size, 2), swallow=True)
    df, dy_, dx_ = tp.ind_tend_vec((f_e, zeros, k_e), np.full(f_e.size, 1), swallow=True)
    no_left = torch.from_numpy(((lf == f_e) & (lx == 0)).reshape(13, n)).to(dev)
    no_down = torch.from_numpy(((df == f_e) & (dy_ == 0)).reshape(13, n)).to(dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    left_land = land[t(lf), t(ly), t(lx)].reshape(13, n)
    down_land = land[t(df), t(dy_), t(dx_)].reshape(13, n)
    scale = 1.0 + 0.3 * np.sin(np.arange(max(nt, 1)) * 1.3)
    shape = (nt, 13, n, n) if nt > 0 else (13, n, n)
    utr = torch.empty(shape, dtype=tdt, device=dev)
    vtr = torch.empty(shape, dtype=tdt, device=dev)
    for f in range(13):
        lon_g, lat_g = _t_face_lonlat(torch, f, gy, gx, n)
        lon_c, lat_c = _t_face_lonlat(torch, f, cy, cx, n)
        pg = _t_sph2cart(torch, lon_g, lat_g)
        tx = pg[:n, 1:] - pg[:n, :n]
        ty = pg[1:, :n] - pg[:n, :n]
        dxg = torch.linalg.norm(tx, dim=-1) * R_EARTH
        dyg = torch.linalg.norm(ty, dim=-1) * R_EARTH
        d1 = pg[1:, 1:] - pg[:n, :n]
        area = (0.5 * torch.linalg.norm(torch.linalg.cross(tx, d1), dim=-1) + 0.5 * torch.linalg.norm(torch.linalg.cross(d1, ty), dim=-1)) * R_EARTH**2
        mu = pg[:n, :n] + 0.5 * ty
        mu = mu / torch.linalg.norm(mu, dim=-1, keepdim=True)
        mv = pg[:n, :n] + 0.5 * tx
        mv = mv / torch.linalg.norm(mv, dim=-1, keepdim=True)
        nu = torch.linalg.cross(ty, mu)
        nv = torch.linalg.cross(mv, tx)
        nu = torch.nan_to_num(nu / torch.linalg.norm(nu, dim=-1, keepdim=True))
        nv = torch.nan_to_num(nv / torch.linalg.norm(nv, dim=-1, keepdim=True))
        xdir = 0.5 * (pg[:n, 1:] + pg[1:, 1:]) - 0.5 * (pg[:n, :n] + pg[1:, :n])
        lam, phi = torch.deg2rad(lon_c), torch.deg2rad(lat_c)
        east = torch.stack([-torch.sin(lam), torch.cos(lam), torch.zeros_like(lam)], dim=-1)
        north = torch.stack([-torch.sin(phi) * torch.cos(lam), -torch.sin(phi) * torch.sin(lam), torch.cos(phi)], dim=-1)
        ce, cn = (xdir * east).sum(-1), (xdir * north).sum(-1)
        nrm = torch.hypot(ce, cn)
        nrm = torch.where(nrm == 0, torch.ones_like(nrm), nrm)
        for key, val in (("XC", lon_c), ("YC", lat_c), ("XG", lon_g[:n, :n]), ("YG", lat_g[:n, :n]), ("dxG", dxg), ("dyG", dyg),
                         ("rA", area), ("CS", ce / nrm), ("SN", cn / nrm)):
            host[key][f] = val.to(torch.float32).cpu().numpy()
        # a wall carries flow only if the cells on both sides are wet
        own = land[f].to(torch.float64)
        wet_u = own.clone()
        wet_u[:, 1:] *= own[:, :-1]
        wet_u[:, 0] *= torch.where(no_left[f], torch.zeros_like(own[:, 0]), left_land[f].to(torch.float64))
        wet_v = own.clone()
        wet_v[1:, :] *= own[:-1, :]
        wet_v[0, :] *= torch.where(no_down[f], torch.zeros_like(own[0, :]), down_land[f].to(torch.float64))
        u_face = u0 * (_t_flow_rot(torch, mu) * nu).sum(-1) * wet_u * dyg
        v_face = u0 * (_t_flow_rot(torch, mv) * nv).sum(-1) * wet_v * dxg
        if nt > 0:
            for k in range(nt):
                utr[k, f] = (u_face * scale[k]).to(tdt)
                vtr[k, f] = (v_face * scale[k]).to(tdt)
        else:
            utr[f], vtr[f] = u_face.to(tdt), v_face.to(tdt)
        del pg, tx, ty, d1, mu, mv, nu, nv, xdir, east, north
    coords = dict(
        XC=(["face", "Y", "X"], host["XC"]), YC=(["face", "Y", "X"], host["YC"]),
        XG=(["face", "Yp1", "Xp1"], host["XG"]), YG=(["face", "Yp1", "Xp1"], host["YG"]),
        dxG=(["face", "Yp1", "X"], host["dxG"]), dyG=(["face", "Y", "Xp1"], host["dyG"]),
        rA=(["face", "Y", "X"], host["rA"]), CS=(["face", "Y", "X"], host["CS"]), SN=(["face", "Y", "X"], host["SN"]),
    )
    lead = ["time"] if nt > 0 else []
    if nt > 0:
        coords["time"] = (["time"], np.arange(nt) * 86400.0 * 10)
    fields = {"utrans": (lead + ["face", "Y", "Xp1"], utr), "vtrans": (lead + ["face", "Yp1", "X"], vtr)}
    return minixr.Dataset(coords=coords, data_vars={}), fields
