"""GPU: the device's cell-to-cell move (``cell_move`` of csrc/sdk_device.cuh, through ``sdk_selftest_cell_move``) against
the host ``Topology`` -- itself pinned on the reference's known answers (tests/test_topology.py) -- on EVERY cell next to a
face edge, for the 13-face LLC table, the 4-face ASTE table (reference topology.py:26, known answers
tests/test_topology.py:92-108), a closed box and a zonally periodic grid."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dataset(shape, periodic=False):
    from seaduck_b200 import minixr

    rng = np.random.default_rng(3)
    ny, nx = shape[-2:]
    faces = shape[0] if len(shape) == 3 else 1
    lon = np.linspace(0.0, 359.0 if periodic else 40.0, nx)[None, :] + np.zeros((ny, 1))
    lat = np.linspace(-40.0, 40.0, ny)[:, None] + np.zeros((1, nx))
    xc = np.stack([lon + 0.01 * rng.random((ny, nx)) for _ in range(faces)])
    yc = np.stack([lat + 0.01 * rng.random((ny, nx)) for _ in range(faces)])
    if len(shape) == 2:
        xc, yc = xc[0], yc[0]
        dims, gdims = ["Y", "X"], ["Yp1", "Xp1"]
    else:
        dims, gdims = ["face", "Y", "X"], ["face", "Yp1", "Xp1"]
    return minixr.Dataset(coords=dict(XC=(dims, xc), YC=(dims, yc), XG=(gdims, xc - 0.1), YG=(gdims, yc - 0.1)), data_vars={})


@pytest.mark.parametrize("shape,periodic,typ", [((13, 24, 24), False, "LLC"), ((4, 18, 18), False, "LLC"), ((12, 20), False, "box"), ((12, 20), True, "x_periodic")])
def test_device_cell_move_matches_host_topology(shape, periodic, typ):
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import _lib

    od = sd.OceData(_dataset(shape, periodic))
    tp = od.tp
    assert tp.typ == typ
    ny, nx = shape[-2:]
    faces = shape[0] if len(shape) == 3 else 1
    f, y, x = np.meshgrid(np.arange(faces), np.arange(ny), np.arange(nx), indexing="ij")
    rim = (y == 0) | (y == ny - 1) | (x == 0) | (x == nx - 1)
    f, y, x = f[rim], y[rim], x[rim]
    cells = np.stack([np.repeat(f, 4), np.repeat(y, 4), np.repeat(x, 4), np.tile(np.arange(4), len(f))], axis=1).astype(np.int32)
    n = len(cells)
    d_in = torch.from_numpy(cells).cuda()
    d_out = torch.empty((n, 5), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().sdk_selftest_cell_move(od.grid_handle(), n, d_in.data_ptr(), d_out.data_ptr(), _lib.current_stream()), "sdk_selftest_cell_move")
    got = d_out.cpu().numpy()
    ind = (cells[:, 0], cells[:, 1], cells[:, 2]) if len(shape) == 3 else (cells[:, 1], cells[:, 2])
    want, status = tp.ind_tend_vec(ind, cells[:, 3], swallow=True, return_status=True)
    want = np.asarray(want)
    if len(shape) == 2:
        want = np.vstack([np.zeros(n, np.int64), want])
    ok = status == 0
    if typ == "LLC":
        # where the host moves, the device moves to the same cell; where the host has no neighbour, the device flags it
        np.testing.assert_array_equal(got[ok, :3], want[:, ok].T)
        assert (got[ok, 4] == 0).all() and (got[~ok, 4] != 0).all()
        crossed = got[:, 3] >= 0
        assert (got[crossed & ok, 0] != cells[crossed & ok, 0]).all()
        assert crossed.sum() > 0 and (~ok).sum() > 0
    else:
        # closed edges: the reference returns -1 indices there (topology.py:100-146), the device freezes the particle
        stays = (want[1] >= 0) & (want[2] >= 0)
        np.testing.assert_array_equal(got[stays, 1:3], want[1:, stays].T)
        assert (got[stays, 4] == 0).all() and (got[~stays, 4] != 0).all()


def test_known_answers_of_the_reference_on_the_device():
    """reference tests/test_topology.py:73-108: LLC (n = 90) and ASTE (n = 270) moves across face edges."""
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import _lib

    for shape, table in (
        ((13, 90, 90), [((5, 89, 89), 3, (7, 0, 0)), ((5, 89, 89), 0, (6, 0, 89)), ((6, 89, 89), 0, (10, 0, 0)), ((6, 0, 0), 2, (2, 89, 89)), ((1, 45, 45), 0, (1, 46, 45))]),
        ((4, 270, 270), [((0, 269, 269), 0, (1, 0, 0)), ((1, 0, 0), 2, (0, 269, 269)), ((1, 269, 0), 0, (3, 269, 0)), ((1, 269, 269), 3, (2, 269, 0))]),
    ):
        od = sd.OceData(_dataset(shape))
        cells = np.array([list(ind) + [tend] for ind, tend, _ in table], np.int32)
        d_in = torch.from_numpy(cells).cuda()
        d_out = torch.empty((len(cells), 5), dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().sdk_selftest_cell_move(od.grid_handle(), len(cells), d_in.data_ptr(), d_out.data_ptr(), _lib.current_stream()), "cell_move")
        got = d_out.cpu().numpy()
        for row, (_, _, result) in zip(got, table):
            assert tuple(row[:3]) == result and row[4] == 0
