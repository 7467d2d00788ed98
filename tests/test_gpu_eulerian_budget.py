"""The Eulerian-budget stencils on the device against the unmodified reference's outputs
(tests/golden/eulerian_budget.npz): bit for bit, NaN land included."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "eulerian_budget.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def eb():
    from seaduck_b200 import eulerian_budget

    return eulerian_budget


@pytest.fixture(scope="module")
def tp(gold):
    from seaduck_b200.topology import Topology

    n = gold["s"].shape[-1]
    return Topology.from_shape((13, n, n), typ="LLC")


def same(got, want):
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)  # NaN == NaN here; everything else to the bit


@pytest.mark.parametrize("face", range(13))
def test_buffers_and_dst3_per_face(gold, eb, tp, face):
    s, u = gold["s"], gold["u"]
    xb = eb.buffer_x_withface(s, face, 2, 1, tp)
    yb = eb.buffer_y_withface(s, face, 2, 1, tp)
    same(xb, gold[f"xbuf_{face}"])
    same(yb, gold[f"ybuf_{face}"])
    same(eb.third_order_DST_x(xb, u[:, face]), gold[f"dstx_{face}"])
    same(eb.third_order_DST_y(yb, u[:, face]), gold[f"dsty_{face}"])


def test_wide_and_simple_buffers(gold, eb, tp):
    same(eb.buffer_x_withface(gold["s"], 5, 3, 2, tp), gold["xbuf_wide_5"])
    same(eb.buffer_y_withface(gold["s"], 10, 1, 3, tp), gold["ybuf_wide_10"])
    same(eb.buffer_x_periodic(gold["sp"], 2, 2), gold["xper"])
    same(eb.buffer_y_periodic(gold["sp"], 1, 3), gold["yper"])
    same(eb.buffer_z_nearest(gold["sp"], 2, 1), gold["znear"])


def test_flux_limiters(gold, eb):
    same(eb.second_order_flux_limiter_x(gold["sp"], gold["ux"]), gold["fl_x"])
    same(eb.second_order_flux_limiter_y(gold["sp"], gold["uy"]), gold["fl_y"])
    same(eb.second_order_flux_limiter_z(gold["sp"], gold["uz"]), gold["fl_z"])


def test_upwind_z_and_its_side_effect(gold, eb):
    w = gold["w"].copy()
    same(eb.third_order_upwind_z(gold["sp"], w), gold["up3_z"])
    same(w, gold["w_after"])


def test_tensors_stay_on_the_device(gold, eb, tp):
    import torch

    s = torch.from_numpy(gold["s"]).cuda()
    u = torch.from_numpy(gold["u"][:, 4]).cuda()
    xb = eb.buffer_x_withface(s, 4, 2, 1, tp)
    assert isinstance(xb, torch.Tensor) and xb.is_cuda
    sx = eb.third_order_DST_x(xb, u)
    assert isinstance(sx, torch.Tensor) and sx.is_cuda
    same(sx.cpu().numpy(), gold["dstx_4"])
    w = torch.from_numpy(gold["w"]).cuda()
    sp = torch.from_numpy(gold["sp"]).cuda()
    same(eb.third_order_upwind_z(sp, w).cpu().numpy(), gold["up3_z"])
    same(w.cpu().numpy(), gold["w_after"])
    with pytest.raises(TypeError):
        eb.buffer_x_periodic(torch.zeros(3, 4), 1, 1)


def test_shape_errors_like_numpy(gold, eb):
    with pytest.raises(ValueError):
        eb.third_order_DST_x(gold["xbuf_0"], gold["u"][:, 0, :, :-1])
    with pytest.raises(ValueError):
        eb.third_order_upwind_z(gold["sp"], gold["w"][1:])


def test_llc270_size_properties(eb):
    """At LLC270 size (13 x 270 x 270 x 50): a constant field gives the constant on every wall whose stencil is
    complete; a field that is linear along the axis is reproduced exactly half-way between the cells at
    CFL 0 ... and the halo of a field holding its own flat index is the map itself."""
    import torch

    from seaduck_b200.topology import Topology

    n, nz = 270, 50
    tp = Topology.from_shape((13, n, n), typ="LLC")
    dev = torch.device("cuda", 0)
    idx = torch.arange(13 * n * n, dtype=torch.float64, device=dev).reshape(13, n, n)
    for face in (0, 2, 6, 7, 12):
        m = eb._withface_map(tp, face, "x", 2, 1)
        xb = eb.buffer_x_withface(idx, face, 2, 1, tp).cpu().numpy()
        np.testing.assert_array_equal(xb, np.where(m >= 0, m, 0).astype(float))
    const = torch.full((nz, 13, n, n), 3.25, dtype=torch.float64, device=dev)
    u = torch.rand((nz, n, n), dtype=torch.float64, device=dev) * 1.6 - 0.8
    xb = eb.buffer_x_withface(const, 1, 2, 1, tp)
    sx = eb.third_order_DST_x(xb, u)
    assert float((sx - 3.25).abs().max()) < 1e-14
    # linear in x: s = ix  ->  wall value at CFL -> 0 is the mean of the two cells around the wall
    ramp = torch.arange(n + 3, dtype=torch.float64, device=dev).expand(nz, n, n + 3).contiguous()
    tiny = torch.full((nz, n, n), 1e-300, dtype=torch.float64, device=dev)
    sx = eb.third_order_DST_x(ramp, tiny)
    want = torch.arange(n, dtype=torch.float64, device=dev) + 1.5
    assert float((sx - want).abs().max()) < 1e-12
    # the limited value never leaves the range of the two cells around the wall (TVD), 0 <= CFL < 1
    s3 = torch.randn((nz, n, n), dtype=torch.float64, device=dev)
    ux = torch.rand((nz, n, n + 1), dtype=torch.float64, device=dev) * 0.9
    fl = eb.second_order_flux_limiter_x(s3, ux)
    xb = eb.buffer_x_periodic(s3, 2, 2)
    lo = torch.minimum(xb[..., 1:-2], xb[..., 2:-1])
    hi = torch.maximum(xb[..., 1:-2], xb[..., 2:-1])
    assert bool(((fl >= lo - 1e-12) & (fl <= hi + 1e-12)).all())
