"""GPU: the branch-free fp64 helpers of the advection kernel (csrc/fastmath.cuh).

* fdiv must be the IEEE quotient bit for bit wherever it does not ask for the slow path, and must
  not ask for it on ordinary operands (including a zero numerator).
* flog / fexp must stay within 1 ulp of an 80-bit reference (numpy longdouble), the bound CUDA's own
  log / exp document."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIV, LOG, EXP, DIV_IEEE, LOG_CUDA, EXP_CUDA = range(6)


def run(op, a, b=None):
    import torch

    from seaduck_b200 import _lib

    dev = torch.device("cuda")
    ta = torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev)
    tb = None if b is None else torch.from_numpy(np.ascontiguousarray(b, np.float64)).to(dev)
    out = torch.empty_like(ta)
    bad = torch.empty(ta.shape, dtype=torch.int32, device=dev)
    _lib.check(
        _lib.lib().sdk_selftest_math(op, ta.numel(), ta.data_ptr(), None if tb is None else tb.data_ptr(), out.data_ptr(), bad.data_ptr(), _lib.current_stream()),
        "sdk_selftest_math",
    )
    return out.cpu().numpy(), bad.cpu().numpy().astype(bool)


def ulp_error(got, ref_ld):
    """|got - ref| in units of the last place of the (double rounded) reference."""
    ref = ref_ld.astype(np.float64)
    ulp = np.spacing(np.abs(ref)).astype(np.longdouble)
    return np.abs(got.astype(np.longdouble) - ref_ld) / ulp


def test_fdiv_is_ieee_division():
    rng = np.random.default_rng(0)
    n = 2_000_000
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    a[: n // 10] = 0.0
    a[n // 20 : n // 10] = -0.0
    a[n // 10 : n // 5] = rng.integers(-1000, 1000, n // 10)  # exact small integers
    b[n // 10 : n // 5] = rng.integers(1, 1000, n // 10)
    got, bad = run(DIV, a, b)
    with np.errstate(all="ignore"):
        ref = a / b
    assert not bad.any(), f"slow path requested for {int(bad.sum())} ordinary operand pairs"
    np.testing.assert_array_equal(got.view(np.int64), ref.view(np.int64))


def test_fdiv_flags_what_it_cannot_do():
    a = np.array([1.0, 1.0, 0.0, np.inf, np.nan, 1.0, 1e-310, 1.0, 0.0, 0.0, 1e300, 1e-300])
    b = np.array([0.0, np.inf, 0.0, 2.0, 1.0, np.nan, 3.0, 1e-310, np.inf, np.nan, 1e-300, 1e300])
    got, bad = run(DIV, a, b)
    with np.errstate(all="ignore"):
        ref = a / b
    ok = ~bad
    np.testing.assert_array_equal(got[ok].view(np.int64), ref[ok].view(np.int64))
    assert bad[[0, 1, 2, 4, 5, 6, 8, 9]].all()  # zero / inf / nan / denormal operands go to the slow path


@pytest.mark.parametrize("lo,hi", [(1e-300, 1e300), (0.25, 4.0), (1 - 1e-3, 1 + 1e-3), (1 - 1e-9, 1 + 1e-9)])
def test_flog_within_one_ulp(lo, hi):
    rng = np.random.default_rng(1)
    n = 1_000_000
    if hi / lo > 100:
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    else:
        x = rng.uniform(lo, hi, n)
    got, bad = run(LOG, x)
    assert not bad.any()
    ref = np.log(x.astype(np.longdouble))
    err = ulp_error(got, ref)
    cuda, _ = run(LOG_CUDA, x)
    err_cuda = ulp_error(cuda, ref)
    print(f"log [{lo},{hi}]: max err {err.max():.3f} ulp (CUDA log {err_cuda.max():.3f})")
    assert err.max() < 1.0


def test_flog_flags_special_inputs():
    x = np.array([0.0, -1.0, np.inf, np.nan, 1e-310, -0.0, 1.0, 2.0])
    got, bad = run(LOG, x)
    assert bad[:6].all() and not bad[6:].any()
    assert got[6] == 0.0 and got[7] == np.log(2.0)


@pytest.mark.parametrize("scale", [1e-12, 1e-6, 1e-2, 1.0, 30.0, 690.0])
def test_fexp_within_one_ulp(scale):
    rng = np.random.default_rng(2)
    x = rng.uniform(-scale, scale, 1_000_000)
    got, bad = run(EXP, x)
    assert not bad.any()
    ref = np.exp(x.astype(np.longdouble))
    err = ulp_error(got, ref)
    cuda, _ = run(EXP_CUDA, x)
    print(f"exp scale {scale}: max err {err.max():.3f} ulp (CUDA exp {ulp_error(cuda, ref).max():.3f})")
    assert err.max() < 1.0


def test_fexp_flags_out_of_range():
    x = np.array([701.0, -701.0, np.inf, -np.inf, np.nan, 0.0, -0.0])
    got, bad = run(EXP, x)
    assert bad[:5].all() and not bad[5:].any()
    assert got[5] == 1.0 and got[6] == 1.0


def test_analytic_step_on_the_reference_test_vectors():
    """One analytic in-cell step on the device (axis_wall_times -> first event -> axis_increment, the helpers of the
    advection kernel) on the vectors of the reference's own unit tests (tests/test_particle.py:75-209 there; outputs
    of the unmodified reference in tests/golden/analytic_step.npz): the event is the same one, times and positions
    agree to 1e-11 (north star 1e-9)."""
    import os

    import torch

    from seaduck_b200 import _lib

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "analytic_step.npz"))
    n = len(g["tf"])
    packed = np.concatenate([g["x0"], g["u"], g["du"], g["tf"][:, None]], axis=1)
    assert packed.shape == (n, 10)
    dev = torch.device("cuda")
    a = torch.from_numpy(np.ascontiguousarray(packed)).to(dev)
    out = torch.full((n, 8), np.nan, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().sdk_selftest_math(6, n, a.data_ptr(), None, out.data_ptr(), None, _lib.current_stream()), "sdk_selftest_math")
    out = out.cpu().numpy()
    np.testing.assert_array_equal(out[:, 0].astype(int), g["tend"])
    np.testing.assert_allclose(out[:, 1], g["t_event"], rtol=1e-11, atol=0)
    np.testing.assert_allclose(out[:, 2:5], g["newx"], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(out[:, 5:8], g["newu"], rtol=1e-11, atol=1e-25)
