"""CPU: host-side API of ``Position`` -- ``fatten`` (reference eulerian.py:520) against the index sets the
reference produced for the golden points, and the argument checks of ``from_latlon`` / ``interpolate``
(reference tests/test_eulerian.py:178-238) that fire before any device work."""

import os
import re

import numpy as np
import pytest

import cases_interp as ci
import seaduck_b200 as sd
from seaduck_b200 import interp
from seaduck_b200.ocedata import HRel, TLinRel, TRel, VLinRel, VlLinRel, VlRel, VRel

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FATTEN_JOBS = [
    ("all", {}, {}),
    ("all_lin", dict(vkernel="linear", tkernel="linear"), {}),
    ("four_d_YX", dict(kernel="cross5"), dict(four_d=True, required=("Y", "X"))),
    ("U", ci.UKNW, dict(required=("Z", "face", "Y", "X"), four_d=True, ind_moves_kwarg={"cuvwg": "U"})),
    ("Zl", dict(vkernel="dz"), dict(required=("Zl", "Y", "X"))),
]


def _host_position(case, gold):
    """A Position with the reference's indices filled in by hand (no device involved)."""
    od = sd.OceData(case["ds"])
    pos = sd.Position()
    pos.ocedata, pos.tp = od, od.tp
    g = lambda k: gold[f"pos_{k}"] if f"pos_{k}" in gold else None  # noqa: E731
    pos.N = len(gold["pos_iy"])
    pos.rel.update(HRel(g("face"), g("iy"), g("ix"), g("rx"), g("ry"), None, None, None, None, None, None))
    pos.rel.update(VRel(g("iz"), g("rz"), None, None))
    pos.rel.update(VlRel(g("izl"), g("rzl"), None, None))
    pos.rel.update(VlLinRel(g("izl_lin"), g("rzl_lin"), None, None))
    pos.rel.update(TRel(g("it"), g("rt"), None, None))
    if g("iz_lin") is not None:
        pos.rel.update(VLinRel(g("iz_lin"), g("rz_lin"), None, None))
    if g("it_lin") is not None:
        pos.rel.update(TLinRel(g("it_lin"), g("rt_lin"), None, None))
    return pos


@pytest.mark.parametrize("make", ci.ALL_CASES, ids=lambda f: f.__name__)
def test_fatten_matches_reference(make):
    case = make()
    gold = np.load(os.path.join(GOLD, f"interp_{case['name']}.npz"))
    pos = _host_position(case, gold)
    checked = 0
    for tag, spec, kw in FATTEN_JOBS:
        want = sorted((k for k in gold.files if re.fullmatch(rf"fatten_{tag}_\d+", k)), key=lambda k: int(k.rsplit("_", 1)[1]))
        if not want:
            continue  # the reference could not do this one on this dataset
        res = pos.fatten(ci.make_knw(sd.KnW, spec), **kw)
        assert len(res) == len(want), tag
        for arr, key in zip(res, want):
            np.testing.assert_array_equal(np.asarray(arr), gold[key], err_msg=f"{tag}: {key}")
            checked += 1
    assert checked > 0


def test_from_latlon_argument_checks():
    """reference tests/test_eulerian.py:178-218 (the checks come before any lookup)."""
    ds = ci.case_gyre()["ds"]
    od = sd.OceData(ds)
    with pytest.raises(ValueError):
        sd.Position().from_latlon(x=-0.5, y=np.ones(2) * 0.1, data=None)
    with pytest.raises(ValueError):
        sd.Position().from_latlon(x=-0.5, y=np.ones(2) * 0.1, data=np.zeros(3))
    with pytest.raises(ValueError):
        sd.Position().from_latlon(x=np.ones((2, 2)), y=np.ones(2), data=od)
    with pytest.raises(ValueError):
        sd.Position().from_latlon(x=np.ones(12), y=np.ones(2), data=od)


@pytest.mark.parametrize(
    "var_name,knw,prefetched,prefix",
    [
        (1, "k", None, None),
        ("SALT", ["k", "k"], None, None),
        (("SALT", "SALT"), ("k", "k", "k"), None, None),
        ("SALT", {"s": "k"}, None, None),
        ("SALT", 1, None, None),
        ("SALT", "k", [1, 1], None),
        ("SALT", "k", None, [1, 1]),
        ("SALT", "k", 1, None),
        ("SALT", "k", None, 1),
    ],
)
def test_interpolate_argument_checks(var_name, knw, prefetched, prefix):
    """reference tests/test_eulerian.py:221-238"""
    k = sd.KnW()
    swap = lambda v: k if v == "k" else v  # noqa: E731
    if isinstance(knw, list):
        knw = [swap(v) for v in knw]
    elif isinstance(knw, tuple):
        knw = tuple(swap(v) for v in knw)
    elif isinstance(knw, dict):
        knw = {a: swap(b) for a, b in knw.items()}
    else:
        knw = swap(knw)
    with pytest.raises(ValueError):
        interp._broadcast_inputs(None, var_name, knw, prefetched, prefix)
