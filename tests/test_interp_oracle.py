"""CPU: the numpy/C oracle of the Eulerian path (oracle/interp_oracle.py) against the golden vectors made
by the unmodified reference (oracle/make_golden_interp.py).  This is what pins the oracle that the GPU
tests and bench_interp.py use at sizes the goldens do not cover."""

import os
import re

import numpy as np
import pytest

import cases_interp as ci

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _setup(case, gold):
    import seaduck_b200 as sd
    from oracle import interp_oracle as io

    od = sd.OceData(case["ds"])  # host side only: no device is touched
    get = lambda k: gold[f"in_{k}"] if f"in_{k}" in gold else None  # noqa: E731
    pos = io.OraclePosition(od, x=get("x"), y=get("y"), z=get("z"), t=get("t"))
    return sd, io, od, pos


@pytest.mark.parametrize("make", ci.ALL_CASES + ci.ORACLE_ONLY_CASES, ids=lambda f: f.__name__)
def test_oracle_position_matches_reference(make, oracle_lib):
    case = make()
    gold = np.load(os.path.join(GOLD, f"interp_{case['name']}.npz"))
    _, _, _, pos = _setup(case, gold)
    for key in ("face", "iy", "ix", "iz", "izl", "izl_lin", "it"):
        if f"pos_{key}" in gold:
            np.testing.assert_array_equal(getattr(pos, key), gold[f"pos_{key}"], err_msg=key)
    for key in ("rx", "ry", "rz", "rzl", "rzl_lin", "rt"):
        if f"pos_{key}" in gold:
            ci.assert_values_close(key, getattr(pos, key), gold[f"pos_{key}"])
    if "pos_iz_lin" in gold:
        np.testing.assert_array_equal(pos.lin("z")[0], gold["pos_iz_lin"])
        ci.assert_values_close("rz_lin", pos.lin("z")[1], gold["pos_rz_lin"])
    if "pos_it_lin" in gold:
        np.testing.assert_array_equal(pos.lin("t")[0], gold["pos_it_lin"])
        ci.assert_values_close("rt_lin", pos.lin("t")[1], gold["pos_rt_lin"])


@pytest.mark.parametrize("make", ci.ALL_CASES + ci.ORACLE_ONLY_CASES, ids=lambda f: f.__name__)
def test_oracle_interpolate_matches_reference(make, oracle_lib):
    case = make()
    gold = np.load(os.path.join(GOLD, f"interp_{case['name']}.npz"))
    sd, io, od, pos = _setup(case, gold)
    orc = io.InterpOracle(od, pos)
    failures = []
    for label, var, spec, vec_transform in case["jobs"]:
        knw = ci.make_knw(sd.KnW, spec)  # read as data only
        flat = ci.flatten_result(orc.interpolate(var, knw, vec_transform))
        assert len(flat) == len([k for k in gold.files if re.fullmatch(rf"job_{label}_\d+", k)]), label
        for k, arr in enumerate(flat):
            try:
                ci.assert_values_close(f"{label}[{k}]", arr, gold[f"job_{label}_{k}"], rtol=1e-10)
            except AssertionError as exc:
                failures.append(str(exc))
    assert not failures, "\n".join(failures)
