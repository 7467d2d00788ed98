"""TEST INFRASTRUCTURE -- golden vectors for ``OceInterp`` (reference ``seaduck/oceinterp.py:14``), both
modes, made by running the UNMODIFIED reference on ``tests/cases_interp.case_oceinterp``.

    python oracle/make_golden_oceinterp.py        # writes tests/golden/oceinterp.npz
"""

import contextlib
import io
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refload import load_reference  # noqa: E402

sd = load_reference()

import cases_interp as ci  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    case = ci.case_oceinterp()
    od = sd.OceData(case["ds"])
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        res = sd.OceInterp(od, case["var_euler"], case["x"], case["y"], case["z"], case["t_euler"])
    for k, arr in enumerate(ci.flatten_result(res)):
        out[f"euler_{k}"] = arr
    with contextlib.redirect_stdout(io.StringIO()):
        stops, res = sd.OceInterp(
            od, case["var_lagrange"], case["x"], case["y"], case["z"], case["t_lagrange"], lagrangian=True,
            lagrange_kwarg=case["lagrange_kwarg"],
        )
    out["stops"] = np.array(stops, float)
    for k, arr in enumerate(ci.flatten_result(res)):
        out[f"lagrange_{k}"] = arr
    path = os.path.join(ROOT, "tests", "golden", "oceinterp.npz")
    np.savez_compressed(path, **out)
    print(f"oceinterp: {len(out) - 1} arrays, stops {out['stops'] / ci.DAY} d -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
