"""TEST INFRASTRUCTURE -- golden vectors for the crossing-log post-processing of the Lagrangian budget
(reference ``seaduck/lagrangian_budget.py``: ``particle2xarray`` :297 -> ``find_ind_frac_tres`` :245 with
``which_wall`` :85, ``wall_index`` :192, ``redo_index`` :223, ``residence_time`` :139, ``tres_update``
:144), made by running the UNMODIFIED reference on Lagrangian cases of ``tests/cases.py``.

    python oracle/make_golden_budget.py        # writes tests/golden/budget_*.npz
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
import importlib  # noqa: E402

lb = importlib.import_module("seaduck.lagrangian_budget")

import cases  # noqa: E402

BUDGET_CASES = [cases.case_llc_transport, cases.case_overturning]


def run_case(case):
    warnings.simplefilter("ignore")
    od = sd.OceData(case["ds"])
    with contextlib.redirect_stdout(io.StringIO()):
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    out = {}
    for i, ts in enumerate(case["stops"]):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        neo = lb.particle2xarray(pt)
        ind1, ind2, frac, tres, last, first = lb.find_ind_frac_tres(neo, od)
        out[f"s{i}_ind1"], out[f"s{i}_ind2"] = np.asarray(ind1), np.asarray(ind2)
        out[f"s{i}_frac"], out[f"s{i}_tres"] = np.asarray(frac, float), np.asarray(tres, float)
        out[f"s{i}_first"], out[f"s{i}_last"] = np.asarray(first), np.asarray(last)
    return out


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    for make in BUDGET_CASES:
        case = make()
        out = run_case(case)
        path = os.path.join(gold, f"budget_{case['name']}.npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']}: {out['s0_ind1'].shape} indices, tres sum {out['s0_tres'].sum():.6e} -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
