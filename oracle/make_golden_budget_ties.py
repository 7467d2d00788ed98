"""TEST INFRASTRUCTURE -- how the reference itself resolves ties of ``which_wall`` (lagrangian_budget.py:85).

``which_wall`` is an argmin over the six distances of a particle to the walls of its cell.  A particle that sits on
two walls at once (it reached a horizontal wall while gliding along a level, rzl == 1 up to rounding) is a tie that
the LAST BIT of rx / ry / rzl decides.  Those last bits come from ``log`` / ``exp``, which the reference takes from
numba's libm when ``rcParam["compilable"]`` is on and from numpy when it is off -- so the reference's own two modes
need not agree there.  This script runs the UNMODIFIED reference on ``tests/cases.py: case_llc_transport`` in both
modes (two child processes) and stores, per stop: the mask of tie records (two smallest distances closer than
1e-15), the ``ind1`` of both modes and the records where they differ.

    python oracle/make_golden_budget_ties.py        # writes tests/golden/budget_ties.npz

``tests/test_gpu_budget.py`` uses it to bound what the device may disagree on: every record outside the tie mask must
match bit for bit, and the number of mismatches inside it is printed next to the reference's own.
"""

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(numba, path):
    import contextlib
    import importlib
    import io
    import warnings

    from oracle.refload import load_reference

    sd = load_reference(numba=numba)
    lb = importlib.import_module("seaduck.lagrangian_budget")
    import cases

    warnings.simplefilter("ignore")
    case = cases.case_llc_transport()
    od = sd.OceData(case["ds"])
    with contextlib.redirect_stdout(io.StringIO()):
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    out = {"numba_active": np.array(bool(sd.runtime_conf.rcParam.get("compilable", False)) if hasattr(sd, "runtime_conf") else numba)}
    for i, ts in enumerate(case["stops"]):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        neo = lb.particle2xarray(pt)
        ind1, ind2, frac, tres, last, first = lb.find_ind_frac_tres(neo, od)
        rx, ry, rzl = (np.asarray(neo[k], float) for k in ("rx", "ry", "rz"))
        d = np.sort(np.abs(np.stack([0.5 + rx, 0.5 - rx, 0.5 + ry, 0.5 - ry, rzl, 1 - rzl])), axis=0)
        tie = (d[1] - d[0]) <= 1e-15
        tie[np.asarray(first)] = False  # first / last records use the trajectory, not the distances
        tie[np.asarray(last)] = False
        out[f"s{i}_ind1"], out[f"s{i}_tie"] = np.asarray(ind1), tie
        out[f"s{i}_shapes"] = np.asarray(neo["shapes"])
    np.savez_compressed(path, **out)


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        child(sys.argv[2] == "on", sys.argv[3])
        return
    paths = {}
    for mode in ("on", "off"):
        paths[mode] = f"/tmp/budget_ties_{mode}.npz"
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", mode, paths[mode]])
    on, off = np.load(paths["on"]), np.load(paths["off"])
    out = {}
    i = 0
    while f"s{i}_ind1" in on:
        a, b = on[f"s{i}_ind1"], off[f"s{i}_ind1"]
        same_log = a.shape == b.shape and np.array_equal(on[f"s{i}_shapes"], off[f"s{i}_shapes"])
        out[f"s{i}_same_log_shape"] = np.array(same_log)
        out[f"s{i}_tie_numba_on"] = on[f"s{i}_tie"]
        out[f"s{i}_ind1_numba_on"] = a
        if same_log:
            differ = (a != b).any(axis=0)
            out[f"s{i}_tie_numba_off"] = off[f"s{i}_tie"]
            out[f"s{i}_modes_differ"] = differ
            tie_any = on[f"s{i}_tie"] | off[f"s{i}_tie"]
            print(f"stop {i}: {a.shape[1]} records, {int(tie_any.sum())} ties, reference numba-on vs numba-off differ on "
                  f"{int(differ.sum())} records, {int((differ & ~tie_any).sum())} of them outside the tie mask")
        else:
            print(f"stop {i}: the two modes logged different numbers of records ({a.shape} vs {b.shape})")
        i += 1
    path = os.path.join(ROOT, "tests", "golden", "budget_ties.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
