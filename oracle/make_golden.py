"""TEST INFRASTRUCTURE -- generate golden vectors by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz

Every case of ``tests/cases.py`` is run through the reference's ``OceData`` / ``Particle`` with
``save_raw=True``; the crossing log (the sequence of visited cells, stamps and float state at every
analytic step) and the state after every stop are stored.  The fixtures are what pins both the CPU
oracle (``tests/test_golden_oracle.py``) and the CUDA path (``tests/test_gpu_parity.py``); the GPU
box never sees the reference itself.
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

import cases  # noqa: E402

LISTS = dict(
    fc="fclist", it="itlist", izl="izlist", rzl="rzlist", iy="iylist", ix="ixlist", rx="rxlist", ry="rylist",
    tt="ttlist", uu="uulist", vv="vvlist", ww="wwlist", du="dulist", dv="dvlist", dw="dwlist",
    xx="xxlist", yy="yylist", zz="zzlist", vs="vslist",
)


def grab_state(pt, out, prefix):
    for k in cases.FINAL_INT + cases.FINAL_FLT:
        v = getattr(pt, k, None)
        if v is not None:
            out[f"{prefix}{k}"] = np.array(v, copy=True)


def grab_log(pt, out, prefix):
    n = pt.N
    lens = np.array([len(v) for v in pt.vslist])
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    out[f"{prefix}offset"] = off
    for key, lname in LISTS.items():
        lst = getattr(pt, lname, None)
        if lst is None:
            continue
        flat = np.concatenate([np.asarray(v) for v in lst]) if off[-1] else np.zeros(0)
        out[f"{prefix}{key}"] = flat.astype(np.int32 if key in cases.LOG_INT else np.float64)


def run_case(case):
    warnings.simplefilter("ignore")
    od = sd.OceData(case["ds"])
    with contextlib.redirect_stdout(io.StringIO()):
        where = case["seeding"] if "seeding" in case else dict(x=case["x"], y=case["y"], z=case["z"], t=case["t"])
        pt = sd.Particle(data=od, save_raw=True, **where, **case["kw"])
    out = {}
    grab_state(pt, out, "init_")
    if case.get("list_of_time"):
        with contextlib.redirect_stdout(io.StringIO()):
            stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
        out["stops"] = np.array(stops, float)
        for i, snap in enumerate(snaps):
            grab_log(snap, out, f"s{i}_log_")
            grab_state(snap, out, f"s{i}_")
    else:
        out["stops"] = np.array(case["stops"], float)
        for i, ts in enumerate(case["stops"]):
            pt.empty_lists()
            pt.note_taking(stamp=15)
            pt.to_next_stop(ts)
            grab_log(pt, out, f"s{i}_log_")
            grab_state(pt, out, f"s{i}_")
    return out


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    only = sys.argv[1:]
    for make in cases.ALL_CASES + cases.ORACLE_ONLY_CASES:
        if only and make.__name__ not in only:
            continue
        case = make()
        out = run_case(case)
        path = os.path.join(gold, f"lagrangian_{case['name']}.npz")
        np.savez_compressed(path, **out)
        ncross = sum(int((out[k] < 6).sum()) for k in out if k.endswith("_log_vs"))
        print(f"{case['name']}: {len(out['stops'])} stops, {ncross} crossings -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
