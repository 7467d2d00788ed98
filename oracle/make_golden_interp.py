"""TEST INFRASTRUCTURE -- golden vectors for the Eulerian interpolation path, made by running the
UNMODIFIED reference (``Position.from_latlon`` + ``Position.interpolate``, reference
``seaduck/eulerian.py:138,1268``) on the seeded cases of ``tests/cases_interp.py``.

Run in the build container (needs /root/reference):

    python oracle/make_golden_interp.py        # writes tests/golden/interp_*.npz

Points at which the reference itself raises (stencil nodes that would have to cross two faces at a
3-face junction, or cells next to an edge without neighbour) are removed first -- the reference
aborts the whole call there, so they carry no expected value; the kept points are stored in the
fixture together with the located indices and every job's output.
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


FATTEN_JOBS = [
    ("all", {}, {}),
    ("all_lin", dict(vkernel="linear", tkernel="linear"), {}),
    ("four_d_YX", dict(kernel="cross5"), dict(four_d=True, required=("Y", "X"))),
    ("U", ci.UKNW, dict(required=("Z", "face", "Y", "X"), four_d=True, ind_moves_kwarg={"cuvwg": "U"})),
    ("Zl", dict(vkernel="dz"), dict(required=("Zl", "Y", "X"))),
]


def _kernels_of(spec):
    if isinstance(spec, (tuple, list)):
        out = []
        for s in spec:
            out += _kernels_of(s)
        return out
    return [spec]


def reference_safe_points(pos, jobs):
    """Boolean mask of the points whose stencils the reference can index for every job."""
    n = pos.N
    ok = np.ones(n, bool)
    if pos.face is None:
        return ok
    tp = pos.ocedata.tp
    ny, nx = tp.h_shape[-2:]
    stencils = {}
    for _, var, spec, _ in jobs:
        for s in _kernels_of(spec):
            knw = ci.make_knw(sd.KnW, s)
            stencils[tuple(map(tuple, knw.kernel.astype(int)))] = knw
    for knw in stencils.values():
        reach = int(np.abs(knw.kernel).max())
        if reach == 0:
            continue
        near = (pos.iy < reach) | (pos.iy >= ny - reach) | (pos.ix < reach) | (pos.ix >= nx - reach)
        for i in np.where(near & ok)[0]:
            sub = pos.subset(np.array([i]))
            for cuvwg in ("C", "U", "V"):
                try:
                    with contextlib.redirect_stdout(io.StringIO()):
                        f, _, _ = sub._fatten_h(knw, ind_moves_kwarg={"cuvwg": cuvwg})
                    if cuvwg != "C":
                        pos.ocedata.tp.four_matrix_for_uv(f)
                except (IndexError, ValueError, TypeError):
                    ok[i] = False
                    break
    return ok


def run_case(case):
    warnings.simplefilter("ignore")
    od = sd.OceData(case["ds"])
    pos = sd.Position().from_latlon(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od)
    ok = reference_safe_points(pos, case["jobs"])
    keep = np.where(ok)[0]
    sel = lambda a: None if a is None else np.asarray(a)[keep]  # noqa: E731
    x, y, z, t = sel(case["x"]), sel(case["y"]), sel(case["z"]), sel(case["t"])
    pos = sd.Position().from_latlon(x=x, y=y, z=z, t=t, data=od)
    out = {"n_dropped": np.array(int((~ok).sum()))}
    for name, arr in (("x", x), ("y", y), ("z", z), ("t", t)):
        if arr is not None:
            out[f"in_{name}"] = arr
    for key in ("face", "iy", "ix", "iz", "izl", "izl_lin", "it"):
        v = getattr(pos, key, None)
        if v is not None:
            out[f"pos_{key}"] = np.asarray(v)
    for key in ("rx", "ry", "rz", "rzl", "rzl_lin", "rt"):
        v = getattr(pos, key, None)
        if v is not None:
            out[f"pos_{key}"] = np.asarray(v, float)
    for label, var, spec, vec_transform in case["jobs"]:
        knw = ci.make_knw(sd.KnW, spec)
        with contextlib.redirect_stdout(io.StringIO()):
            res = pos.interpolate(var, knw, vec_transform=vec_transform)
        for k, arr in enumerate(ci.flatten_result(res)):
            out[f"job_{label}_{k}"] = arr
    # Position.fatten (eulerian.py:520): the index sets themselves, for the host-side API
    for tag, spec, kw in FATTEN_JOBS:
        knw = ci.make_knw(sd.KnW, spec)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                res = pos.fatten(knw, **kw)
        except (IndexError, ValueError, TypeError, AttributeError, KeyError):
            continue  # e.g. a vertical dimension required on a dataset without one
        for k, arr in enumerate(res):
            out[f"fatten_{tag}_{k}"] = np.asarray(arr)
    # lazily computed linear rel-coords, after the jobs that need them ran
    for key in ("iz_lin", "it_lin"):
        try:
            out[f"pos_{key}"] = np.asarray(getattr(pos, key))
        except AttributeError:
            pass
    for key in ("rz_lin", "rt_lin"):
        try:
            out[f"pos_{key}"] = np.asarray(getattr(pos, key), float)
        except AttributeError:
            pass
    return out


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    only = sys.argv[1:]
    for make in ci.ALL_CASES + ci.ORACLE_ONLY_CASES:
        case = make()
        if only and case["name"] not in only:
            continue
        out = run_case(case)
        path = os.path.join(gold, f"interp_{case['name']}.npz")
        np.savez_compressed(path, **out)
        njob = len([k for k in out if k.startswith("job_")])
        nan = sum(int(np.isnan(out[k]).sum()) for k in out if k.startswith("job_"))
        print(
            f"{case['name']}: {len(out['in_x'])} points kept ({int(out['n_dropped'])} dropped), {njob} outputs, "
            f"{nan} NaN values -> {os.path.getsize(path) / 1e3:.0f} kB"
        )


if __name__ == "__main__":
    main()
