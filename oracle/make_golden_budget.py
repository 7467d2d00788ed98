"""TEST INFRASTRUCTURE -- golden vectors for the crossing-log post-processing of the Lagrangian budget
(reference ``seaduck/lagrangian_budget.py``: ``particle2xarray`` :297 -> ``find_ind_frac_tres`` :245 with
``which_wall`` :85, ``wall_index`` :192, ``redo_index`` :223, ``residence_time`` :139, ``tres_update``
:144), made by running the UNMODIFIED reference on Lagrangian cases of ``tests/cases.py``.

For the LLC case the budget itself is evaluated too: ``calculate_budget`` :603 (``lhs_contribution`` :575,
``contr_p_relaxed`` :583, ``prefetch_vector`` :548) on seeded random budget terms and wall concentrations
(``tests/cases.py: budget_fields``, rebuilt from the same seed by the tests), with the in-memory equivalent of what
``dump_to_zarr`` :348 stores.

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


class _Var:
    """The little of an xarray.DataArray that ``calculate_budget`` touches."""

    def __init__(self, a):
        self.values = np.asarray(a)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __sub__(self, other):
        return _Var(self.values - other)


class _Stored(dict):
    """What ``dump_to_zarr`` would have written, held in memory."""

    __getattr__ = dict.__getitem__


def stored_particle_array(neo, ind1, ind2, frac, tres):
    st = _Stored()
    for k in ("tt", "iz", "iy", "ix"):
        st[k] = _Var(np.asarray(neo[k]))
    st["face"] = _Var(np.asarray(neo["fc"]).astype("int16"))
    st["shapes"] = _Var(np.asarray(neo["shapes"]))
    st["ind1"], st["ind2"] = _Var(np.asarray(ind1).astype("int16")), _Var(np.asarray(ind2).astype("int16"))
    st["frac"], st["tres"] = _Var(frac), _Var(tres)
    return st


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
        if od.tp.typ == "LLC":
            fields = cases.budget_fields(np.asarray(case["ds"]["maskC"]).shape, seed=100 + i)
            stored = stored_particle_array(neo, ind1, ind2, frac, tres)
            for tag, kw, direction in (
                ("pad", dict(same_size=False), -1),
                ("same", dict(zname="szprime_same"), 1),
            ):
                with np.errstate(all="ignore"):
                    contr, conc, bfirst, blast = lb.calculate_budget(stored, fields, list(cases.RHS_TERMS), kw, "lhs", direction)
                for k, v in contr.items():
                    out[f"s{i}_bud_{tag}_{k}"] = np.asarray(v, float)
                out[f"s{i}_bud_{tag}_conc"] = np.asarray(conc, float)
                assert np.array_equal(bfirst, first) and np.array_equal(blast, last)
            # the consistency check between the particle log and the Eulerian budget (read_wall_list :406,
            # check_particle_data_compat :470): wall values around every record, wall velocities from
            # (u, du, r), their flux convergence, and the Eulerian convergence it should equal
            neo["face"] = neo["fc"]
            fields["sx"], fields["sy"], fields["sz"] = fields["sxprime"], fields["syprime"], fields["szprime"]
            fields["divus"] = np.random.default_rng(200 + i).normal(size=np.asarray(case["ds"]["maskC"]).shape)
            walls = [fields["sx"], fields["sy"], fields["sz"]]
            with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
                out[f"s{i}_walls_scalar"] = np.asarray(lb.read_wall_list(neo, od.tp, walls, scalar=True), float)
                out[f"s{i}_walls_vector"] = np.asarray(lb.read_wall_list(neo, od.tp, walls, scalar=False), float)
                same, extra = lb.check_particle_data_compat(neo, fields, od.tp, wall_names=("sx", "sy", "sz"), conv_name="divus", debug=True)
            out[f"s{i}_compat_ok"] = np.array(bool(same))
            for k, v in zip(("u_list", "c_list", "lagrangian_conv", "eulerian_conv"), extra):
                out[f"s{i}_compat_{k}"] = np.asarray(v, float)
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
