"""GPU: crossing-log post-processing of the Lagrangian budget (sdk_log_budget) against golden vectors
from the unmodified reference (oracle/make_golden_budget.py: particle2xarray -> find_ind_frac_tres)."""

import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("make", [cases.case_llc_transport, cases.case_overturning], ids=lambda f: f.__name__)
def test_budget_postprocessing_matches_reference(make):
    import seaduck_b200 as sd
    from seaduck_b200 import lagrangian_budget as lb

    case = make()
    gold = np.load(os.path.join(GOLD, f"budget_{case['name']}.npz"))
    od = sd.OceData(case["ds"])
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    for i, ts in enumerate(case["stops"]):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        ind1, ind2, frac, tres, last, first = lb.find_ind_frac_tres(pt, od)
        np.testing.assert_array_equal(first, gold[f"s{i}_first"])
        np.testing.assert_array_equal(last, gold[f"s{i}_last"])
        # which_wall is an argmin over six distances: a particle that sits on two walls at once (on a
        # horizontal wall while gliding along the surface, rzl == 1) is a tie that the last bit of rx
        # decides -- in the reference as well.  Compare where the nearest wall is unambiguous.
        rec = pt.raw_arrays()[-1]
        d = np.sort(np.abs(np.stack([0.5 + rec["rx"], 0.5 - rec["rx"], 0.5 + rec["ry"], 0.5 - rec["ry"], rec["rzl"], 1 - rec["rzl"]])), axis=0)
        clear = (d[1] - d[0]) > 1e-15
        clear[first] = True  # first / last records use the trajectory, not the distances
        clear[last] = True
        np.testing.assert_array_equal(ind1[:, clear], gold[f"s{i}_ind1"][:, clear], err_msg="ind1")
        same_on_ties = (ind1[:, ~clear] == gold[f"s{i}_ind1"][:, ~clear]).all(axis=0)
        # The tie records are the ones the reference itself does not resolve reproducibly: tests/golden/budget_ties.npz
        # (oracle/make_golden_budget_ties.py) holds what its numba-on and numba-off runs -- same code, libm of numba vs
        # numpy -- give for this case, and how often THEY differ.  The device may differ from the golden (numba on) on
        # tie records only, and not more often than twice the reference differs from itself (+2).
        n_tie, n_diff = int((~clear).sum()), int((~same_on_ties).sum())
        ref_self = 0
        if case["name"] == "llc_transport":
            ties = np.load(os.path.join(GOLD, "budget_ties.npz"))
            if bool(ties[f"s{i}_same_log_shape"]):
                ref_self = int(ties[f"s{i}_modes_differ"].sum())
        print(f"{case['name']} stop {i}: {ind1.shape[1]} records, {n_tie} which_wall ties, device differs from the golden on {n_diff} "
              f"of them; the reference's numba-on and numba-off runs differ on {ref_self} records")
        assert n_diff <= 2 * ref_self + 2, (n_diff, ref_self, n_tie)
        np.testing.assert_array_equal(ind2, gold[f"s{i}_ind2"], err_msg="ind2")
        cases.assert_close("frac", frac, gold[f"s{i}_frac"])
        cases.assert_close("tres", tres, gold[f"s{i}_tres"])


def test_budget_matches_oracle_on_own_log(oracle_lib):
    """Larger seeded run: the device post-processing against the numpy oracle fed with the very same log
    (so ties of which_wall have the same inputs on both sides): indices bit-exact, floats 1e-9."""
    import seaduck_b200 as sd
    from oracle import budget_oracle as bo
    from seaduck_b200 import lagrangian_budget as lb
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=45, nz=12)
    n = 5000
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=13, depth_range=(5, 700))
    od = sd.OceData(ds)
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, save_raw=True, uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    pt.note_taking(stamp=15)
    pt.to_next_stop(40 * cases.DAY)
    ind1, ind2, frac, tres, last, first = lb.find_ind_frac_tres(pt, od)
    rec = pt.raw_arrays()[-1]
    og = oracle_lib.OracleGrid(od)
    r1, r2, rfrac, rtres, rlast, rfirst = bo.find_ind_frac_tres(og, rec, rec["offset"])
    assert ind1.shape[1] > 50_000
    np.testing.assert_array_equal(first, rfirst)
    np.testing.assert_array_equal(last, rlast)
    np.testing.assert_array_equal(ind1, r1)
    np.testing.assert_array_equal(ind2, r2)
    cases.assert_close("frac", frac, rfrac)
    cases.assert_close("tres", tres, rtres)


BUDGET_VARIANTS = (("pad", dict(same_size=False), -1), ("same", dict(zname="szprime_same"), 1))


def test_calculate_budget_matches_reference():
    """calculate_budget (:603) on the device log against the reference's on its own log: wherever both chose the
    same walls (see the tie note above) the concentrations and every contribution agree to 1e-9."""
    import seaduck_b200 as sd
    from seaduck_b200 import lagrangian_budget as lb

    case = cases.case_llc_transport()
    gold = np.load(os.path.join(GOLD, "budget_llc_transport.npz"))
    od = sd.OceData(case["ds"])
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    shape = np.asarray(case["ds"]["maskC"]).shape
    for i, ts in enumerate(case["stops"]):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        ind1, ind2, _, _, _, _ = lb.find_ind_frac_tres(pt, od)
        same = (ind1 == gold[f"s{i}_ind1"]).all(axis=0) & (ind2 == gold[f"s{i}_ind2"]).all(axis=0)
        assert same.mean() > 0.98
        fields = cases.budget_fields(shape, seed=100 + i)
        for tag, kw, direction in BUDGET_VARIANTS:
            contr, conc, first, last = lb.calculate_budget(pt, fields, list(cases.RHS_TERMS), kw, "lhs", direction)
            np.testing.assert_array_equal(first, gold[f"s{i}_first"])
            np.testing.assert_array_equal(last, gold[f"s{i}_last"])
            want = gold[f"s{i}_bud_{tag}_conc"]
            np.testing.assert_array_equal(np.isnan(conc[same]), np.isnan(want[same]))
            cases.assert_close("conc", conc[same], want[same])
            pair = same[:-1] & same[1:]
            assert set(contr) == set(cases.RHS_TERMS) | {"error", "lhs"}
            for k, v in contr.items():
                w = gold[f"s{i}_bud_{tag}_{k}"]
                assert v.shape == w.shape
                np.testing.assert_array_equal(np.isnan(v[pair]), np.isnan(w[pair]), err_msg=k)
                if k == "error":  # a rounding residue of O(1e-16): only its size is meaningful
                    assert np.nanmax(np.abs(v[pair])) < 1e-13
                else:
                    cases.assert_close(k, v[pair], w[pair], rtol=1e-8)


def test_calculate_budget_matches_oracle_on_own_log():
    """Same inputs on both sides (the device's own log, indices, fractions and residence times): bit-exact."""
    import seaduck_b200 as sd
    from oracle import budget_oracle as bo
    from seaduck_b200 import lagrangian_budget as lb
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=30, nz=10)
    n = 1500
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=21, depth_range=(5, 300))
    od = sd.OceData(ds)
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, save_raw=True, uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    pt.note_taking(stamp=15)
    pt.to_next_stop(-30 * cases.DAY)
    ind1, ind2, frac, tres, last, first = lb.find_ind_frac_tres(pt, od)
    rec = pt.raw_arrays()[-1]
    fields = cases.budget_fields(np.asarray(ds["maskC"]).shape, seed=5)
    assert ind1.shape[1] > 10_000
    for tag, kw, direction in BUDGET_VARIANTS:
        contr, conc, bfirst, blast = lb.calculate_budget(pt, fields, list(cases.RHS_TERMS), kw, "lhs", direction)
        names = None if tag == "pad" else ("sxprime", "syprime", "szprime_same")
        want, wconc = bo.calculate_budget(rec, rec["offset"], ind1, ind2, frac, tres, fields, cases.RHS_TERMS, "lhs", direction, names)
        np.testing.assert_array_equal(bfirst, first)
        np.testing.assert_array_equal(blast, last)
        np.testing.assert_array_equal(conc, wconc)
        for k in want:
            np.testing.assert_array_equal(contr[k], want[k], err_msg=f"{tag} {k}")


def test_calculate_budget_argument_checks():
    import seaduck_b200 as sd
    from seaduck_b200 import lagrangian_budget as lb

    case = cases.case_llc_transport(n=20)
    od = sd.OceData(case["ds"])
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    pt.note_taking(stamp=15)
    pt.to_next_stop(5 * cases.DAY)
    fields = cases.budget_fields(np.asarray(case["ds"]["maskC"]).shape, seed=1)
    with pytest.raises(KeyError):
        lb.calculate_budget(pt, fields, ["nope"])
    with pytest.raises(ValueError):
        lb.calculate_budget(pt, fields, [])
    with pytest.raises(ValueError):
        lb.calculate_budget(pt, fields, ["adv"], dir_of_time=0)
    with pytest.raises(RuntimeError):  # sxprime / syprime / szprime differ in shape: same_size=True cannot stack them
        lb.calculate_budget(pt, fields, ["adv"])
    bad = dict(fields, adv=fields["adv"][:-1])
    with pytest.raises(ValueError):
        lb.calculate_budget(pt, bad, ["adv"], dict(same_size=False))


def test_wall_values_and_compat_check_match_oracle_and_reference(oracle_lib):
    """read_wall_list (:406) / check_particle_data_compat (:470) on the device: bit-exact against the numpy oracle fed
    with the device's own log (wall values incl. the 1e-16 leak of the unrounded rotation coefficients, wall
    velocities, both convergences), and 1e-9 against the reference's run wherever both logs sit in the same cell."""
    import seaduck_b200 as sd
    from oracle import budget_oracle as bo
    from seaduck_b200 import lagrangian_budget as lb

    case = cases.case_llc_transport()
    gold = np.load(os.path.join(GOLD, "budget_llc_transport.npz"))
    logs = np.load(os.path.join(GOLD, "lagrangian_llc_transport.npz"))
    od = sd.OceData(case["ds"])
    og = oracle_lib.OracleGrid(od)
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, save_raw=True, **case["kw"])
    shape = np.asarray(case["ds"]["maskC"]).shape
    for i, ts in enumerate(case["stops"]):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        rec = pt.raw_arrays()[-1]
        fields = cases.budget_fields(shape, seed=100 + i)
        fields.update(sx=fields["sxprime"], sy=fields["syprime"], sz=fields["szprime"],
                      divus=np.random.default_rng(200 + i).normal(size=shape))
        walls = [fields["sx"], fields["sy"], fields["sz"]]
        for scalar in (True, False):
            got = lb.read_wall_list(pt, od.tp, walls, scalar=scalar)
            np.testing.assert_array_equal(got, bo.read_wall_list(og, rec, walls, scalar=scalar))
        ok, (u_list, c_list, lag, eul) = lb.check_particle_data_compat(pt, fields, od.tp, debug=True)
        wu, wc, wlag, weul = bo.particle_data_compat(og, rec, fields)
        np.testing.assert_array_equal(u_list, wu)
        np.testing.assert_array_equal(c_list, wc)
        np.testing.assert_array_equal(lag, wlag)
        np.testing.assert_array_equal(eul, weul)
        assert ok == bool(gold[f"s{i}_compat_ok"])
        # against the reference's own run: same number of records, same cells (indices are bit-exact), values 1e-9
        assert c_list.shape == gold[f"s{i}_compat_c_list"].shape
        same_cell = np.ones(c_list.shape[1], bool)
        for k in ("fc", "iy", "ix", "izl"):
            same_cell &= rec[k] == logs[f"s{i}_log_{k}"]
        assert same_cell.all()
        np.testing.assert_array_equal(c_list, gold[f"s{i}_compat_c_list"])
        np.testing.assert_array_equal(eul, gold[f"s{i}_compat_eulerian_conv"])
        cases.assert_close("u_list", u_list, gold[f"s{i}_compat_u_list"])
        cases.assert_close("lagrangian_conv", lag, gold[f"s{i}_compat_lagrangian_conv"])
    # a field whose Eulerian convergence is made to fit passes the check, the random one above does not
    fields["divus"] = np.zeros(shape)
    fields["divus"][rec["izl"] - 1, rec["fc"], rec["iy"], rec["ix"]] = lag
    cells, first_at = np.unique(np.stack([rec["izl"], rec["fc"], rec["iy"], rec["ix"]]), axis=1, return_index=True)
    if cells.shape[1] == len(lag):  # only meaningful when no two records share a cell
        assert lb.check_particle_data_compat(pt, fields, od.tp)[0]
    with pytest.raises(ValueError):
        lb.read_wall_list(pt, od.tp, walls[:2])
    with pytest.raises(ValueError):
        lb.check_particle_data_compat(pt, fields, od.tp, use_tracer_name=3)
