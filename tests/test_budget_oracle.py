"""CPU: the numpy oracle of the Lagrangian-budget post-processing (oracle/budget_oracle.py) on the crossing
logs of the Lagrangian goldens against the budget goldens -- both produced by the unmodified reference."""

import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("make", [cases.case_llc_transport, cases.case_overturning], ids=lambda f: f.__name__)
def test_budget_oracle_matches_reference(make, oracle_lib):
    import seaduck_b200 as sd
    from oracle import budget_oracle as bo

    case = make()
    logs = np.load(os.path.join(GOLD, f"lagrangian_{case['name']}.npz"))
    gold = np.load(os.path.join(GOLD, f"budget_{case['name']}.npz"))
    og = oracle_lib.OracleGrid(sd.OceData(case["ds"]))
    for i in range(len(case["stops"])):
        rec = {k: logs[f"s{i}_log_{k}"] for k in cases.LOG_INT + cases.LOG_FLT if f"s{i}_log_{k}" in logs}
        if "fc" not in rec:
            rec["fc"] = np.zeros_like(rec["iy"])
        ind1, ind2, frac, tres, last, first = bo.find_ind_frac_tres(og, rec, logs[f"s{i}_log_offset"])
        np.testing.assert_array_equal(first, gold[f"s{i}_first"])
        np.testing.assert_array_equal(last, gold[f"s{i}_last"])
        np.testing.assert_array_equal(ind1, gold[f"s{i}_ind1"])
        np.testing.assert_array_equal(ind2, gold[f"s{i}_ind2"])
        np.testing.assert_allclose(frac, gold[f"s{i}_frac"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(tres, gold[f"s{i}_tres"], rtol=1e-12, atol=0)


def test_calculate_budget_oracle_matches_reference():
    """oracle/budget_oracle.calculate_budget against the reference's calculate_budget (:603) on the reference's
    own ind / frac / tres: every term, the error and the concentrations, to the bit (NaN where it has NaN)."""
    from oracle import budget_oracle as bo

    case = cases.case_llc_transport()
    logs = np.load(os.path.join(GOLD, "lagrangian_llc_transport.npz"))
    gold = np.load(os.path.join(GOLD, "budget_llc_transport.npz"))
    shape = np.asarray(case["ds"]["maskC"]).shape
    for i in range(len(case["stops"])):
        rec = {k: logs[f"s{i}_log_{k}"] for k in ("tt", "izl", "fc", "iy", "ix")}
        fields = cases.budget_fields(shape, seed=100 + i)
        for tag, names, direction in (("pad", None, -1), ("same", ("sxprime", "syprime", "szprime_same"), 1)):
            out, conc = bo.calculate_budget(
                rec, logs[f"s{i}_log_offset"], gold[f"s{i}_ind1"], gold[f"s{i}_ind2"], gold[f"s{i}_frac"], gold[f"s{i}_tres"],
                fields, cases.RHS_TERMS, "lhs", direction, names,
            )
            np.testing.assert_array_equal(conc, gold[f"s{i}_bud_{tag}_conc"])
            assert set(out) == set(cases.RHS_TERMS) | {"error", "lhs"}
            for k, v in out.items():
                np.testing.assert_array_equal(v, gold[f"s{i}_bud_{tag}_{k}"], err_msg=f"{tag} {k}")
            assert np.isnan(out["adv"]).any()  # the 0 / 0 cells are in the sample


def test_wall_values_and_compat_oracle_match_reference(oracle_lib):
    """read_wall_list (:406) and check_particle_data_compat (:470) of the reference on its own log against the
    oracle's restatement: everything to the bit."""
    import seaduck_b200 as sd
    from oracle import budget_oracle as bo

    case = cases.case_llc_transport()
    logs = np.load(os.path.join(GOLD, "lagrangian_llc_transport.npz"))
    gold = np.load(os.path.join(GOLD, "budget_llc_transport.npz"))
    og = oracle_lib.OracleGrid(sd.OceData(case["ds"]))
    shape = np.asarray(case["ds"]["maskC"]).shape
    for i in range(len(case["stops"])):
        rec = {k: logs[f"s{i}_log_{k}"] for k in cases.LOG_INT + cases.LOG_FLT if f"s{i}_log_{k}" in logs}
        fields = cases.budget_fields(shape, seed=100 + i)
        fields.update(sx=fields["sxprime"], sy=fields["syprime"], sz=fields["szprime"],
                      divus=np.random.default_rng(200 + i).normal(size=shape))
        walls = [fields["sx"], fields["sy"], fields["sz"]]
        np.testing.assert_array_equal(bo.read_wall_list(og, rec, walls, scalar=True), gold[f"s{i}_walls_scalar"])
        np.testing.assert_array_equal(bo.read_wall_list(og, rec, walls, scalar=False), gold[f"s{i}_walls_vector"])
        u_list, c_list, lag, eul = bo.particle_data_compat(og, rec, fields)
        np.testing.assert_array_equal(u_list, gold[f"s{i}_compat_u_list"])
        np.testing.assert_array_equal(c_list, gold[f"s{i}_compat_c_list"])
        np.testing.assert_array_equal(eul, gold[f"s{i}_compat_eulerian_conv"])
        np.testing.assert_array_equal(lag, gold[f"s{i}_compat_lagrangian_conv"])
