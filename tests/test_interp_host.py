"""CPU: host logic of the interpolation path -- the rim tables (neighbour indices near face edges) against
the oracle's topology walk on every rim cell, and the polynomial form of the Lagrange weights that the
device evaluates against the stencil weights of the host KnW mirror (itself checked against the
reference in test_kernel_weight.py)."""

import numpy as np
import pytest

import cases_interp as ci
from seaduck_b200 import interp, synthetic
from seaduck_b200.kernel_weight import KnW
from seaduck_b200.topology import Topology


def _unpack(packed):
    return packed >> 26 & 0x1F, packed >> 13 & 0x1FFF, packed & 0x1FFF


@pytest.mark.parametrize("stencil,cuvwg", [("cross9", "C"), ("cross9", "U"), ("cross9", "V"), ("square9", "C"), ("uv", "U"), ("uv", "V"), ("cross5", "G")])
def test_rim_table_matches_oracle_walk(stencil, cuvwg, oracle_lib):
    from oracle import interp_oracle as io

    import seaduck_b200 as sd

    n = 12
    ds = synthetic.llc(n=n, nz=0)
    od = sd.OceData(ds)
    og = oracle_lib.OracleGrid(od)
    kernel = ci.STENCILS[stencil]
    reach, dense, rim, code = interp.build_rim(od.tp, kernel, cuvwg)
    iy1, ix1 = interp.rim_cells(n, n, reach, dense)
    f = np.repeat(np.arange(13), len(iy1))
    iy, ix = np.tile(iy1, 13), np.tile(ix1, 13)
    ff, fy, fx, err = io.fatten_h(og, f, iy, ix, kernel, cuvwg)
    ok = err == 0
    assert ((rim == interp.NODE_RAISES) == ~ok).all(), "cells where the walk fails must be flagged, and only those"
    pf, py, px = _unpack(rim.astype(np.int64))
    np.testing.assert_array_equal(pf[ok], ff[ok])
    np.testing.assert_array_equal(py[ok], fy[ok])
    np.testing.assert_array_equal(px[ok], fx[ok])
    if cuvwg in ("U", "V"):
        # component swap / sign codes against the oracle's four_matrix_for_uv
        safe_f = np.where(ok, ff, f[:, None])
        (ufu, ufv, vfu, vfv), err2 = io.four_matrix_for_uv(og, safe_f)
        own, other = (ufu, ufv) if cuvwg == "U" else (vfv, vfu)
        good = ok & (err2 == 0)
        np.testing.assert_array_equal((code & 1).astype(bool)[good], np.abs(other).astype(bool)[good])
        np.testing.assert_array_equal((code >> 1 & 1).astype(bool)[good], ((own + other) < 0)[good])


def test_rim_table_on_closed_box():
    """Stepping out of a closed box indexes [-1, -1] in the reference: NODE_LAST."""
    tp = Topology.from_shape((6, 7), typ="box")
    reach, dense, rim, code = interp.build_rim(tp, ci.STENCILS["cross5"], "C")
    assert dense == 1 and code is None
    corner = rim[0]  # cell (0, 0): nodes [0,0], [0,1], [0,-1], [-1,0], [1,0]
    assert corner[0] == interp._pack(0, 0, 0)
    assert corner[2] == interp.NODE_LAST and corner[3] == interp.NODE_LAST
    assert corner[1] == interp._pack(0, 1, 0) and corner[4] == interp._pack(0, 0, 1)


SPECS = [
    {}, dict(kernel="cross5"), dict(kernel="square9"),
    dict(hkernel="dx", h_order=1, inheritance=[list(range(9)), [0, 1, 3, 5, 7]]),
    dict(hkernel="dy", h_order=2, inheritance=[list(range(9))]),
    dict(kernel="square9", hkernel="dx", h_order=1, inheritance=[list(range(9))]),
    dict(kernel="xline5", hkernel="dx", h_order=2, inheritance=[[0, 1, 2, 3, 4], [0, 1, 2]]),
    ci.UKNW, ci.DVKNW, ci.WKNW,
]


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: "-".join(f"{k}={v}" for k, v in s.items() if k != "inheritance") or "default")
def test_weight_polynomials_equal_stencil_weights(spec):
    knw = ci.make_knw(KnW, spec)
    rng = np.random.default_rng(0)
    rx, ry = rng.uniform(-0.5, 0.5, 200), rng.uniform(-0.5, 0.5, 200)
    for doll, st in zip(knw.inheritance, knw.hfuncs):
        ref = st(rx, ry)  # (N, nodes of the level)
        polys = interp.level_polynomials(st, [knw.kernel[i] for i in doll])
        for j, (mode, a, c) in enumerate(polys):
            pa = np.polynomial.polynomial.polyval(rx, a)
            pc = np.polynomial.polynomial.polyval(ry, c)
            w = {interp.W_A: pa, interp.W_C: pc, interp.W_A_PLUS_C: pa + pc, interp.W_A_TIMES_C: pa * pc}[mode]
            np.testing.assert_allclose(w, ref[:, j], rtol=1e-12, atol=1e-13)


def test_launch_plan_groups_scalars_and_keeps_the_nesting():
    """``_plan`` (what has to be launched for a request) and ``_assemble`` (results back in the request's nesting):
    scalars that share kernel, grid position and dtype go into one multi-field launch of at most four, a repeated
    (variable, kernel) pair is computed once, vectors stay whole on grids with faces and split without."""
    import seaduck_b200 as sd

    ds = synthetic.llc(n=8, nz=3, tracers=True)
    od = sd.OceData(ds)
    k1, k2 = KnW(), KnW(vkernel="linear")
    uk, vk = ci.make_knw(KnW, (ci.UKNW, ci.VKNW))
    var = ["T", "S", ("UVELMASS", "VVELMASS"), "T", "S"]
    knw = [k1, k1, (uk, vk), k2, k1]
    single, v, k, pre, prefix = interp._broadcast_inputs(None, list(var), list(knw), None, None)
    assert not single
    plan = interp._plan(od, True, v, k, pre, prefix)
    kinds = [(kind, names) for _, kind, names, *_ in plan]
    assert kinds == [("m", ["T", "S"]), ("v", ("UVELMASS", "VVELMASS")), ("s", "T")]
    final = {}
    for keys, kind, names, kernels, _, _ in plan:
        for key in keys:
            final[key] = ("result", key[0] if kind != "v" else names)
    out = interp._assemble(final, v, k, True, single)
    assert [o[1] for o in out] == ["T", "S", ("UVELMASS", "VVELMASS"), "T", "S"]
    assert out[0] is out[0] and out[1] is out[4]  # the repeated (S, k1) request is answered by the same launch
    # without faces a vector is two scalar jobs on staggered grids
    plan2 = interp._plan(od, False, [("UVELMASS", "VVELMASS")], [(k1, k1)], [None], [None])
    assert [kind for _, kind, *_ in plan2] == ["s", "s"]
    tup = interp._assemble({("UVELMASS", k1): 1, ("VVELMASS", k1): 2}, [("UVELMASS", "VVELMASS")], [(k1, k1)], False, True)
    assert tup == (1, 2)
    # five scalars with one kernel: one launch of four and one single
    plan3 = interp._plan(od, True, ["T", "S", "maskC", "T", "S"], [k1] * 5, [None] * 5, [None] * 5)
    assert [kind for _, kind, *_ in plan3][:1] == ["m"] and sum(len(keys) for keys, *_ in plan3) == 3


def test_index_family_selectors_follow_the_kernel():
    """Which of sdk_locate's index families a variable reads in the host pipeline: Z or Zl by the variable's dims, nearest
    or linear by the kernel's vkernel / tkernel -- and the reference's complaints when the points lack the coordinate."""
    import seaduck_b200 as sd

    od = sd.OceData(synthetic.llc(n=8, nz=3, nt=3, tracers=True))
    sel = interp._vertical_time_selectors
    tdims = od["T"].dims
    wdims = od["WVELMASS"].dims
    assert sel(od, tdims, KnW(), True, True) == (interp.SEL_Z, interp.SEL_T)
    assert sel(od, tdims, KnW(vkernel="linear", tkernel="dt"), True, True) == (interp.SEL_Z_LIN, interp.SEL_T_LIN)
    assert sel(od, wdims, KnW(vkernel="dz"), True, True) == (interp.SEL_ZL_LIN, interp.SEL_T)
    assert sel(od, wdims, KnW(), True, True)[0] == interp.SEL_ZL
    assert sel(od, ("face", "Y", "X"), KnW(), False, False) == (interp.SEL_NONE, interp.SEL_NONE)
    with pytest.raises(ValueError, match="no depth"):
        sel(od, tdims, KnW(), False, True)
    with pytest.raises(ValueError, match="no time"):
        sel(od, tdims, KnW(), True, False)
    with pytest.raises(ValueError, match="vkernel"):
        sel(od, tdims, type("K", (), {"vkernel": "cubic", "tkernel": "nearest"})(), True, True)
