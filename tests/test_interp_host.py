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
