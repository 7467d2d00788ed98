"""TEST INFRASTRUCTURE -- golden vectors for the stencils of the Eulerian budget (reference
``seaduck/eulerian_budget.py``: ``buffer_x_withface`` :214, ``buffer_y_withface`` :262, ``buffer_x_periodic``
:309, ``buffer_y_periodic`` :321, ``buffer_z_nearest`` :333, ``second_order_flux_limiter_x/y/z`` :363-445,
``third_order_upwind_z`` :448, ``third_order_DST_x/y`` :477-532), made by calling the UNMODIFIED reference
on seeded random fields with NaN land.

    python oracle/make_golden_eulerian_budget.py        # writes tests/golden/eulerian_budget.npz
"""

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refload import load_reference  # noqa: E402

sd = load_reference()
eb = importlib.import_module("seaduck.eulerian_budget")
tpm = importlib.import_module("seaduck.topology")

N, NZ = 8, 5


def inputs(seed=0):
    rng = np.random.default_rng(seed)
    s = rng.normal(size=(NZ, 13, N, N))
    s[rng.random(s.shape) < 0.08] = np.nan  # land
    u = rng.uniform(-0.9, 0.9, size=(NZ, 13, N, N))
    u[rng.random(u.shape) < 0.05] = 0.0
    sp = rng.normal(size=(NZ, N, N + 2))  # a single-face (periodic) field
    sp[rng.random(sp.shape) < 0.08] = np.nan
    return s, u, sp, rng


def main():
    s, u, sp, rng = inputs()
    tp = tpm.Topology.__new__(tpm.Topology)
    # a Topology of an n x n LLC grid without a dataset behind it
    tp.typ, tp.face_connect = "LLC", tpm.LLC_FACE_CONNECT
    tp.h_shape = tp.g_shape = (13, N, N)
    tp.num_face, tp.iymax, tp.ixmax, tp.itmax, tp.izmax = 13, N - 1, N - 1, 0, NZ - 1
    out = {"s": s, "u": u, "sp": sp}
    for f in range(13):
        xb = eb.buffer_x_withface(s, f, 2, 1, tp)
        yb = eb.buffer_y_withface(s, f, 2, 1, tp)
        out[f"xbuf_{f}"], out[f"ybuf_{f}"] = xb, yb
        out[f"dstx_{f}"] = eb.third_order_DST_x(xb, u[:, f])
        out[f"dsty_{f}"] = eb.third_order_DST_y(yb, u[:, f])
    out["xbuf_wide_5"] = eb.buffer_x_withface(s, 5, 3, 2, tp)
    out["ybuf_wide_10"] = eb.buffer_y_withface(s, 10, 1, 3, tp)
    out["xper"] = eb.buffer_x_periodic(sp, 2, 2)
    out["yper"] = eb.buffer_y_periodic(sp, 1, 3)
    out["znear"] = eb.buffer_z_nearest(sp, 2, 1)
    ux = rng.uniform(-0.9, 0.9, size=(NZ, N, N + 3))
    uy = rng.uniform(-0.9, 0.9, size=(NZ, N + 1, N + 2))
    uz = rng.uniform(-0.9, 0.9, size=(NZ, N, N + 2))
    ux[rng.random(ux.shape) < 0.05] = 0.0
    out["ux"], out["uy"], out["uz"] = ux, uy, uz
    out["fl_x"] = eb.second_order_flux_limiter_x(sp, ux)
    out["fl_y"] = eb.second_order_flux_limiter_y(sp, uy)
    out["fl_z"] = eb.second_order_flux_limiter_z(sp, uz)
    w = rng.uniform(-1, 1, size=(NZ, N, N + 2))
    out["w"] = w.copy()
    out["up3_z"] = eb.third_order_upwind_z(sp, w)
    out["w_after"] = w
    path = os.path.join(ROOT, "tests", "golden", "eulerian_budget.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {os.path.getsize(path) / 1e3:.0f} kB; nan in dstx_0: {int(np.isnan(out['dstx_0']).sum())}")


if __name__ == "__main__":
    main()
