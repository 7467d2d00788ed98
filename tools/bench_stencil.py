#!/usr/bin/env python
"""Bandwidth of the Eulerian-budget stencil kernels (csrc/stencil.cu) at LLC270 size: halo gather of one
face for all levels, and the third-order DST wall values from it.  Prints one JSON line per kernel."""

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seaduck_b200 import eulerian_budget as eb  # noqa: E402
from seaduck_b200.topology import Topology  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n, nz, nt = 270, 50, 8  # eight time records at once: 8 x 50 x 13 x 270 x 270 x 8 B = 3.0 GB
    dev = torch.device("cuda", 0)
    tp = Topology.from_shape((13, n, n), typ="LLC")
    s = torch.randn((nt, nz, 13, n, n), dtype=torch.float64, device=dev)
    u = torch.rand((nt, nz, n, n), dtype=torch.float64, device=dev) - 0.5
    peak = 6650.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    face = 4
    xb = eb.buffer_x_withface(s, face, 2, 1, tp)
    ms = timed(lambda: eb.buffer_x_withface(s, face, 2, 1, tp))
    by = xb.numel() * 16  # read one value, write one value
    print(json.dumps({"kernel": "gather_plane_kernel (buffer_x_withface, one face)", "ms": ms, "GB/s": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                      "includes": "torch.empty of the output + ctypes call"}))
    ms = timed(lambda: eb.third_order_DST_x(xb, u))
    by = (xb.numel() + 2 * u.numel()) * 8
    print(json.dumps({"kernel": "wall_stencil_kernel (third_order_DST_x)", "ms": ms, "GB/s": by / ms / 1e6, "frac": by / ms / 1e6 / peak}))
    s3 = s[:, :, face].contiguous()
    ux = torch.rand((nt, nz, n, n + 1), dtype=torch.float64, device=dev) - 0.5
    ms = timed(lambda: eb.second_order_flux_limiter_x(s3, ux))
    by = (s3.numel() * 3 + 2 * ux.numel()) * 8  # gather read+write, stencil read, vel, out
    print(json.dumps({"kernel": "gather + wall_stencil_kernel (second_order_flux_limiter_x)", "ms": ms, "GB/s": by / ms / 1e6, "frac": by / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
