#!/usr/bin/env python
"""Experiment: how much of an advection launch is its tail, and what a work-aware particle order is worth.

The persistent kernel hands particles to lanes in queue order.  Particles need between 1 and > 50 analytic steps;
when the queue runs dry the launch waits for whoever still has many steps to go.  This script times one launch of the
bench workload (LLC90 x 50, 10^6 particles, 30 d) for several queue orders: cell order (what the product keeps),
longest-first by the TRUE step count of a dry run (an upper bound on what any estimate can give), and
longest-first by an estimate available before the launch (|u| + |v| + |w| in cells per second times the leg).
One line per order on stdout.
"""

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--days", type=float, default=30.0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=50)
    od = sd.OceData(ds)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    n, t_stop = a.particles, a.days * 86400.0

    def launch_ms(lon, lat, dep, sort):
        pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, sort=sort, **kw)
        pt.lazy = True
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pt.to_next_stop(t_stop)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), pt

    results = {}
    for rep in range(a.reps):
        lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=500 + rep)
        ms, pt = launch_ms(lon, lat, dep, True)
        inv = pt._inv
        niter = pt._st["niter"][inv].cpu().numpy()
        rank_in_cell_order = inv.cpu().numpy()  # position of particle i in cell order
        # estimate known before the launch: cells per second along the three axes times the leg
        pt0 = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, sort=False, **kw)
        est = t_stop * (np.abs(pt0.u) + np.abs(pt0.v) + np.abs(pt0.w))
        orders = {
            "cell order (product)": np.argsort(rank_in_cell_order, kind="stable"),
            "true steps, longest first (then cell)": np.lexsort((rank_in_cell_order, -niter)),
            "true steps in 8 log2 buckets, longest first (then cell)": np.lexsort((rank_in_cell_order, -np.floor(np.log2(np.maximum(niter, 1))))),
            "estimate in log2 buckets, longest first (then cell)": np.lexsort((rank_in_cell_order, -np.floor(np.log2(np.maximum(est, 0.5))))),
            "estimate, longest first (then cell)": np.lexsort((rank_in_cell_order, -est)),
            "seeding order": np.arange(n),
        }
        if rep == 0:
            q = np.quantile(niter, [0.5, 0.9, 0.99, 0.999, 1.0])
            print(f"steps per particle: mean {niter.mean():.2f}, quantiles 50/90/99/99.9/100 % = {q}", flush=True)
            print(f"estimate vs true steps: correlation {np.corrcoef(est, niter)[0, 1]:.3f}", flush=True)
        for name, o in orders.items():
            ms, _ = launch_ms(lon[o], lat[o], dep[o], False)
            results.setdefault(name, []).append(ms)
    for name, v in results.items():
        print(f"{name:60s} {np.mean(v):.3f} ms  (launch + finish_stop, {a.reps} seeds: {[round(x, 3) for x in v]})")


if __name__ == "__main__":
    main()
