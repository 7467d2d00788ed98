#!/usr/bin/env python
"""The reference's LLC4320 notebook at full size on ONE GPU (docs/sciserver_notebooks/LLC4320.md:91-131): a surface-only
LLC4320-shaped field with time snapshots, 1.25e7 particles, 239 three-hour stops through
``Particle.to_list_of_time(snapshots="records")`` -- 239 x 400 MB of output records kept in HBM.
Prints one JSON line: wall time, crossings/s, device memory high-water mark, and a spot check of two snapshots."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    n = int(os.environ.get("N_PARTICLES", 12_500_000))
    grid = int(os.environ.get("GRID", 4320))
    stops_n = int(os.environ.get("N_STOPS", 239))
    dev = torch.device("cuda", 0)
    t0 = time.perf_counter()
    ds, fields = synthetic.llc_surface_device(n=grid, nt=4, device=dev)
    od = sd.OceData(ds)
    for name, (dims, tensor) in fields.items():
        od.stage_device_field(name, tensor, dims)
    del fields
    lon, lat, _ = synthetic.llc_seed_particles(None, n, seed=4)
    pt = sd.Particle(x=lon, y=lat, z=None, t=np.zeros(n), data=od, uname="utrans", vname="vtrans", wname=None, transport=True)
    pt.lazy = True
    setup = time.perf_counter() - t0
    stops = 3 * 3600.0 * (1 + np.arange(stops_n))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    got_stops, snaps = pt.to_list_of_time(stops, snapshots="records")
    crossings = pt.crossings  # (waits for the device)
    run = time.perf_counter() - t1
    first, last = snaps[0], snaps[-1]
    moved = np.abs(((last.lon - first.lon + 180) % 360) - 180)
    line = {
        "what": f"LLC{grid} surface, 4 time snapshots, {n} particles, {stops_n} three-hour stops, to_list_of_time(snapshots='records')",
        "setup_s": setup, "run_s": run, "stops_returned": len(snaps), "crossings": int(crossings), "crossings_per_s": crossings / run,
        "ms_per_stop": 1e3 * run / len(snaps), "record_bytes_in_hbm": int(sum(s._records.numel() for s in snaps)),
        "max_memory_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
        "spot_check": {"particles": int(last.N), "median_lon_displacement_deg": float(np.nanmedian(moved)), "t_first": float(first.t[0]),
                       "t_last": float(last.t[0]), "cells_in_range": bool((last.ix >= 0).all() and (last.ix < grid).all() and (last.face < 13).all())},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
