#!/usr/bin/env python
"""Experiment: where the time of sdk_track_host goes (bench workload: LLC90 x 50, 10^6 particles, 30 d).

Times the call for several chunk counts (SDK_TRACK_CHUNKS) with and without the device-side sort, and the pieces on
their own: the two copies, and the kernels of one unchunked pass by CUDA events.  One line per setting.
"""

import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, synthetic

    ds = synthetic.llc(n=90, nz=50)
    od = sd.OceData(ds)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    n, t_stop = 1_000_000, 30 * 86400.0
    lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=1000)
    pt = sd.Particle(x=lon[:1000], y=lat[:1000], z=dep[:1000], t=np.zeros(1000), data=od, **kw)  # velocity descriptor
    L = _lib.lib()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    hin = [pin(lon), pin(lat), pin(dep), pin(np.zeros(n))]
    rec = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    hstats = torch.zeros(4, dtype=torch.int64).pin_memory()
    scratch = torch.empty(int(L.sdk_track_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
    par = pt._params(200)
    par.stats = None

    def call(sort, reps=8):
        io = _lib.TrackIO()
        io.n, io.sort = n, sort
        io.lon, io.lat, io.dep, io.t = (h.data_ptr() for h in hin)
        io.out_records, io.out_stats = rec.data_ptr(), hstats.data_ptr()
        ts = []
        for i in range(reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(L.sdk_track_host(od.grid_handle(), C.byref(pt._vel), C.byref(io), None, 0, t_stop, C.byref(par), scratch.data_ptr(),
                                        _lib.current_stream()), "sdk_track_host")
            ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.mean(ts[2:])), 1e3 * float(np.min(ts[2:])), int(hstats[1])

    for chunks in (1, 2, 3, 4, 6, 8, 12, 16, 24):
        os.environ["SDK_TRACK_CHUNKS"] = str(chunks)
        for sort in (0, 1):
            mean, best, cross = call(sort)
            print(f"chunks {chunks:2d} sort {sort}: {mean:.3f} ms mean, {best:.3f} ms best -> {cross / mean / 1e6:.2f}e9 crossings/s", flush=True)
    # the copies alone
    dev_in = torch.empty(4 * n, dtype=torch.float64, device="cuda")
    for name, fn in (
        ("H2D 32 MB (4 arrays)", lambda: [dev_in[k * n : (k + 1) * n].copy_(hin[k], non_blocking=True) for k in range(4)]),
        ("D2H 32 MB (records)", lambda: rec.copy_(scratch[: 32 * n].view(n, 32), non_blocking=True)),
    ):
        ts = []
        for i in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(f"{name}: {1e3 * min(ts[1:]):.3f} ms", flush=True)
    # kernels of the resident path by CUDA events
    def ev():
        return torch.cuda.Event(enable_timing=True)

    for sort in (False, True):
        marks = [ev() for _ in range(4)]
        torch.cuda.synchronize()
        marks[0].record()
        p2 = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, sort=sort, **kw)
        marks[1].record()
        p2.lazy = True
        p2.to_next_stop(t_stop)
        marks[2].record()
        r = p2.record_snapshot(t_stop)
        marks[3].record()
        torch.cuda.synchronize()
        print(f"Particle(sort={sort}): construct (H2D pageable + locate + sort + get_u_du) {marks[0].elapsed_time(marks[1]):.3f} ms, "
              f"advect {marks[1].elapsed_time(marks[2]):.3f} ms, pack {marks[2].elapsed_time(marks[3]):.3f} ms", flush=True)
        del r


if __name__ == "__main__":
    main()
