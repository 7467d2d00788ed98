#!/usr/bin/env python
"""Experiment (EXP_MODES other than 0 need the library built from profiles/r02d_fed_queue_experiment.patch):
(1) the timeline of one sdk_track_host call (SDK_TRACK_TIMELINE: CUDA events between the stages of every
chunk) for the default chunking, with and without the device-side sort; (2) chunk-count sweep of sdk_interp_host on
the configs[2] workload.  Prints to stdout / stderr; nothing here is a bench value."""

import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def track():
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, synthetic

    ds = synthetic.llc(n=90, nz=50)
    od = sd.OceData(ds)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    n, t_stop = 1_000_000, 30 * 86400.0
    lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=1000)
    pt = sd.Particle(x=lon[:1000], y=lat[:1000], z=dep[:1000], t=np.zeros(1000), data=od, **kw)
    L = _lib.lib()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    hin = [pin(lon), pin(lat), pin(dep), pin(np.zeros(n))]
    rec = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    hstats = torch.zeros(4, dtype=torch.int64).pin_memory()
    scratch = torch.empty(int(L.sdk_track_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
    par = pt._params(200)
    par.stats = None

    def call(sort, reps=6):
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
        return 1e3 * float(np.min(ts[2:] or ts)), int(hstats[1])

    for mode in os.environ.get("EXP_MODES", "0").split(","):
        os.environ["SDK_TRACK_MODE"] = mode[0]
        os.environ.pop("SDK_TRACK_SCATTER", None)
        if mode.endswith("s"):
            os.environ["SDK_TRACK_SCATTER"] = "1"
        if len(mode) > 1 and mode[1:].isdigit():
            os.environ["SDK_TRACK_FEED_CTAS"] = mode[1:]  # "23" = mode 2 with 3 CTA slots for the fed launch
        for chunks in os.environ.get("EXP_CHUNKS", "3").split(","):
            os.environ["SDK_TRACK_CHUNKS"] = chunks
            for sort in ((0,) if mode[0] == "2" else (0, 1)):
                os.environ.pop("SDK_TRACK_TIMELINE", None)
                best, cross = call(sort)
                print(f"mode {mode} chunks {chunks} sort {sort}: best {best:.3f} ms -> {cross / best / 1e6:.2f}e9 crossings/s", flush=True)
                if os.environ.get("EXP_TIMELINE", "1") == "1":
                    os.environ["SDK_TRACK_TIMELINE"] = "1"
                    sys.stderr.write(f"--- timeline mode {mode} chunks {chunks} sort {sort}\n")
                    sys.stderr.flush()
                    call(sort, reps=-1)
    os.environ.pop("SDK_TRACK_TIMELINE", None)


def interp_sweep():
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import cases_interp as ci

    ds = synthetic.llc(n=270, nz=50, land_blobs=True, tracers=True)
    od = sd.OceData(ds)
    n = 10_000_000
    lon, lat, _ = synthetic.llc_seed_particles(ds, n, seed=3, lat_range=(-72.0, 88.0))
    dep = -np.random.default_rng(4).uniform(5.0, 3000.0, n)
    px, py, pz = (sd.pinned(n) for _ in range(3))
    px[:], py[:], pz[:] = lon, lat, dep
    uknw, vknw = ci.make_knw(sd.KnW, (ci.UKNW, ci.VKNW))
    var, kernels = ["T", "S", ("UVELMASS", "VVELMASS")], [sd.KnW(), sd.KnW(), (uknw, vknw)]
    for chunks in (4, 6, 8, 12, 16, 24, 32):
        os.environ["SDK_INTERP_CHUNKS"] = str(chunks)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = sd.OceInterp(od, list(var), px, py, pz, 0.0, kernel_list=list(kernels))
            ts.append(time.perf_counter() - t0)
            del res
        print(f"sdk_interp_host chunks {chunks:2d}: best {1e3 * min(ts[1:]):.3f} ms -> {n / min(ts[1:]) / 1e9:.3f}e9 points/s", flush=True)


def hostfields():
    """One 30-day stop of 10^5 particles on LLC90 x 50 with the transports resident in HBM and left in host memory."""
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=50)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    n = 100_000
    lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=1000)
    for limit in (None, 1):
        od = sd.OceData(ds, device_memory_limit=limit)
        best = 1e9
        for rep in range(3):
            pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, **kw)
            pt.lazy = True
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pt.to_next_stop(30 * 86400.0)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"fields {'in HBM' if limit is None else 'in pinned host memory'}: {best:.3f} ms per stop of {n} particles, "
              f"{pt.crossings / best / 1e6:.3f}e9 crossings/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["track", "interp"]
    if "hostfields" in what:
        hostfields()
    if "track" in what:
        track()
    if "interp" in what:
        interp_sweep()
