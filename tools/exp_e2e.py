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

    for zero_copy in (1, 0):
        if zero_copy:
            os.environ.pop("SDK_TRACK_NO_ZEROCOPY", None)
        else:
            os.environ["SDK_TRACK_NO_ZEROCOPY"] = "1"
        for chunks in (1, 2, 3, 4, 6, 8):
            os.environ["SDK_TRACK_CHUNKS"] = str(chunks)
            for sort in (0, 1):
                mean, best, cross = call(sort)
                print(f"records {'stored by the kernel (zero copy)' if zero_copy else 'packed + copied'}, chunks {chunks:2d} sort {sort}: "
                      f"{mean:.3f} ms mean, {best:.3f} ms best -> {cross / mean / 1e6:.2f}e9 crossings/s", flush=True)
    os.environ.pop("SDK_TRACK_NO_ZEROCOPY", None)
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
    # the stages of one unchunked pass through the C ABI, CUDA events in between
    import ctypes as C2

    st_in = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(lon=lon, lat=lat, dep=dep, t=np.zeros(n)).items()}
    for sort in (0, 1):
        loc = od.locate(x=lon, y=lat, z=dep)  # warm-up + allocation pattern
        del loc
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        names = ["locate", "cell sort", "gather", "get_u_du", "advect", "pack"]
        marks[0].record()
        p3 = sd.Particle.__new__(sd.Particle)
        # (stage timing through the Python API pieces that map 1:1 onto C-ABI calls)
        loc = od.locate(x=st_in["lon"].cpu().numpy(), y=st_in["lat"].cpu().numpy(), z=st_in["dep"].cpu().numpy())
        marks[1].record()
        del p3
        pt2 = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, sort=bool(sort), **kw)
        torch.cuda.synchronize()
        # sort + gather + get_u_du separately on the located state of pt2
        stt = pt2._st
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        perm = torch.empty(n, dtype=torch.int32, device="cuda")
        nbytes = int(L.sdk_cell_sort_scratch_bytes(n))
        scr = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        _lib.check(L.sdk_cell_sort(od.grid_handle(), n, stt["face"].data_ptr(), stt["iy"].data_ptr(), stt["ix"].data_ptr(), stt["izl"].data_ptr(),
                                   perm.data_ptr(), scr.data_ptr(), nbytes, _lib.current_stream()), "sort")
        e[1].record()
        new = {k: (None if v is None else torch.empty_like(v)) for k, v in stt.items()}
        src, dst = pt2._cparticles(state=stt), pt2._cparticles(state=new)
        _lib.check(L.sdk_particles_gather(C2.byref(src), C2.byref(dst), perm.data_ptr(), _lib.current_stream()), "gather")
        e[2].record()
        pt2.get_u_du()
        e[3].record()
        pt2.lazy = True
        pt2.to_next_stop(t_stop)
        e[4].record()
        torch.cuda.synchronize()
        print(f"stages on {n} particles (state {'cell-sorted' if sort else 'in seeding order'}): locate incl. pageable H2D {marks[0].elapsed_time(marks[1]):.3f} ms, "
              f"cell sort {e[0].elapsed_time(e[1]):.3f}, gather {e[1].elapsed_time(e[2]):.3f}, get_u_du {e[2].elapsed_time(e[3]):.3f}, "
              f"advect + finish {e[3].elapsed_time(e[4]):.3f}", flush=True)
    # locate alone on device-resident inputs
    import ctypes as C3
    for rep in range(2):
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outl = {}
        locs = _lib.Located()
        for nme, dt, mult in (("face", torch.int32, 1), ("iy", torch.int32, 1), ("ix", torch.int32, 1), ("rx", torch.float64, 1), ("ry", torch.float64, 1),
                              ("px", torch.float64, 4), ("py", torch.float64, 4), ("dx", torch.float32, 1), ("dy", torch.float32, 1),
                              ("izl", torch.int32, 1), ("rzl", torch.float64, 1), ("dzl", torch.float64, 1), ("bzl", torch.float64, 1), ("status", torch.int32, 1)):
            outl[nme] = torch.empty(n * mult, dtype=dt, device="cuda")
            setattr(locs, nme, outl[nme].data_ptr())
        a0.record()
        _lib.check(L.sdk_locate(od.grid_handle(), n, st_in["lon"].data_ptr(), st_in["lat"].data_ptr(), st_in["dep"].data_ptr(), None, None, 0,
                                C3.byref(locs), _lib.current_stream()), "locate")
        a1.record()
        torch.cuda.synchronize()
        print(f"sdk_locate alone (device inputs, seeding order): {a0.elapsed_time(a1):.3f} ms", flush=True)

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
