"""Where a stop of BASELINE.json configs[3] goes (LLC4320 surface, 1.25e7 particles, 3-hour stops): the whole
``to_list_of_time(snapshots="records")`` loop against its pieces, CUDA events on the launching stream, one GPU.

    python tools/exp_c4_stages.py [--grid 4320] [--particles 12500000] [--stops 10]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4320)
    ap.add_argument("--particles", type=int, default=12_500_000)
    ap.add_argument("--stops", type=int, default=10)
    a = ap.parse_args()
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, synthetic

    DAY = 86400.0
    dev = torch.device("cuda:0")
    ds, fields = synthetic.llc_surface_device(n=a.grid, nt=3, device=dev)
    od = sd.OceData(ds)
    for name, (dims, tensor) in fields.items():
        od.stage_device_field(name, tensor, dims)
    del fields
    n = a.particles
    lon, lat, _ = synthetic.llc_seed_particles(None, n, seed=4000)
    t_first, dt = 4.45 * DAY, 3 * 3600.0
    pt = sd.Particle(x=lon, y=lat, z=None, t=np.full(n, t_first), data=od, uname="utrans", vname="vtrans", wname=None, transport=True)
    pt.lazy = True
    K = a.stops
    pt.reserve_records(4 * K + 8)
    L = _lib.lib()

    def timed(fn, reps=3):
        out = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            e0.record()
            fn()
            e1.record()
            w_issue = time.perf_counter() - w0
            torch.cuda.synchronize()
            out.append((e0.elapsed_time(e1), w_issue * 1e3))
        return out

    first = [0]

    def loop():
        stops = t_first + dt * (first[0] + 1 + np.arange(K))
        first[0] += K
        pt.to_list_of_time(stops, snapshots="records")

    res = timed(loop, reps=3)
    for ms, issue in res:
        print(f"to_list_of_time, {K} stops: {ms / K:.3f} ms per stop on the device, host issued the call in {issue / K:.3f} ms per stop")

    # pieces, on the state as it is now (restored before every repetition)
    keep = {k: (None if v is None else v.clone()) for k, v in pt._st.items()}
    t_now = float(pt._st["t"][0].item())
    stats = torch.zeros(4, dtype=torch.int64, device=dev)

    def restore():
        for k, v in keep.items():
            if v is not None:
                pt._st[k].copy_(v)

    sig = torch.zeros(_lib.SIGNAL_WORDS + 1, dtype=torch.int64, device=dev)
    alive = []

    def local_exchange(block):
        """sdk_exchange with this GPU as its only rank and `block` as the receive buffer: sdk_advect_params.ship then makes
        the advection kernel store every particle's record there when it writes the particle back."""
        x = _lib.Exchange()
        x.world, x.rank, x.n_slots, x.n_pad = 1, 0, 1, n
        x.recv[0] = block.data_ptr()
        x.recv_multicast = None
        x.signal[0] = sig.data_ptr()
        x.ticket = sig.data_ptr() + 8 * _lib.SIGNAL_WORDS
        alive.append(x)
        return x

    def advect(target):
        def run():
            p = pt._cparticles()
            par = pt._params(pt.max_iteration)
            stats.zero_()
            par.stats = stats.data_ptr()
            if target is not None:
                x = local_exchange(target)
                par.ship = C.addressof(x)
                par.ship_index = None if pt._perm is None else pt._perm.data_ptr()
                par.ship_step = 1
            _lib.check(L.sdk_advect(od.grid_handle(), C.byref(pt._vel), C.byref(p), t_now + dt, C.byref(par), None, 1, _lib.current_stream()), "sdk_advect")
        return run

    block = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    for label, fn in (("sdk_advect", advect(None)), ("sdk_advect + record store", advect(block))):
        vals = []
        for _ in range(3):
            restore()
            vals.append(timed(fn, reps=1)[0][0])
        print(f"{label}: {np.mean(vals):.3f} ms ({[round(v, 3) for v in vals]}), {int(stats[1].item())} crossings")
    midp = od.to_device("time_midp", od.time_midp, np.float64)

    def finish():
        _lib.check(L.sdk_finish_stop(pt._st["t"].data_ptr(), pt._st["it"].data_ptr(), midp.data_ptr(), len(od.time_midp), t_now + dt,
                                     pt._stats.data_ptr(), pt._total.data_ptr(), pt.N, _lib.current_stream()), "sdk_finish_stop")

    print("sdk_finish_stop:", [round(v[0], 3) for v in timed(finish)])
    print("record_snapshot (sdk_pack_records):", [round(v[0], 3) for v in timed(lambda: pt.record_snapshot(t_now, out=block))])
    print("get_u_du:", [round(v[0], 3) for v in timed(pt.get_u_du)])
    print("update_uvw_array (device / host ms):", [(round(v[0], 3), round(v[1], 3)) for v in timed(pt.update_uvw_array)])
    print("_stats.zero_:", [round(v[0], 4) for v in timed(pt._stats.zero_)])


if __name__ == "__main__":
    main()
