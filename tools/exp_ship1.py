#!/usr/bin/env python
"""Experiment: cost of the fused output exchange inside the advection kernel, on ONE GPU (the rank is its own peer):
kernel time of sdk_advect on the bench workload without and with sdk_advect_params.ship."""
import ctypes as C
import os
import sys

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
    slots = 2

    class OneRank:
        def __init__(self):
            self.buf = torch.zeros(4096 + slots * n * 32, dtype=torch.uint8, device="cuda")
            self.ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
            self.x = _lib.Exchange()
            self.x.world, self.x.rank, self.x.n_slots, self.x.n_pad = 1, 0, slots, n
            self.x.signal[0], self.x.recv[0], self.x.ticket = self.buf.data_ptr(), self.buf.data_ptr() + 4096, self.ticket.data_ptr()
            self.step = 0

        def next_step(self):
            self.step += 1
            return self.step

    ex = OneRank()
    L = _lib.lib()
    res = {(s, sh): [] for s in (True, False) for sh in (False, True)}
    for rep in range(6):
        lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=700 + rep)
        for sort, ship in res:
            pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, sort=sort, **kw)
            pt.lazy = True
            if ship:
                pt.ship_to(ex)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pt.to_next_stop(t_stop)
            e1.record()
            if ship:
                _lib.check(L.sdk_exchange_wait(C.byref(ex.x), pt.last_ship_step, 1, 1, _lib.current_stream()), "wait")
            torch.cuda.synchronize()
            if rep:
                res[(sort, ship)].append(e0.elapsed_time(e1))
    for (sort, ship), v in res.items():
        print(f"advect + finish_stop, {'cell-sorted' if sort else 'seeding order'}, fused exchange {'on ' if ship else 'off'}: {np.mean(v):.4f} ms  {[round(x, 3) for x in v]}")


if __name__ == "__main__":
    main()
