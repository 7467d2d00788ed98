#!/usr/bin/env python
"""Multi-GPU check of the record exchange and of ShardedParticle (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/check_exchange.py

1. sdk_pack_ship / sdk_exchange_wait: every rank ships known records for 40 steps through 2 slots, with plain peer
   stores and (where the NVSwitch can replicate) with multimem.st; every rank checks every block of every step.
2. bandwidth of the exchange alone (1e6 and 1.25e7 records per rank).
3. ShardedParticle.to_list_of_time with gather="all" and gather="owner" against one unsharded Particle run on rank 0:
   positions and cells bit for bit.
Rank 0 prints one line per check and exits non-zero if anything differs.
"""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    import torch.distributed as dist

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, parallel

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    def cparticles(state, n):
        p = _lib.Particles()
        p.n = n
        for k, v in state.items():
            setattr(p, k, v.data_ptr())
        return p

    def state_of(r, step, n):
        """What rank r ships in step `step`: a known function of (rank, step, particle)."""
        i = np.arange(n)
        return {
            "lon": (i * 1e-3 + r + 0.25 * step), "lat": (-i * 2e-3 + step), "dep": -(i % 977) - r * 0.5,
            "face": ((i + r) % 13).astype(np.int32), "iy": ((i * 7 + step) % 4320).astype(np.int32),
            "ix": ((i * 3 + r) % 4320).astype(np.int32), "izl": ((i + step) % 90).astype(np.int32),
        }

    # ---- 1. correctness, both store paths ----
    n, n_pad = 50_000, 50_016
    for multicast in (False, True):
        ex = parallel.RecordExchange(n_pad, dev, slots=2, multicast=multicast)
        if multicast and ex.x.recv_multicast is None:
            say("exchange: no multicast mapping on this box (NVLS not available): multimem path not exercised")
            continue
        perm = torch.from_numpy(np.random.default_rng(7 + rank).permutation(n).astype(np.int32)).to(dev)
        bad = 0
        for step in range(1, 41):
            st = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in state_of(rank, step, n).items()}
            scatter = step % 3 == 0  # every third step through out_index (a cell-sorted state)
            if scatter:
                st = {k: v[perm.long()] for k, v in st.items()}
            s = ex.ship(cparticles(st, n), out_index=perm if scatter else None)
            assert s == step
            ex.wait(s)
            got = ex.slot_view(s).cpu().numpy()
            ex.release(s)
            for r in range(world):
                want = parallel.pack_records_host(state_of(r, step, n), n_pad=n_pad).view(np.uint8).reshape(n_pad, 32)
                if not np.array_equal(got[r][:n], want[:n]):
                    bad += 1
        ex.check()
        t = torch.tensor([bad], device=dev)
        dist.all_reduce(t)
        say(f"exchange ({ex.mode}): 40 steps x {world} blocks on {world} ranks, {int(t.item())} wrong blocks")
        ok &= int(t.item()) == 0
        del ex

    # ---- 2. bandwidth ----
    for n in (() if os.environ.get("CHECK_SKIP_BW") else (1_000_000, 12_500_000)):
        for multicast in (False, True):
            ex = parallel.RecordExchange(n, dev, slots=2, multicast=multicast)
            if multicast and ex.x.recv_multicast is None:
                continue
            st = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in state_of(rank, 1, n).items()}
            cp = cparticles(st, n)
            for _ in range(3):
                s = ex.ship(cp)
                ex.wait(s, release=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                s = ex.ship(cp)
                ex.wait(s, release=True)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ex.check()
            say(f"exchange bandwidth ({ex.mode}), {n} records per rank: {ms.item():.3f} ms per step, "
                f"{32 * n * (world - 1) / (ms.item() * 1e-3) / 1e9:.0f} GB/s received per rank from its peers")
            del ex

    # ---- 3. ShardedParticle against one unsharded run ----
    import cases
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=30, nz=8, nt=4)
    od = sd.OceData(ds)
    n = 20_001
    lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=77, depth_range=(5, 300))
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    stops = [4 * cases.DAY, 12 * cases.DAY, 17 * cases.DAY]
    ref = None
    if rank == 0:
        whole = sd.Particle(x=lon, y=lat, z=dep, t=np.full(n, cases.DAY), data=od, sort=False, **kw)
        _, ref = whole.to_list_of_time(stops, snapshots="records")
    for gather in ("all", "owner"):
        for sort in (False, True):
            sp = parallel.ShardedParticle(lon, lat, dep, np.full(n, cases.DAY), data=od, gather=gather, seed=3, sort=sort, **kw)
            got_stops, snaps = sp.to_list_of_time(stops)
            outs = [s.assemble(dst=0) for s in snaps]
            if rank == 0:
                assert len(outs) == len(ref)
                bad = 0
                for j, (o, r) in enumerate(zip(outs, ref)):
                    for k, attr in (("lon", "lon"), ("lat", "lat"), ("dep", "dep"), ("face", "face"), ("iy", "iy"), ("ix", "ix"), ("izl", "izl_lin")):
                        want = getattr(r, attr)
                        if not np.array_equal(o[k], want, equal_nan=(o[k].dtype.kind == 'f')):
                            bad += 1
                            d = np.flatnonzero((o[k] != want) & ~(np.isnan(o[k].astype(float)) & np.isnan(np.asarray(want, float))))
                            say(f"   snapshot {j} {k}: {d.size} of {n} differ, first at {d[:3]}: got {o[k][d[:3]]} want {want[d[:3]]}")
                say(f"ShardedParticle(gather={gather!r}, sort={sort}): {len(outs)} snapshots of {n} particles on {world} ranks, {bad} arrays differ from the unsharded run")
                ok &= bad == 0
            dist.barrier()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
