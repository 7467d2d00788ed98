"""GPU: the CUDA path (through the seaduck-compatible Python API -> C ABI) against
(a) the golden vectors from the unmodified reference and (b) the CPU oracle on larger seeded inputs.
Index records bit-exact, floats within 1e-9 relative."""

import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sd():
    import seaduck_b200 as sd

    return sd


def log_of(pt):
    return pt.raw_arrays(merged=True)


def compare_log(gold, prefix, rec):
    np.testing.assert_array_equal(rec["offset"], gold[f"{prefix}offset"])
    for k in cases.LOG_INT:
        if f"{prefix}{k}" in gold:
            np.testing.assert_array_equal(rec[k], gold[f"{prefix}{k}"], err_msg=k)
    for k in cases.LOG_FLT:
        if f"{prefix}{k}" in gold:
            cases.assert_close(k, rec[k], gold[f"{prefix}{k}"])


def compare_state(gold, prefix, pt):
    for k in cases.FINAL_INT:
        if f"{prefix}{k}" in gold:
            np.testing.assert_array_equal(getattr(pt, k), gold[f"{prefix}{k}"], err_msg=k)
    for k in cases.FINAL_FLT:
        if f"{prefix}{k}" in gold and gold[f"{prefix}{k}"].dtype.kind == "f":
            cases.assert_close(k, getattr(pt, k), gold[f"{prefix}{k}"])


@pytest.mark.parametrize("make", cases.ALL_CASES, ids=lambda f: f.__name__)
def test_cuda_matches_reference_golden(make):
    sd = _sd()
    case = make()
    gold = np.load(os.path.join(GOLD, f"lagrangian_{case['name']}.npz"))
    od = sd.OceData(case["ds"])
    where = case["seeding"] if "seeding" in case else dict(x=case["x"], y=case["y"], z=case["z"], t=case["t"])
    pt = sd.Particle(data=od, save_raw=True, **where, **case["kw"])
    compare_state(gold, "init_", pt)
    if case.get("list_of_time"):
        stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
        np.testing.assert_allclose(stops, gold["stops"])
        for i, snap in enumerate(snaps):
            compare_log(gold, f"s{i}_log_", log_of(snap))
            compare_state(gold, f"s{i}_", snap)
    else:
        for i, ts in enumerate(case["stops"]):
            pt.empty_lists()
            pt.note_taking(stamp=15)
            pt.to_next_stop(ts)
            compare_log(gold, f"s{i}_log_", log_of(pt))
            compare_state(gold, f"s{i}_", pt)


def test_cuda_matches_oracle_llc90(oracle_lib):
    """Larger seeded run (LLC90-shaped 3-D, 20k particles): CUDA vs the C oracle, full crossing log."""
    sd = _sd()
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=50)
    n = 20000
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=11)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    od = sd.OceData(ds)
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, save_raw=True, **kw)
    op = oracle_lib.oracle_particle_from_od(od, lon, lat, dep, np.zeros(n), save_raw=True, nthreads=8, **kw)
    # nearest-centre search: bucket grid + walk (device) vs cKDTree (oracle)
    same = (pt.face == op.face) & (pt.iy == op.iy) & (pt.ix == op.ix)
    assert same.mean() > 1 - 1e-4, f"locate mismatch fraction {1 - same.mean()}"
    assert same.all(), "a locate tie would make the trajectories below incomparable; change the seed"
    total = 0
    for ts in (30 * cases.DAY, 90 * cases.DAY):
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        op.to_next_stop(ts, emit_start=True)
        rec, ref = log_of(pt), op.records
        np.testing.assert_array_equal(rec["offset"], ref["offset"])
        for k in cases.LOG_INT:
            np.testing.assert_array_equal(rec[k], ref[k], err_msg=k)
        for k in cases.LOG_FLT:
            cases.assert_close(k, rec[k], ref[k])
        total += int((ref["vs"] < 6).sum())
    assert pt.crossings == total
    assert total > 2e5


@pytest.mark.parametrize("scheme", ["rectilinear", "local_cartesian"])
def test_cuda_matches_oracle_on_the_schemes_without_corners(scheme, oracle_lib):
    """The rectilinear (1-D lon / lat, BASELINE.json configs[0]'s grid kind, 10^4 particles as there) and local_cartesian
    (LLC without XG / YG) horizontal schemes at a size the goldens do not reach: CUDA against the C oracle, full crossing
    log, indices bit for bit (reference lagrangian.py:647-704, utils.py:214,577-612)."""
    sd = _sd()
    from seaduck_b200 import minixr, synthetic

    if scheme == "rectilinear":
        ds = synthetic.rectilinear(nlat=100, nlon=100)
        n = 10_000
        rng = np.random.default_rng(31)
        lon, lat, dep = rng.uniform(-179.5, 179.5, n), rng.uniform(-70.0, 70.0, n), np.full(n, -1.0)
        kw = dict(uname="u", vname="v", wname=None, transport=False)
        stops = (60 * cases.DAY, 20 * cases.DAY)
    else:
        full = synthetic.llc(n=45, nz=12)
        keep = {k: (tuple(full[k].dims), np.asarray(full[k])) for k in full.variables if k not in ("XG", "YG")}
        dims = {d for dd, _ in keep.values() for d in dd}
        ds = minixr.Dataset(coords={k: v for k, v in keep.items() if k in dims}, data_vars={k: v for k, v in keep.items() if k not in dims})
        n = 10_000
        lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=32, depth_range=(5, 600))
        kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
        stops = (30 * cases.DAY, 10 * cases.DAY)
    od = sd.OceData(ds)
    assert od.readiness["h"] == scheme
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, save_raw=True, **kw)
    op = oracle_lib.oracle_particle_from_od(od, lon, lat, dep, np.zeros(n), save_raw=True, nthreads=8, **kw)
    same = (pt.iy == op.iy) & (pt.ix == op.ix)
    assert same.all(), f"{int((~same).sum())} particles located in another cell"
    for k, ref in (("rx", op.rx), ("ry", op.ry), ("u", op.u), ("v", op.v), ("du", op.du), ("dv", op.dv)):
        cases.assert_close(k, getattr(pt, k), ref)
    total = 0
    for ts in stops:
        pt.empty_lists()
        pt.note_taking(stamp=15)
        pt.to_next_stop(ts)
        op.to_next_stop(ts, emit_start=True)
        rec, ref = log_of(pt), op.records
        np.testing.assert_array_equal(rec["offset"], ref["offset"])
        for k in cases.LOG_INT:
            np.testing.assert_array_equal(rec[k], ref[k], err_msg=k)
        for k in cases.LOG_FLT:
            cases.assert_close(k, rec[k], ref[k])
        total += int((ref["vs"] < 6).sum())
    assert pt.crossings == total and total > 3e4, total


def test_lazy_mode_equals_eager():
    """``Particle.lazy``: end-of-stop bookkeeping on the device, same state and counts as the eager path."""
    sd = _sd()
    case = cases.case_llc_transport()
    pts = []
    for lazy in (False, True):
        od = sd.OceData(case["ds"])
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, **case["kw"])
        pt.lazy = lazy
        for ts in case["stops"]:
            pt.to_next_stop(ts)
        pt.to_next_stop(case["stops"][-1])  # nothing left to simulate: must leave everything alone
        pts.append(pt)
    eager, lazy = pts
    assert lazy.crossings == eager.crossings and eager.crossings > 0
    for k in cases.FINAL_INT + cases.FINAL_FLT:
        if getattr(eager, k, None) is not None:
            np.testing.assert_array_equal(getattr(lazy, k), getattr(eager, k), err_msg=k)


def test_lazy_mode_list_of_time_matches_golden():
    """Time-dependent field, slab swaps at the update stops, lazy bookkeeping: still the reference's answer."""
    sd = _sd()
    case = cases.case_llc_time()
    gold = np.load(os.path.join(GOLD, f"lagrangian_{case['name']}.npz"))
    od = sd.OceData(case["ds"])
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, **case["kw"])
    pt.lazy = True
    # streaming ingest: the snapshot a leg ends in is uploaded on a side stream while the leg runs, and
    # update_uvw_array then finds it on the device
    calls, hits = [], []
    prefetch, device_field = od.prefetch_field, od.device_field

    def counting_prefetch(name, sl):
        calls.append((name, sl.start))
        return prefetch(name, sl)

    def counting_device_field(name, sl=None):
        key = od._slab_key(name, sl)
        hits.append(isinstance(od._field_cache.get(key), tuple))
        return device_field(name, sl)

    od.prefetch_field, od.device_field = counting_prefetch, counting_device_field
    stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
    assert len(calls) >= 3 and any(hits), (calls, hits)
    np.testing.assert_allclose(stops, gold["stops"])
    for i, snap in enumerate(snaps):
        compare_state(gold, f"s{i}_", snap)


def test_surface_profile_equals_generic_kernel_and_golden():
    """The surface run of BASELINE.json configs[3] takes its own compile-time variant of the advection kernel (profile 3:
    no vertical axis, five CTAs per SM); CTAs of 256 threads are outside that variant's launch bounds and fall back to the
    generic kernel (profile 0).  Same inputs through both: every state array bit for bit, and both the reference's
    answer at every stop (golden ``lagrangian_llc_surface_time``, which the golden test itself reaches through the
    logging variant of profile 0)."""
    import ctypes as C

    sd = _sd()
    from seaduck_b200 import _lib

    case = cases.case_llc_surface_time()
    gold = np.load(os.path.join(GOLD, f"lagrangian_{case['name']}.npz"))
    runs = {}
    for threads, want in ((0, 3), (256, 0)):
        od = sd.OceData(case["ds"])
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, **case["kw"])
        pt.tuning = dict(threads_per_block=threads)
        par = pt._params(pt.max_iteration)
        p = pt._cparticles()
        assert _lib.lib().sdk_advect_profile(od.grid_handle(), C.byref(pt._vel), C.byref(p), C.byref(par), None) == want
        stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
        np.testing.assert_allclose(stops, gold["stops"])
        for i, snap in enumerate(snaps):
            compare_state(gold, f"s{i}_", snap)
        runs[want] = (pt, snaps)
    assert runs[3][0].crossings == runs[0][0].crossings > 0
    # record snapshots of a lazy run, with and without the cell sort
    for sort in (False, True):
        od = sd.OceData(case["ds"])
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=sort, **case["kw"])
        pt.lazy = True
        _, recs = pt.to_list_of_time(normal_stops=case["stops"], snapshots="records")
        assert len(recs) == len(runs[3][1])
        for full, r in zip(runs[3][1], recs):
            for k in ("lon", "lat", "face", "iy", "ix"):
                np.testing.assert_array_equal(getattr(r, k), getattr(full, k), err_msg=f"{k} (records, sort={sort})")
    for a, b in zip(runs[3][1] + [runs[3][0]], runs[0][1] + [runs[0][0]]):
        for k in cases.FINAL_INT + cases.FINAL_FLT:
            if getattr(a, k, None) is not None:
                np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)


def _cparticles_of(state, n):
    from seaduck_b200 import _lib

    p = _lib.Particles()
    p.n = n
    for k, v in state.items():
        setattr(p, k, None if v is None else v.data_ptr())
    return p


def test_pack_records_device_equals_host():
    """sdk_pack_records (one launch) writes the same 32-byte records as the numpy restatement, also scattered."""
    import torch

    from seaduck_b200 import parallel

    rng = np.random.default_rng(0)
    n = 1000
    host = {
        "lon": rng.uniform(-180, 180, n), "lat": rng.uniform(-90, 90, n), "dep": rng.uniform(-5000, 0, n),
        "face": rng.integers(0, 13, n).astype(np.int32), "iy": rng.integers(0, 4320, n).astype(np.int32),
        "ix": rng.integers(0, 4320, n).astype(np.int32), "izl": rng.integers(0, 90, n).astype(np.int32),
    }
    dev = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
    a = parallel.pack_records_host(host)
    b = parallel.pack_records(_cparticles_of(dev, n), n, device="cuda").cpu().numpy()
    np.testing.assert_array_equal(a.view(np.uint8).reshape(n, 32), b)
    perm = torch.from_numpy(rng.permutation(n).astype(np.int32)).cuda()
    c = parallel.pack_records(_cparticles_of(dev, n), n, out_index=perm, device="cuda").cpu().numpy()
    np.testing.assert_array_equal(c[perm.cpu().numpy()], b)
    # 2-D runs have no depth / level, single-face grids no face
    dev2 = dict(dev, dep=None, izl=None, face=None)
    d = parallel.unpack_records(parallel.pack_records(_cparticles_of(dev2, n), n, device="cuda"))
    assert (d["dep"] == 0).all() and (d["izl"] == 0).all() and (d["face"] == 0).all()
    np.testing.assert_array_equal(d["ix"], host["ix"])


def test_record_exchange_on_one_rank():
    """sdk_pack_ship / sdk_exchange_wait with world = 1 (the rank is its own peer): records land in the slot of the
    step, the flags count up, a slot is reused once it has been released."""
    import ctypes as C

    import torch

    from seaduck_b200 import _lib, parallel

    rng = np.random.default_rng(5)
    n, n_pad, slots = 3000, 3072, 2
    L = _lib.lib()
    buf = torch.zeros(4096 + slots * n_pad * 32, dtype=torch.uint8, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = _lib.Exchange()
    x.world, x.rank, x.n_slots, x.n_pad = 1, 0, slots, n_pad
    x.signal[0], x.recv[0], x.ticket = buf.data_ptr(), buf.data_ptr() + 4096, ticket.data_ptr()
    stream = _lib.current_stream()
    for step in range(1, 6):
        host = {
            "lon": rng.uniform(-180, 180, n), "lat": rng.uniform(-90, 90, n), "dep": rng.uniform(-5000, 0, n),
            "face": rng.integers(0, 13, n).astype(np.int32), "iy": rng.integers(0, 90, n).astype(np.int32),
            "ix": rng.integers(0, 90, n).astype(np.int32), "izl": rng.integers(0, 50, n).astype(np.int32),
        }
        dev = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
        _lib.check(L.sdk_pack_ship(C.byref(x), C.byref(_cparticles_of(dev, n)), None, step, stream), "sdk_pack_ship")
        _lib.check(L.sdk_exchange_wait(C.byref(x), step, 1, 0, stream), "sdk_exchange_wait")
        slot = step % slots
        got = buf[4096 + slot * n_pad * 32 : 4096 + (slot + 1) * n_pad * 32].cpu().numpy().reshape(n_pad, 32)
        np.testing.assert_array_equal(got, parallel.pack_records_host(host, n_pad=n_pad).view(np.uint8).reshape(n_pad, 32))
        _lib.check(L.sdk_exchange_wait(C.byref(x), step, 0, 1, stream), "sdk_exchange_wait")
        words = buf[: _lib.SIGNAL_WORDS * 8].cpu().numpy().view(np.uint64)
        assert words[_lib.SIGNAL_ARRIVED] == step and words[_lib.SIGNAL_CONSUMED] == step and words[_lib.SIGNAL_ERROR] == 0
    assert int(ticket.item()) == 0


@pytest.mark.parametrize("sort", [False, True])
def test_fused_ship_equals_pack_records(sort):
    """sdk_advect with sdk_advect_params.ship (world = 1: the rank is its own peer): the records the advection kernel
    stores while it writes particles back are the records sdk_pack_records packs afterwards, stop after stop."""
    import ctypes as C

    import torch

    from seaduck_b200 import _lib, parallel

    sd = _sd()
    case = cases.case_llc_transport(n=3000)
    od = sd.OceData(case["ds"])
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=sort, **case["kw"])
    plain = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=False, **case["kw"])
    n, n_pad, slots = pt.N, 3008, 2

    class OneRank:
        def __init__(self):
            self.buf = torch.zeros(4096 + slots * n_pad * 32, dtype=torch.uint8, device="cuda")
            self.ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
            self.x = _lib.Exchange()
            self.x.world, self.x.rank, self.x.n_slots, self.x.n_pad = 1, 0, slots, n_pad
            self.x.signal[0], self.x.recv[0], self.x.ticket = self.buf.data_ptr(), self.buf.data_ptr() + 4096, self.ticket.data_ptr()
            self.step = 0

        def next_step(self):
            self.step += 1
            return self.step

    ex = OneRank()
    pt.ship_to(ex)
    L = _lib.lib()
    for k, ts in enumerate([5 * cases.DAY, 20 * cases.DAY, 20 * cases.DAY, 60 * cases.DAY, 31 * cases.DAY]):
        pt.to_next_stop(ts)
        plain.to_next_stop(ts)
        step = pt.last_ship_step
        assert step == k + 1
        _lib.check(L.sdk_exchange_wait(C.byref(ex.x), step, 1, 1, _lib.current_stream()), "sdk_exchange_wait")
        slot = step % slots
        got = ex.buf[4096 + slot * n_pad * 32 : 4096 + (slot + 1) * n_pad * 32].view(n_pad, 32)[:n]
        want = parallel.pack_records(plain._cparticles(), n, device="cuda")
        np.testing.assert_array_equal(got.cpu().numpy(), want.cpu().numpy(), err_msg=f"stop {k}")
    words = ex.buf[: _lib.SIGNAL_WORDS * 8].cpu().numpy().view(np.uint64)
    assert words[_lib.SIGNAL_ARRIVED] == 5 and words[_lib.SIGNAL_CONSUMED] == 5 and words[_lib.SIGNAL_ERROR] == 0
    for k in cases.FINAL_INT + cases.FINAL_FLT:
        if getattr(plain, k, None) is not None:
            np.testing.assert_array_equal(getattr(pt, k), getattr(plain, k), err_msg=k)


@pytest.mark.parametrize("n", [1000, 300_000])
def test_track_host_equals_particle_api(n):
    """sdk_track_host (host buffers in and out; chunked and pipelined over three streams for large
    batches) gives exactly what Particle(...).to_next_stop(...) gives -- as separate arrays, as 32-byte records,
    and with the particles of every chunk cell-sorted on the device."""
    import ctypes as C

    import torch

    from seaduck_b200 import _lib, parallel, synthetic

    sd = _sd()
    ds = synthetic.llc(n=45, nz=12)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=8, depth_range=(5, 600))
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    od = sd.OceData(ds)
    t_stop = 20 * cases.DAY
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, **kw)
    pt.to_next_stop(t_stop)
    L = _lib.lib()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    hin = [pin(lon), pin(lat), pin(dep), pin(np.zeros(n))]
    outf = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    outi = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(6)]
    rec = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    stats = torch.zeros(4, dtype=torch.int64).pin_memory()
    scratch = torch.empty(int(L.sdk_track_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
    par = pt._params(200)
    par.stats = None

    def run(io):
        _lib.check(
            L.sdk_track_host(od.grid_handle(), C.byref(pt._vel), C.byref(io), None, 0, float(t_stop), C.byref(par), scratch.data_ptr(), _lib.current_stream()),
            "sdk_track_host",
        )

    io = _lib.TrackIO()
    io.n = n
    io.lon, io.lat, io.dep, io.t = (h.data_ptr() for h in hin)
    io.out_lon, io.out_lat, io.out_dep = (h.data_ptr() for h in outf)
    io.out_face, io.out_iy, io.out_ix, io.out_izl, io.out_ncross, io.out_status = (h.data_ptr() for h in outi)
    io.out_records, io.out_stats = rec.data_ptr(), stats.data_ptr()
    run(io)
    for got, name in zip(outf, ("lon", "lat", "dep")):
        np.testing.assert_array_equal(got.numpy(), getattr(pt, name), err_msg=name)
    for got, name in zip(outi[:4], ("face", "iy", "ix", "izl_lin")):
        np.testing.assert_array_equal(got.numpy(), getattr(pt, name), err_msg=name)
    assert int(outi[4].numpy().sum()) == pt.crossings == int(stats[1])
    for sort in (0, 1):
        io2 = _lib.TrackIO()
        io2.n, io2.sort = n, sort
        io2.lon, io2.lat, io2.dep, io2.t = (h.data_ptr() for h in hin)
        rec.zero_()
        stats.zero_()
        io2.out_records, io2.out_stats = rec.data_ptr(), stats.data_ptr()
        run(io2)
        back = parallel.unpack_records(rec.numpy())
        for name, attr in (("lon", "lon"), ("lat", "lat"), ("dep", "dep"), ("face", "face"), ("iy", "iy"), ("ix", "ix"), ("izl", "izl_lin")):
            np.testing.assert_array_equal(back[name], getattr(pt, attr), err_msg=f"{name} (sort={sort})")
        assert int(stats[1]) == pt.crossings
    # the sorted path refuses the separate arrays
    io.sort = 1
    assert L.sdk_track_host(od.grid_handle(), C.byref(pt._vel), C.byref(io), None, 0, float(t_stop), C.byref(par), scratch.data_ptr(), _lib.current_stream()) == -1


def test_cell_sorted_particles_equal_plain():
    """Particle(sort=True) keeps the device state in cell order; every attribute, snapshot and interpolation still
    comes back in the caller's order and bit for bit equal to the unsorted run."""
    sd = _sd()
    case = cases.case_llc_time(n=4000)
    runs = []
    for sort in (False, True):
        od = sd.OceData(case["ds"])
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=sort, **case["kw"])
        assert (pt._perm is not None) == sort
        stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
        stops_r, recs = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=sort, **case["kw"]).to_list_of_time(
            normal_stops=case["stops"], snapshots="records")
        # the same through the lazy path (end-of-stop bookkeeping on the device)
        lazy = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, sort=sort, **case["kw"])
        lazy.lazy = True
        _, recs_lazy = lazy.to_list_of_time(normal_stops=case["stops"], snapshots="records")
        assert lazy.crossings == pt.crossings
        for full, r in zip(snaps, recs_lazy):
            for k in ("lon", "lat", "dep", "face", "iy", "ix", "izl_lin"):
                np.testing.assert_array_equal(getattr(r, k), getattr(full, k), err_msg=f"{k} (lazy, sort={sort})")
        assert len(recs_lazy) == len(snaps)
        tvals = pt.interpolate("T", sd.KnW()) if "T" in case["ds"].variables else None
        runs.append((pt, snaps, recs, tvals))
    (a, sa, ra, ta), (b, sb, rb, tb) = runs
    assert a.crossings == b.crossings > 0
    if sort:
        key = (b._st["izl"].long() * 13 + b._st["face"].long())
        assert int(b._perm.max()) == b.N - 1 and key.shape[0] == b.N
    for x, y in zip(sa + [a], sb + [b]):
        for k in cases.FINAL_INT + cases.FINAL_FLT + ["px", "py", "vol", "dx", "dy", "bzl_lin", "dzl_lin"]:
            if getattr(x, k, None) is not None:
                np.testing.assert_array_equal(getattr(y, k), getattr(x, k), err_msg=k)
    # record snapshots: same positions and cells as the full clones, in both orders
    assert len(ra) == len(rb) == len(sa)
    for full, r1, r2 in zip(sa, ra, rb):
        for k in ("lon", "lat", "dep", "face", "iy", "ix", "izl_lin"):
            np.testing.assert_array_equal(getattr(r1, k), getattr(full, k), err_msg=k)
            np.testing.assert_array_equal(getattr(r2, k), getattr(full, k), err_msg=k)
        assert r1.N == full.N and (r1.t == full.t).all()
        with pytest.raises(AttributeError):
            r1.rx
    if ta is not None:
        np.testing.assert_array_equal(ta, tb)
    # writing an attribute goes through the permutation as well
    new_lon = a.lon + 0.125
    a.lon, b.lon = new_lon, new_lon
    np.testing.assert_array_equal(b._st["lon"][b._inv].cpu().numpy(), new_lon)
    np.testing.assert_array_equal(b.lon, a.lon)


def test_sharded_run_equals_single_run():
    """disjoint particle slices advected separately and un-sharded give bit for bit what
    one run over all particles gives (particles do not interact)."""
    import torch

    from seaduck_b200 import parallel, synthetic

    sd = _sd()
    n, world = 5000, 4
    ds = synthetic.llc(n=30, nz=8)
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=9, depth_range=(5, 300))
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    od = sd.OceData(ds)
    t_stop = 30 * cases.DAY
    whole = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, **kw)
    whole.to_next_stop(t_stop)
    per = -(-n // world)
    blocks = []
    for r in range(world):
        idx = parallel.shard_indices(n, world, r, seed=5)
        pt = sd.Particle(x=lon[idx], y=lat[idx], z=dep[idx], t=np.zeros(len(idx)), data=od, sort=bool(r % 2), **kw)
        pt.to_next_stop(t_stop)
        rec = torch.zeros((per, 32), dtype=torch.uint8, device="cuda")
        parallel.pack_records(pt._cparticles(), pt.N, out=rec, out_index=pt._out_index())
        blocks.append(rec)
    out = parallel.unshard(torch.stack(blocks), n, world, seed=5)
    for name, ref in (("lon", whole.lon), ("lat", whole.lat), ("dep", whole.dep), ("face", whole.face), ("iy", whole.iy),
                      ("ix", whole.ix), ("izl", whole.izl_lin)):
        np.testing.assert_array_equal(out[name], ref, err_msg=name)


@pytest.mark.parametrize("sort", [False, True])
def test_sharded_particle_api_on_one_rank(sort):
    """parallel.ShardedParticle without a process group (one rank owns every particle, in the order of the sharding
    permutation): to_list_of_time over a time-dependent field gives the snapshots of a plain Particle run."""
    from seaduck_b200 import parallel, synthetic

    sd = _sd()
    ds = synthetic.llc(n=30, nz=8, nt=4)
    od = sd.OceData(ds)
    n = 5001
    lon, lat, dep = synthetic.llc_seed_particles(None, n, seed=77, depth_range=(5, 300))
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    stops = [4 * cases.DAY, 12 * cases.DAY, 17 * cases.DAY]
    whole = sd.Particle(x=lon, y=lat, z=dep, t=np.full(n, cases.DAY), data=od, sort=False, **kw)
    ref_stops, ref = whole.to_list_of_time(stops, snapshots="full")
    sp = parallel.ShardedParticle(lon, lat, dep, np.full(n, cases.DAY), data=od, gather="all", seed=3, sort=sort, **kw)
    got_stops, snaps = sp.to_list_of_time(stops)
    np.testing.assert_allclose(got_stops, ref_stops)
    assert len(snaps) == len(ref)
    for i, (s, r) in enumerate(zip(snaps, ref)):
        out = s.assemble()
        assert s.t == float(r.t[0])
        for k, attr in (("face", "face"), ("iy", "iy"), ("ix", "ix"), ("izl", "izl_lin"), ("lon", "lon"), ("lat", "lat"), ("dep", "dep")):
            np.testing.assert_array_equal(out[k], getattr(r, attr), err_msg=f"snapshot {i}: {k}")


def test_full_size_round_trip_and_idempotence():
    """BASELINE.json configs[1] at full size (LLC90 x 50 levels, 10^6 particles): size-independent
    properties of the analytic solver in a steady field -- going to a stop twice changes nothing, and
    30 days forward followed by 30 days backward returns (almost) every particle to where it started,
    through the same number of walls."""
    sd = _sd()
    from seaduck_b200 import synthetic

    ds = synthetic.llc(n=90, nz=50)
    n = 1_000_000
    lon, lat, dep = synthetic.llc_seed_particles(ds, n, seed=21)
    od = sd.OceData(ds)
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    start = {k: getattr(pt, k).copy() for k in ("lon", "lat", "dep", "face", "iy", "ix", "izl_lin")}
    pt.to_next_stop(30 * cases.DAY)
    forward = pt.crossings
    assert forward > 1e7
    there = {k: getattr(pt, k).copy() for k in ("lon", "lat", "dep", "ix")}
    pt.to_next_stop(30 * cases.DAY)  # nothing left to simulate
    assert pt.crossings == forward
    for k, v in there.items():
        np.testing.assert_array_equal(getattr(pt, k), v, err_msg=k)
    assert (np.abs(to180(there["lon"] - start["lon"])) > 1e-3).mean() > 0.9  # they did move
    pt.to_next_stop(0.0)
    backward = pt.crossings - forward
    status = pt._st["status"].cpu().numpy()
    # (a dozen seeds fall into the degenerate land cells at the 3-face junctions of the cap, where two
    # corners coincide and the reference's inverse bilinear map returns inf / nan as well)
    clean = (status == 0) & np.isfinite(pt.lon) & np.isfinite(there["lon"])
    assert clean.mean() > 0.99
    back_lon = np.abs(to180(pt.lon - start["lon"]))[clean]
    back_lat = np.abs(pt.lat - start["lat"])[clean]
    back_dep = np.abs(pt.dep - start["dep"])[clean]
    same_cell = ((pt.face == start["face"]) & (pt.iy == start["iy"]) & (pt.ix == start["ix"]) & (pt.izl_lin == start["izl_lin"]))[clean]
    # (not every trajectory is reversible in floating point: a particle that converges onto a stagnation
    # point next to land forgets how long it sat there; ~1.5 % of this field's trajectories do)
    q = [float(np.quantile(v, 0.9)) for v in (back_lon, back_lat, back_dep)]
    assert q[0] < 1e-6 and q[1] < 1e-6 and q[2] < 1e-3, (q, float(same_cell.mean()))
    assert same_cell.mean() > 0.97, float(same_cell.mean())
    assert abs(backward - forward) < 2e-2 * forward, (forward, backward)


def to180(x):
    return (x + 180.0) % 360.0 - 180.0


def test_stuck_particle_callbacks():
    """get_masks.which_not_stuck / abandon_stuck (reference get_masks.py:219,231): the mask of the cell a particle sits
    in, as a ``callback`` that takes grounded particles out of the integration."""
    sd = _sd()
    from seaduck_b200 import get_masks

    import cases_interp as ci

    case = ci.case_llc3d(n=600)
    ds = case["ds"]
    od = sd.OceData(ds)
    kw = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)
    t0 = np.zeros(len(case["x"]))
    pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=t0, data=od, callback=get_masks.which_not_stuck, **kw)
    wet = get_masks.which_not_stuck(pt)
    maskc = np.asarray(ds["maskC"])
    np.testing.assert_array_equal(wet, maskc[pt.izl_lin - 1, pt.face, pt.iy, pt.ix] != 0)
    assert 0 < wet.sum() < pt.N  # both kinds are present
    start = pt.lon.copy()
    pt.to_next_stop(5 * cases.DAY)
    moved = pt.lon != start
    assert not moved[~wet].any() and moved[wet].mean() > 0.5  # grounded particles stay where they are
    sub = pt.subset(np.arange(pt.N))
    get_masks.abandon_stuck(sub)
    assert sub.N == int(get_masks.which_not_stuck(pt).sum()) == len(sub.lon) == len(sub.ix)
    mc, mu, mv, mw = get_masks.get_mask_arrays(od)
    assert mc.shape == maskc.shape and mu.ndim == mv.ndim == mw.ndim == 4
    common = tuple(slice(0, min(a, b)) for a, b in zip(mc.shape, mu.shape))
    assert ((mu[common] != 0) | (mc[common] == 0)).all()  # the west wall of a wet cell is wet


def test_fields_left_in_host_memory_give_the_same_run():
    """Out-of-core fields (the reference's ``too_large`` route): data variables that are not uploaded but read in place
    from page-locked host memory -- forced by name (``keep_on_host``) or by size (``device_memory_limit``) -- give bit
    for bit the run and the interpolated values of fields resident in HBM, time slabs included."""
    import torch

    sd = _sd()
    case = cases.case_llc_time(n=3000)
    runs = []
    for how in ("resident", "by name", "by size"):
        od = sd.OceData(case["ds"], device_memory_limit=1 if how == "by size" else None)
        if how == "by name":
            od.keep_on_host(case["kw"]["uname"], case["kw"]["vname"], case["kw"]["wname"], "T")
        pt = sd.Particle(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od, **case["kw"])
        stops, snaps = pt.to_list_of_time(normal_stops=case["stops"])
        tvals = pt.interpolate("T", sd.KnW()) if "T" in case["ds"].variables else None
        where = {k[0]: v.device.type for k, v in od._field_cache.items() if isinstance(v, torch.Tensor)}
        runs.append((pt, snaps, tvals, where))
    (a, sa, ta, wa), (b, sb, tb, wb), (c, sc, tc, wc) = runs
    assert set(wa.values()) == {"cuda"} and wb[case["kw"]["uname"]] == "cpu" and set(wc.values()) == {"cpu"}
    assert a.crossings == b.crossings == c.crossings > 0
    for other, snaps in ((b, sb), (c, sc)):
        for x, y in zip(sa + [a], snaps + [other]):
            for k in cases.FINAL_INT + cases.FINAL_FLT:
                if getattr(x, k, None) is not None:
                    np.testing.assert_array_equal(getattr(y, k), getattr(x, k), err_msg=k)
    if ta is not None:
        np.testing.assert_array_equal(tb, ta)
        np.testing.assert_array_equal(tc, ta)
    # Eulerian interpolation of host-resident tracers and velocities (time-dependent, masked and vector kernels)
    import cases_interp as ci

    case = ci.case_llc3d(n=2000)
    res = []
    for limit in (None, 1):
        od = sd.OceData(case["ds"], device_memory_limit=limit)
        pos = sd.Position().from_latlon(x=case["x"], y=case["y"], z=case["z"], t=case["t"], data=od)
        res.append(ci.flatten_result(pos.interpolate(["T", "S", ("UVELMASS", "VVELMASS")], [sd.KnW(), sd.KnW(vkernel="linear", tkernel="linear"), (sd.KnW(), sd.KnW())])))
    for x, y in zip(*res):
        np.testing.assert_array_equal(x, y)
