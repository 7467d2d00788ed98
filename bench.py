#!/usr/bin/env python
"""Headline benchmark: particle cell-crossings per second (BASELINE.json `metric`).

Workload (BASELINE.json configs[1]): a 3-D LLC90-shaped synthetic field (13 faces x 90 x 90 x 50
levels, volume transports, fp64) and batches of 10^6 particles per GPU seeded uniformly on the
sphere, so trajectories cross cell walls, vertical levels and LLC face edges.

One *step* = one pass of the hot path over one batch: ``Particle.to_next_stop`` (one launch of the
fused advection kernel) carries a fresh batch of 10^6 located particles `--days` days forward.
Every step works on a different batch (all batches are resident in HBM before the timed region;
fields 168 MB + 220 MB of state per batch >> L2, so nothing is served from a warm cache).
A crossing is an analytic step that ended on a cell wall (``tend < 6``), counted on the device
exactly like the instrumented reference does (wrap ``cross_cell_wall``, sum ``tend < 6``).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line.  `value` = crossings/s with the state resident in HBM; `e2e` = the same
metric through the C-ABI host-buffer call ``sdk_track_host`` (pinned host lon/lat/dep/t -> H2D ->
locate -> velocity read -> advect -> D2H of positions and indices, all inside the timed region).
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DAY = 86400.0
ALGO_BYTES_PER_CROSSING = 88  # 6 wall transports + Vol (fp64) + 4 corners x (XG,YG) f32
STATE_BYTES_PER_PARTICLE = 440  # 220 B of state read + 220 B written once per launch
KW = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU and batch")
    ap.add_argument("--days", type=float, default=30.0, help="simulated days per step")
    ap.add_argument("--grid", type=int, default=90)
    ap.add_argument("--levels", type=int, default=50, help="vertical levels; 0 = surface-only 2-D run (configs[3] shape)")
    ap.add_argument("--snapshots", type=int, default=0, help="time snapshots of the velocity field (0 = steady)")
    ap.add_argument("--save-raw", action="store_true", help="write the reference's crossing log (configs[4]): count pass + fill pass per step")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="field storage (math is fp64)")
    ap.add_argument("--no-sort", action="store_true", help="keep particles in seeding order")
    ap.add_argument("--max-batches", type=int, default=24, help="distinct batches kept in HBM (reused cyclically)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="particles of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sweep", default="", help="tuning aid: chunks:blocks_per_sm,... for sdk_track_host")
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--peer-gather", action="store_true", help="ship snapshots with copy-engine writes into peer memory instead of NCCL")
    ap.add_argument("--eager", action="store_true", help="reference-style to_next_stop (host reads the call statistics every step)")
    ap.add_argument("--no-gather", action="store_true", help="(diagnostic) skip the snapshot all-gather at N > 1")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs the advection kernel leaves to the snapshot all-gather (default 0: measured best, NCCL runs in the kernel tails)")
    return ap.parse_args()


def workload_name(a):
    shape = f"3D synthetic transports (13x{a.grid}x{a.grid}x{a.levels})" if a.levels else f"surface-only 2D synthetic transports (13x{a.grid}x{a.grid})"
    extra = (f", {a.snapshots} time snapshots" if a.snapshots else "") + (", crossing log written (save_raw)" if a.save_raw else "")
    return f"LLC{a.grid}-shaped {shape}{extra}, {a.particles} particles per GPU and step, {a.days:g} d per step"


def kwargs_of(a):
    return dict(KW) if a.levels else dict(uname="utrans", vname="vtrans", wname=None, transport=True)


def make_dataset(a):
    from seaduck_b200 import synthetic

    return synthetic.llc(n=a.grid, nz=a.levels, nt=a.snapshots, dtype=np.float64 if a.dtype == "f64" else np.float32)


def seeds(ds, n, seed):
    from seaduck_b200 import synthetic

    return synthetic.llc_seed_particles(ds, n, seed=seed)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, via NVML (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        nv = self.nv
        while nv is not None and not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        reasons = [k for k, bit in names.items() if any(s[2] & bit for s in self.samples)]
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": float(max(s[1] for s in self.samples)),
                "reasons": reasons, "samples": len(self.samples)}


def oracle_particle(a, n, threads, seed=99):
    from oracle import oracle
    from seaduck_b200.ocedata import OceData

    ds = make_dataset(a)
    od = OceData(ds)
    lon, lat, dep = seeds(ds, n, seed)
    return oracle.oracle_particle_from_od(od, lon, lat, dep if a.levels else np.full(n, -1.0), np.zeros(n), nthreads=threads, **kwargs_of(a))


def cpu_baseline(a, threads):
    """Time the CPU oracle (C port of the reference loop) on a bounded sample of the same workload."""
    op = oracle_particle(a, a.cpu_sample, threads)
    t0 = time.perf_counter()
    c = op.to_next_stop(a.days * DAY)
    return c, time.perf_counter() - t0


def run_reference(a):
    """`--impl reference`: the reference's CPU algorithm on all host threads.  The reference itself
    is Python and cannot travel to the GPU box, so this is its C port (oracle/sd_oracle.c, pinned
    to the reference by golden vectors); each step is a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    n = a.cpu_sample
    total, elapsed = 0, 0.0
    for i in range(a.warmup + a.steps):
        op = oracle_particle(a, n, threads, seed=99 + i)  # a fresh batch per step, like the CUDA arm
        t0 = time.perf_counter()
        c = op.to_next_stop(a.days * DAY)
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            total += c
            elapsed += dt
    value = total / elapsed
    sample = (f"{n} particles x {a.days:g} d per step (same field, same seeding law), C port of the reference loop "
              f"(oracle/sd_oracle.c), {threads} threads")
    line = {
        "impl": "reference", "metric": "particle cell-crossings/sec", "value": value, "unit": "crossings/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * elapsed / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "crossings/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "crossings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_cuda(a):
    import ctypes as C

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    reserve = max(a.reserve_sms, 0)
    if world > 1:
        if reserve > 0:
            # one NCCL channel = one CTA: no more channels than SMs the advection kernel leaves free,
            # or the collective's remaining CTAs would have to wait for the end of that kernel
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(reserve))
        dist.init_process_group("nccl", device_id=dev)
    import seaduck_b200 as sd
    from seaduck_b200 import _lib

    L = _lib.lib()
    ds = make_dataset(a)
    od = sd.OceData(ds)
    handle = od.grid_handle()
    n = a.particles
    t_stop = a.days * DAY
    nb = min(a.warmup + a.steps, a.max_batches)

    side = torch.cuda.Stream(dev) if world > 1 else None

    def make_batch(b):
        lon, lat, dep = seeds(ds, n, seed=1000 + 97 * rank + b)  # disjoint particle slices per rank / batch
        if not a.no_sort:
            # cell-major order so that the lanes of a warp touch neighbouring memory (K5)
            loc = od.locate(x=lon, y=lat, z=dep if a.levels else None)
            lev = loc["izl"].long() if a.levels else 0
            key = ((lev * 13 + loc["face"].long()) * a.grid + loc["iy"].long()) * a.grid + loc["ix"].long()
            order = torch.argsort(key).cpu().numpy()
            lon, lat, dep = lon[order], lat[order], dep[order]
        pt = sd.Particle(x=lon, y=lat, z=dep if a.levels else np.full(n, -1.0), t=np.zeros(n), data=od, save_raw=a.save_raw, **kwargs_of(a))
        pt.lazy = not a.eager  # device-side end-of-stop bookkeeping: the host never waits inside a step
        pt.tuning = dict(refill_threshold=a.refill, blocks_per_sm=a.blocks_per_sm, threads_per_block=a.threads)
        if reserve > 0:
            # one 512-thread CTA per SM on all but `reserve` SMs: the NCCL kernel that ships the previous
            # step's snapshot runs on the SMs left free
            pt.tuning.update(blocks_per_sm=1, threads_per_block=512, reserve_sms=reserve)
        return pt, (lon, lat, dep)

    batches, host_inputs = [], []
    for b in range(nb):
        pt, xyz = make_batch(b)
        batches.append(pt)
        if b == 0:
            host_inputs = xyz
    # initial state of every batch: restored (device copy) when a batch is reused
    reuse = a.warmup + a.steps > nb
    saved = [{k: (None if v is None else v.clone()) for k, v in pt._st.items()} for pt in batches]
    from seaduck_b200 import parallel

    gatherer = parallel.SnapshotGather(parallel.SNAPSHOT_ROWS, n, dev, use_peer=a.peer_gather) if world > 1 else None

    def restore(b):
        for k, v in saved[b].items():
            if v is not None:
                batches[b]._st[k].copy_(v)

    def step(i):
        b = i % nb
        pt = batches[b]
        if i >= nb:
            restore(b)
        if a.save_raw:
            pt.empty_lists()  # the log of this step only (the reference's to_list_of_time does the same)
        pt.to_next_stop(t_stop)
        if world > 1 and not a.no_gather:
            # output trajectories of all ranks -- the only exchange of the path: shipped over
            # NVLink peer memory by the copy engines while the next step's kernel runs
            # (side stream: packing and the collective stay off the critical path of the next step)
            done = torch.cuda.Event()
            done.record()
            side.wait_event(done)
            with torch.cuda.stream(side):
                gatherer.post(parallel.pack_snapshot(pt._st), slot=i % 2)

    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    c0 = sum(pt.crossings for pt in batches)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    t_start.record()
    for i in range(a.steps):
        ev[i][0].record()
        step(a.warmup + i)
        ev[i][1].record()
    if world > 1:
        for w in gatherer.work:  # the last snapshots have landed (device-side wait, inside the timed region)
            if w is not None:
                w.wait()
        if gatherer.stream is not None:
            torch.cuda.current_stream().wait_stream(gatherer.stream)
        torch.cuda.current_stream().wait_stream(side)
    t_end.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    crossings = sum(pt.crossings for pt in batches) - c0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]

    budget = None
    if a.save_raw:
        # crossing-log post-processing of the Lagrangian budget on the log of the last step (K5)
        from seaduck_b200 import lagrangian_budget as lb

        pt = batches[(a.warmup + a.steps - 1) % nb]
        lb.find_ind_frac_tres_device(pt)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(3):
            out_b = lb.find_ind_frac_tres_device(pt)
        b1.record()
        torch.cuda.synchronize()
        n_rec = int(out_b[2].shape[0])
        bms = b0.elapsed_time(b1) / 3
        budget = {"records": n_rec, "ms": bms, "records_per_s": n_rec / (bms * 1e-3),
                  "what": "sdk_log_budget: wall indices, cell-traverse fractions and residence times of every log record"}

    # ---- the dominant kernel alone: CUDA events right around the C-ABI launch, fresh state each time
    kern_ms, kern_cross = [], []
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for i in range(min(nb, 8)):
        pt = batches[i]
        restore(i)
        p = pt._cparticles()
        par = pt._params(pt.max_iteration)
        stats.zero_()
        par.stats = stats.data_ptr()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.sdk_advect(handle, C.byref(pt._vel), C.byref(p), float(t_stop), C.byref(par), None, 1, _lib.current_stream()), "sdk_advect")
        e1.record()
        torch.cuda.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
        kern_cross.append(int(stats[1].item()))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    kms, kcr = float(np.mean(kern_ms)), float(np.mean(kern_cross))
    per_crossing = ALGO_BYTES_PER_CROSSING if a.levels else 72  # 2-D crossing = 4 walls + Vol + corners
    if a.dtype == "f32":
        per_crossing = 60 if a.levels else 52
    algo_bytes = kcr * per_crossing + n * STATE_BYTES_PER_PARTICLE
    achieved = algo_bytes / (kms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    e2e = None
    if not a.no_e2e and a.levels and not a.save_raw:
        pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()  # noqa: E731
        hin = [pin(x) for x in host_inputs] + [torch.zeros(n, dtype=torch.float64).pin_memory()]
        hout_f = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
        hout_i = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(6)]
        io = _lib.TrackIO()
        io.n = n
        io.lon, io.lat, io.dep, io.t = (h.data_ptr() for h in hin)
        io.out_lon, io.out_lat, io.out_dep = (h.data_ptr() for h in hout_f)
        io.out_face, io.out_iy, io.out_ix, io.out_izl, io.out_ncross, io.out_status = (h.data_ptr() for h in hout_i)
        scratch = torch.empty(int(L.sdk_track_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        par = batches[0]._params(200)
        par.stats = None
        in_bytes, out_bytes = 4 * 8 * n, 3 * 8 * n + 6 * 4 * n

        def run_e2e(steps):
            cross, secs = 0, 0.0
            for i in range(2 + steps):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _lib.check(L.sdk_track_host(handle, C.byref(batches[0]._vel), C.byref(io), None, 0, float(t_stop), C.byref(par),
                                            scratch.data_ptr(), _lib.current_stream()), "sdk_track_host")
                dt = time.perf_counter() - t0
                if i >= 2:
                    secs += dt
                    cross += int(hout_i[4].sum().item())
            return cross, secs

        if a.e2e_sweep:  # tuning aid: "chunks:blocks_per_sm,..." -> one line per setting on stderr
            for item in a.e2e_sweep.split(","):
                chunks, bps = (int(v) for v in item.split(":"))
                os.environ["SDK_TRACK_CHUNKS"] = str(chunks)
                par.blocks_per_sm = bps
                cross, secs = run_e2e(a.steps)
                print(f"e2e sweep chunks={chunks} blocks_per_sm={bps}: {cross / secs:.4e} crossings/s, {secs / a.steps * 1e3:.3f} ms/step", file=sys.stderr)
            os.environ.pop("SDK_TRACK_CHUNKS", None)
            par.blocks_per_sm = a.blocks_per_sm
        e2e_cross, e2e_t = run_e2e(a.steps)
        e2e = (e2e_cross, e2e_t, in_bytes, out_bytes)

    # ---- aggregate over ranks ------------------------------------------------------------------------------
    vals = torch.tensor([elapsed_ms, float(crossings), e2e[1] if e2e else 0.0, float(e2e[0]) if e2e else 0.0],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx, sm = vals.clone(), vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, crossings_all, e2e_time, e2e_cross_all = float(mx[0]), float(sm[1]), float(mx[2]), float(sm[3])
    else:
        crossings_all = float(crossings)
        e2e_time, e2e_cross_all = (e2e[1], float(e2e[0])) if e2e else (0.0, 0.0)
    if rank == 0:
        value = crossings_all / (elapsed_ms * 1e-3)
        line = {
            "metric": "particle cell-crossings/sec", "value": value, "unit": "crossings/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {
                "workload": workload_name(a), "transport": True, "particles_per_gpu_per_step": n, "days_per_step": a.days,
                "crossings_per_step": crossings_all / a.steps,
                "particle_order": "seeding" if a.no_sort else "cell-sorted at seeding time",
                "batches": f"{nb} distinct batches resident in HBM" + (", reused cyclically with a device-side state restore inside the timed region" if reuse else ", one per step"),
                "l2": "inputs larger than L2 (fields 168 MB + 220 MB of particle state per batch, a different batch every step); no flush",
                "collective": (f"all-gather of the output snapshot per step ({gatherer.mode}: "
                               + ("copy-engine writes into peer memory" if gatherer.mode == "peer" else "async NCCL all_gather_into_tensor on a side stream" + (f", {reserve} SMs left free by the advection kernel" if reserve else ""))
                               + ", overlapped with the next step"
                               + f", {parallel.SNAPSHOT_ROWS * 8 * n * world} B received per rank and step)") if world > 1 else "none",
                "step_ms_first10": [round(x, 3) for x in step_ms[:10]], "host_wall_s": wall,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "advect_kernel<%s,false>" % ("float" if a.dtype == "f32" else "double"),
                         "kernel_ms": kms, "crossings_per_launch": kcr, "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "instruction-supply bound, not bandwidth bound (DRAM 4.5 % busy, fields L2 resident, "
                                 "SM instruction cache hit rate 88 %, GPC instruction cache at 94 % of its request rate; the "
                                 "launch time follows the number of GPC instruction-cache requests): profiles/README.md r01k-r01m"},
            # this package's kernels inside the timed region, per rank and step: advect_kernel (two passes with the
            # crossing log: count, then fill) + finish_stop_kernel in lazy mode + pack_snapshot_kernel before the gather
            "gpu_launches": a.steps * world * ((2 if a.save_raw else 1) + (0 if (a.eager or a.save_raw) else 1)
                                               + (1 if (world > 1 and not a.no_gather) else 0)),
            "clocks": sampler.summary(),
        }
        if budget:
            line["config"]["budget_postprocessing"] = budget
        if e2e:
            line["e2e"] = {"value": e2e_cross_all / e2e_time if e2e_time else None, "unit": "crossings/s",
                           "h2d_bytes_per_step": e2e[2], "d2h_bytes_per_step": e2e[3],
                           "api": "sdk_track_host (C ABI): pinned host lon/lat/dep/t in, positions + indices out; "
                                  "locate + velocity read + advect on the device", "gpu_launches_per_step": 3}
        if not a.no_cpu_baseline:
            tot, el = cpu_baseline(a, threads=1)
            line["cpu_baseline"] = {"value": tot / el, "unit": "crossings/s", "cores": 1, "kind": "port",
                                    "sample": f"{a.cpu_sample} particles x {a.days:g} d, one step, C port of the reference loop (oracle/sd_oracle.c), 1 thread"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cuda(a)


if __name__ == "__main__":
    main()
