#!/usr/bin/env python
"""Headline benchmark: particle cell-crossings per second (BASELINE.json `metric`).

Headline workload (BASELINE.json configs[1]): a 3-D LLC90-shaped synthetic field (13 faces x 90 x 90 x 50
levels, volume transports, fp64) and batches of 10^6 particles per GPU seeded uniformly on the sphere IN
SEEDING ORDER (scattered; the product puts them into cell order itself, ``Particle(sort="auto")``), so
trajectories cross cell walls, vertical levels and LLC face edges.

One *step* = one pass of the hot path over one batch: ``Particle.to_next_stop`` (one launch of the fused
advection kernel) carries a fresh batch of 10^6 located particles `--days` days forward.  Every step works on
a different batch (all batches are resident in HBM before the timed region; fields 168 MB + 220 MB of state
per batch >> L2, so nothing is served from a warm cache).  A crossing is an analytic step that ended on a
cell wall (``tend < 6``), counted on the device exactly like the instrumented reference does (wrap
``cross_cell_wall``, sum ``tend < 6``).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--extras auto|none|c4,c5,f32,c3]

Prints ONE JSON line.  `value` = crossings/s with the state resident in HBM; `e2e` = the same metric through
the C-ABI host-buffer call ``sdk_track_host`` (pinned host lon/lat/dep/t in seeding order -> H2D -> locate ->
cell sort -> velocity read -> advect -> 32-byte records -> D2H, all inside the timed region).
`extra_workloads` carries the other configurations of BASELINE.json measured by the same process right after
the headline: configs[3] (LLC4320 surface, time snapshots, 1.25e7 particles per GPU, 3-hour stops, through
``to_list_of_time``), configs[4] (crossing log), fp32 field storage, and configs[2] (Eulerian sampling) --
each with its own clock record.  With N > 1 ranks the output records of every step go to all ranks through
``sdk_pack_ship`` (stores into peer memory from inside the packing kernel).
"""

import argparse
import gc
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DAY = 86400.0
# SURVEY.md §8d, algorithmic bytes: a 3-D crossing reads 6 wall transports + Vol (fp64) + 4 corners x (XG, YG) f32;
# a 2-D one 4 walls + Vol + corners; the state costs ~128 B read + 128 B written per particle and launch
BYTES_PER_CROSSING = {("f64", True): 88, ("f64", False): 72, ("f32", True): 60, ("f32", False): 52}
STATE_BYTES_8D = 256
STATE_BYTES_LAYOUT = 440  # what the 220-byte device layout really moves per particle and launch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU and batch")
    ap.add_argument("--days", type=float, default=30.0, help="simulated days per step")
    ap.add_argument("--grid", type=int, default=90)
    ap.add_argument("--levels", type=int, default=50, help="vertical levels; 0 = surface-only 2-D run")
    ap.add_argument("--snapshots", type=int, default=0, help="time snapshots of the velocity field (0 = steady)")
    ap.add_argument("--save-raw", action="store_true", help="write the reference's crossing log (configs[4]): count pass + fill pass per step")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="field storage (math is fp64)")
    ap.add_argument("--sort", default="auto", choices=["auto", "on", "off"], help="Particle(sort=...): cell order kept by the product")
    ap.add_argument("--max-batches", type=int, default=24, help="distinct batches kept in HBM (reused cyclically)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="particles of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extras", default="auto", help="auto (all at 1 GPU, c4 at N > 1), none, or a list of c4,c5,f32,c3")
    ap.add_argument("--c4-grid", type=int, default=4320)
    ap.add_argument("--c4-particles", type=int, default=12_500_000)
    ap.add_argument("--c4-hours", type=float, default=3.0)
    ap.add_argument("--c3-points", type=int, default=10_000_000)
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--eager", action="store_true", help="reference-style to_next_stop (host reads the call statistics every step)")
    ap.add_argument("--no-gather", action="store_true", help="(diagnostic) skip the record exchange at N > 1")
    ap.add_argument("--no-multicast", action="store_true", help="record exchange with plain peer stores even where NVLS is available")
    return ap.parse_args()


def workload_name(a):
    shape = f"3D synthetic transports (13x{a.grid}x{a.grid}x{a.levels})" if a.levels else f"surface-only 2D synthetic transports (13x{a.grid}x{a.grid})"
    extra = (f", {a.snapshots} time snapshots" if a.snapshots else "") + (", crossing log written (save_raw)" if a.save_raw else "")
    return f"LLC{a.grid}-shaped {shape}{extra}, {a.particles} particles per GPU and step, {a.days:g} d per step"


def kwargs_of(a):
    return dict(uname="utrans", vname="vtrans", wname="wtrans" if a.levels else None, transport=True)


def make_dataset(a):
    from seaduck_b200 import synthetic

    return synthetic.llc(n=a.grid, nz=a.levels, nt=a.snapshots, dtype=np.float64 if a.dtype == "f64" else np.float32)


def seeds(n, seed):
    from seaduck_b200 import synthetic

    return synthetic.llc_seed_particles(None, n, seed=seed)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

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


# ---------------------------------------------------------------------------------------------------------------
# the reference arm and the CPU baseline: the oracle (C port of the reference loop), nothing of the product on the timed path
# ---------------------------------------------------------------------------------------------------------------
def oracle_particle(a, n, threads, seed=99):
    """Setup only: the dataset comes from seaduck_b200.synthetic and the float32 grid extraction from the product's
    OceData (host code, no GPU); what is timed afterwards is `oracle.OracleParticle.to_next_stop` alone."""
    from oracle import oracle
    from seaduck_b200.ocedata import OceData

    ds = make_dataset(a)
    od = OceData(ds)
    lon, lat, dep = seeds(n, seed)
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
              f"(oracle/sd_oracle.c), {threads} threads; only OracleParticle.to_next_stop is timed -- dataset and grid "
              f"extraction (seaduck_b200.synthetic / OceData host code) are setup outside the timed region")
    line = {
        "impl": "reference", "metric": "particle cell-crossings/sec", "value": value, "unit": "crossings/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * elapsed / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "crossings/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "crossings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the CUDA arm: device, ranks, the record exchange."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        self.a = a
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.affinity = self._bind_to_gpu_cpus(self.local) if self.world > 1 else "unchanged (one rank)"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.torch, self.dist = torch, dist
        self.side = torch.cuda.Stream(self.dev)

    @staticmethod
    def _bind_to_gpu_cpus(index):
        """Run this rank on the CPUs next to its GPU (NVML's ideal affinity), so that its pinned buffers are first touched
        -- and therefore allocated -- on the GPU's own NUMA node.  Best effort: a container that does not own those
        CPUs keeps what it has."""
        try:
            import pynvml

            pynvml.nvmlInit()
            before = sorted(os.sched_getaffinity(0))
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
            after = sorted(os.sched_getaffinity(0))
            return f"cpus {after[0]}-{after[-1]} ({len(after)}), was {before[0]}-{before[-1]} ({len(before)})"
        except Exception as exc:  # noqa: BLE001
            return f"unchanged ({type(exc).__name__})"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, vals):
        """(max over ranks, sum over ranks) of a list of floats."""
        torch = self.torch
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world == 1:
            return list(vals), list(vals)
        mx, sm = t.clone(), t.clone()
        self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(sm, op=self.dist.ReduceOp.SUM)
        return mx.tolist(), sm.tolist()


def nvlink_bytes(index):
    """(tx, rx) data bytes this GPU has moved over all its NVLinks so far (NVML throughput counters, KiB), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        every_link = 0xFFFFFFFF
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, every_link),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, every_link)])
        if any(v.nvmlReturn != 0 for v in vals):
            return None
        return tuple(int(v.value.ullVal) * 1024 for v in vals)
    except Exception:  # noqa: BLE001
        return None


def roofline_of(kernel, kms, crossings, n, dtype, three_d, traffic=None, traffic_source=None):
    """Roofline of the advection kernel from a live CUDA-event timing: algorithmic bytes per SURVEY.md §8d."""
    peak, peak_src = hbm_peak()
    per_crossing = BYTES_PER_CROSSING[(dtype, three_d)]
    algo = crossings * per_crossing + n * STATE_BYTES_8D
    achieved = algo / (kms * 1e-3) / 1e9
    layout = (crossings * per_crossing + n * STATE_BYTES_LAYOUT) / (kms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_source, "peak_source": peak_src, "kernel": kernel, "kernel_ms": kms,
            "crossings_per_launch": crossings, "algorithmic_bytes_per_launch": algo,
            "bytes_model": f"{per_crossing} B per crossing + {STATE_BYTES_8D} B per particle and launch (SURVEY.md 8d)",
            "frac_with_layout_state_bytes": layout / peak,
            "layout_state_bytes": f"{STATE_BYTES_LAYOUT} B per particle and launch: the 220-byte device state read and written once"}


def load_traffic(kind):
    """DRAM bytes per launch of the advection kernel from the last ncu --set full capture of this build (profiles/)."""
    path = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if not os.path.exists(path):
        return None, None
    rec = json.load(open(path))
    if kind == "c4":
        rec = rec.get("configs[3]")
    elif kind != "c2":
        rec = None
    if not rec:
        return None, None
    return rec.get("dram_bytes_per_launch"), f"profiles/traffic_bytes_per_launch.json ({rec.get('source', 'ncu --set full')})"


def run_batches(ctx, a, kind="c2", with_e2e=False, with_orders=False):
    """Batch workload (configs[1] and its variants): every step carries a fresh resident batch `days` days forward."""
    import ctypes as C

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, parallel

    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    L = _lib.lib()
    ds = make_dataset(a)
    od = sd.OceData(ds)
    handle = od.grid_handle()
    n = a.particles
    t_stop = a.days * DAY
    nb = min(a.warmup + a.steps, a.max_batches)
    sort = {"auto": "auto", "on": True, "off": False}[a.sort]

    def make_batch(b, sort_flag):
        lon, lat, dep = seeds(n, seed=1000 + 97 * rank + b)  # disjoint particle slices per rank / batch, seeding order
        pt = sd.Particle(x=lon, y=lat, z=dep if a.levels else np.full(n, -1.0), t=np.zeros(n), data=od, save_raw=a.save_raw,
                         sort=sort_flag, **kwargs_of(a))
        pt.lazy = not a.eager  # device-side end-of-stop bookkeeping: the host never waits inside a step
        pt.tuning = dict(refill_threshold=a.refill, blocks_per_sm=a.blocks_per_sm, threads_per_block=a.threads)
        return pt, (lon, lat, dep)

    fused = os.environ.get("SDK_BENCH_EXCH", "fused") == "fused"  # (experiments: "kernel" = separate sdk_pack_ship launch)

    batches, host_inputs = [], None
    for b in range(nb):
        pt, xyz = make_batch(b, sort)
        batches.append(pt)
        if b == 0:
            host_inputs = xyz
    sorted_by_product = batches[0]._perm is not None
    reuse = a.warmup + a.steps > nb
    saved = [{k: (None if v is None else v.clone()) for k, v in pt._st.items()} for pt in batches]
    exchange = None
    if world > 1 and not a.no_gather:
        exchange = parallel.RecordExchange(n, dev, slots=int(os.environ.get("SDK_BENCH_SLOTS", "2")), multicast=not a.no_multicast)
    if exchange is not None and fused and not a.save_raw:
        # the advection kernel stores the output records into every rank's buffer itself, in the cell order of its state
        # (contiguous stores); the permutation that says which record is which particle is exchanged once per batch,
        # at set-up time, like ShardedParticle does at construction
        order = os.environ.get("SDK_BENCH_SHIP_ORDER", "state")
        perms = []
        for pt in batches:
            pt.ship_to(exchange, order=order)
            if order == "state" and pt._perm is not None:
                everyone = torch.empty((world, n), dtype=torch.int32, device=dev)
                ctx.dist.all_gather_into_tensor(everyone, pt._perm)
                perms.append(everyone)

    def restore(group, b):
        for k, v in group[1][b].items():
            if v is not None:
                group[0][b]._st[k].copy_(v)

    def step(group, i):
        b = i % nb
        pt = group[0][b]
        if i >= nb:
            restore(group, b)
        if a.save_raw:
            pt.empty_lists()  # the log of this step only (the reference's to_list_of_time does the same)
        pt.to_next_stop(t_stop)
        if exchange is not None:
            # output records of all ranks -- the only exchange of the path.  Fused: the advection kernel has
            # stored them into every rank's receive buffer while it ran; what is left is a one-warp kernel on a side
            # stream that waits for the other ranks' blocks of this step and frees the slot.
            done = torch.cuda.Event()
            done.record()
            ctx.side.wait_event(done)
            with torch.cuda.stream(ctx.side):
                s = pt.last_ship_step if pt.__dict__.get("_exchange") is not None else exchange.ship(pt._cparticles(), out_index=pt._out_index())
                exchange.wait(s, release=True)  # everybody's records of this step are here; the slot is free again

    def timed(group):
        for i in range(a.warmup):
            step(group, i)
        torch.cuda.synchronize()
        ctx.barrier()
        c0 = sum(pt.crossings for pt in group[0])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        t_start.record()
        for i in range(a.steps):
            ev[i][0].record()
            step(group, a.warmup + i)
            ev[i][1].record()
        if exchange is not None:
            torch.cuda.current_stream().wait_stream(ctx.side)  # the last records have landed (inside the timed region)
        t_end.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        ctx.barrier()
        crossings = sum(pt.crossings for pt in group[0]) - c0
        return t_start.elapsed_time(t_end), crossings, [e0.elapsed_time(e1) for e0, e1 in ev], wall

    sampler = ClockSampler(ctx.local)
    sampler.start()
    elapsed_ms, crossings, step_ms, wall = timed((batches, saved))
    nvlink = {}
    if exchange is not None:
        exchange.check()
        # NVLink data bytes of this GPU per step: NVML throughput counters around a few more (untimed) steps -- reading
        # them next to the timed region stalls a step by ~10 ms
        link0 = nvlink_bytes(ctx.local)
        extra = 5
        for i in range(extra):
            step((batches, saved), a.warmup + a.steps + i)
        torch.cuda.current_stream().wait_stream(ctx.side)
        torch.cuda.synchronize()
        ctx.barrier()
        link1 = nvlink_bytes(ctx.local)
        if link0 is not None and link1 is not None:
            nvlink = {"tx_bytes_per_step": (link1[0] - link0[0]) / extra, "rx_bytes_per_step": (link1[1] - link0[1]) / extra}

    # ---- the dominant kernel alone: CUDA events right around the C-ABI launch, fresh state each time
    kern_ms, kern_cross = [], []
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for i in range(min(nb, 8)):
        pt = batches[i]
        restore((batches, saved), i)
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
    clocks = sampler.finish()
    kms, kcr = float(np.mean(kern_ms)), float(np.mean(kern_cross))
    traffic, traffic_src = load_traffic(kind)
    kname = "advect_kernel<%s,%s,%d>" % ("float" if a.dtype == "f32" else "double", "true" if a.save_raw else "false", 1 if a.levels else 0)
    roof = roofline_of(kname, kms, kcr, n, a.dtype, bool(a.levels), traffic, traffic_src)

    mx, sm = ctx.reduce([elapsed_ms, float(crossings)])
    out = {
        "value": sm[1] / (mx[0] * 1e-3), "ms_per_step": mx[0] / a.steps, "crossings_per_step": sm[1] / a.steps,
        "roofline": roof, "clocks": clocks, "step_ms_first10": [round(x, 3) for x in step_ms[:10]], "host_wall_s": wall,
        "batches": f"{nb} distinct batches resident in HBM" + (", reused cyclically with a device-side state restore inside the timed region" if reuse else ", one per step"),
        "particle_order": "seeding order in; kept in cell order on the device by the product (Particle(sort=...), sdk_cell_sort)" if sorted_by_product else "seeding order (no sort)",
        "collective": ((f"fused into the advection kernel (sdk_advect_params.ship): the write-back of every finished particle also stores its 32-byte record into "
                        f"every rank's receive buffer ({exchange.mode}), in the cell order of the sender's state (its permutation was "
                        f"exchanged once, at set-up)" if (fused and not a.save_raw) else
                        f"sdk_pack_ship per step on a side stream ({exchange.mode})")
                       + f"; one-warp arrival wait + slot release per step on a side stream; {32 * n * world} B received per rank and step") if exchange is not None else "none",
        "nvlink": dict(nvlink, source="NVML NVLINK_THROUGHPUT_DATA_TX / RX of rank 0's GPU, all links, around 5 untimed steps after the timed region",
                       expected_rx_bytes_per_step=32 * n * (world - 1)) if nvlink else None,
        # this package's kernels inside the timed region, per rank and step: advect_kernel (two passes with the crossing
        # log: count, then fill) + finish_stop_kernel in lazy mode + exchange_wait_kernel at N > 1 (+ pack_ship_kernel when not fused)
        "gpu_launches": a.steps * world * ((2 if a.save_raw else 1) + (0 if (a.eager or a.save_raw) else 1)
                                           + ((1 if (fused and not a.save_raw) else 2) if exchange is not None else 0)),
    }

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
        out["budget_postprocessing"] = {"records": n_rec, "ms": bms, "records_per_s": n_rec / (bms * 1e-3),
                                        "what": "sdk_log_budget: wall indices, cell-traverse fractions and residence times of every log record"}

    if with_orders and not a.save_raw:
        # the same steps with the product's sort switched the other way: what the cell order is worth
        other_sort = not sorted_by_product
        ob = [make_batch(b, other_sort)[0] for b in range(nb)]
        osaved = [{k: (None if v is None else v.clone()) for k, v in pt._st.items()} for pt in ob]
        ex, exchange = exchange, None  # (single-rank comparison of the kernels; no exchange)
        o_ms, o_cross, _, _ = timed((ob, osaved))
        exchange = ex
        mx2, sm2 = ctx.reduce([o_ms, float(o_cross)])
        out["orders"] = {("cell order kept by the product" if other_sort else "seeding order (sort=False)"): sm2[1] / (mx2[0] * 1e-3),
                         ("cell order kept by the product" if sorted_by_product else "seeding order (sort=False)"): (sm[1] / (mx[0] * 1e-3)) if exchange is None else None,
                         "note": "resident crossings/s, same seeds, no record exchange in the comparison run"}
        del ob, osaved

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    if with_e2e and a.levels and not a.save_raw:
        pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()  # noqa: E731
        hin = [pin(x) for x in host_inputs] + [torch.zeros(n, dtype=torch.float64).pin_memory()]
        rec = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
        hstats = torch.zeros(4, dtype=torch.int64).pin_memory()
        scratch = torch.empty(int(L.sdk_track_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        par = batches[0]._params(200)
        par.stats = None
        in_bytes, out_bytes = 4 * 8 * n, 32 * n + 32

        def run_e2e(inputs, sort_flag, steps):
            io = _lib.TrackIO()
            io.n, io.sort = n, int(sort_flag)
            io.lon, io.lat, io.dep, io.t = (h.data_ptr() for h in inputs)
            io.out_records, io.out_stats = rec.data_ptr(), hstats.data_ptr()
            cross, secs = 0, 0.0
            for i in range(2 + steps):
                ctx.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _lib.check(L.sdk_track_host(handle, C.byref(batches[0]._vel), C.byref(io), None, 0, float(t_stop), C.byref(par),
                                            scratch.data_ptr(), _lib.current_stream()), "sdk_track_host")
                dt = time.perf_counter() - t0
                if i >= 2:
                    secs += dt
                    cross += int(hstats[1].item())
            return cross, secs

        e_cross, e_secs = run_e2e(hin, False, a.steps)
        mxe, sme = ctx.reduce([e_secs, float(e_cross)])
        out["e2e"] = {"value": sme[1] / mxe[0], "unit": "crossings/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                      "api": "sdk_track_host (C ABI): pinned host lon/lat/dep/t in seeding order in, one 32-byte record per particle "
                             "(position + cell, caller's order) and 4 counters out; locate + velocity read + advect on the device, the advection "
                             "kernel stores the records into the caller's pinned buffer itself (no pack pass, no D2H copy); 3 chunks pipelined "
                             "over three streams", "cpu_affinity_rank0": ctx.affinity, "gpu_launches_per_step": 9,
                      "gpu_launches_note": "3 chunks x (locate, get_u_du, advect)"}
        if with_orders:
            # the same call without the device-side sort, and with host input that already is in cell order
            u_cross, u_secs = run_e2e(hin, True, max(3, a.steps // 2))
            order = batches[0]._perm.cpu().numpy() if batches[0]._perm is not None else np.arange(n)
            hsorted = [pin(x[order]) for x in host_inputs] + [hin[3]]
            s_cross, s_secs = run_e2e(hsorted, False, max(3, a.steps // 2))
            out["e2e"]["orders"] = {"seeding order in, no sort (the headline e2e)": e_cross / e_secs,
                                    "seeding order in, (level, bucket) sort on the device before locating (sdk_track_io.sort)": u_cross / u_secs,
                                    "cell-sorted host input, no sort": s_cross / s_secs, "note": "this rank only"}
    return out


def run_stops(ctx, a, grid, particles, hours, nt=3):
    """configs[3]: LLC4320-shaped surface field with time snapshots, the SAME particles carried from stop to stop through
    ``to_list_of_time`` (3-hour stops as in the reference's LLC4320 notebook, docs/sciserver_notebooks/LLC4320.md:119-121),
    one 32-byte record per particle and stop kept in HBM (``snapshots="records"``).  The run starts 0.5 d before
    time_midp[0], so the velocity slab is swapped once inside the timed region: two of the three snapshots are used."""
    import ctypes as C

    import psutil

    import seaduck_b200 as sd
    from seaduck_b200 import _lib, parallel, synthetic

    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    need_host = 12 * 13 * grid * grid * 4 * world * 1.5
    if psutil.virtual_memory().available < need_host:
        return {"skipped": f"host memory: {psutil.virtual_memory().available / 1e9:.0f} GB available, {need_host / 1e9:.0f} GB wanted for the float32 grid of {world} ranks"}
    t0 = time.perf_counter()
    ds, fields = synthetic.llc_surface_device(n=grid, nt=nt, device=dev)
    od = sd.OceData(ds)
    for name, (dims, tensor) in fields.items():
        od.stage_device_field(name, tensor, dims)
    del fields
    build_s = time.perf_counter() - t0
    n = particles
    lon, lat, _ = seeds(n, seed=4000 + rank)
    t_first = 4.45 * DAY  # time_midp[0] = 5 d, not itself one of the stops
    dt = hours * 3600.0
    kw = dict(uname="utrans", vname="vtrans", wname=None, transport=True)
    pt = sd.Particle(x=lon, y=lat, z=None, t=np.full(n, t_first), data=od, **kw)
    pt.lazy = True
    pt.tuning = dict(refill_threshold=a.refill, blocks_per_sm=a.blocks_per_sm, threads_per_block=a.threads)
    setup_s = time.perf_counter() - t0
    exchange = None

    def run(first, count, gather):
        """`count` stops through to_list_of_time; returns (ms, crossings, number of snapshots)."""
        nonlocal exchange
        stops = t_first + dt * (first + 1 + np.arange(count))
        ctx.barrier()
        c0 = pt.crossings
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        if gather == "owner":
            _, snaps = pt.to_list_of_time(stops, snapshots="records")
        else:
            if exchange is None:
                exchange = parallel.RecordExchange(n, dev, slots=2, multicast=not a.no_multicast)
            pt.ship_to(exchange)
            snaps = []

            def on_stop(t_stop):
                done = torch.cuda.Event()
                done.record()
                ctx.side.wait_event(done)
                with torch.cuda.stream(ctx.side):
                    exchange.wait(pt.last_ship_step, release=True)
                snaps.append(pt.last_ship_step)

            pt.to_list_of_time(stops, _on_stop=on_stop)
            torch.cuda.current_stream().wait_stream(ctx.side)
            pt.ship_to(None)
        e1.record()
        torch.cuda.synchronize()
        ctx.barrier()
        return e0.elapsed_time(e1), pt.crossings - c0, len(snaps)

    W, K = a.warmup, a.steps
    pt.reserve_records(W + K + 4)  # the records of every stop of this run, allocated once, outside the timed region
    run(0, W, "owner")
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ms, cross, nsnap = run(W, K, "owner")
    mx, sm = ctx.reduce([ms, float(cross)])
    out = {
        "config": {
            "workload": f"LLC{grid}-shaped surface-only 2D synthetic transports (13x{grid}x{grid}) with {nt} time snapshots (2 used: the run "
                        f"crosses time_midp[0]), {n} particles per GPU carried from stop to stop, {hours:g} h per stop, through "
                        f"Particle.to_list_of_time(snapshots='records')",
            "crossings_per_step": sm[1] / K, "crossings_per_particle_and_stop": sm[1] / K / (n * world),
            "records": f"{nsnap} record snapshots of 32 B x {n} kept in HBM per rank (every rank keeps its own slice; assembled on demand)",
            "l2": f"inputs larger than L2 (fields {2 * 13 * grid * grid * 8 / 1e6:.0f} MB per snapshot, particle state {220 * n / 1e6:.0f} MB); no flush",
            "dataset_build_s": build_s, "setup_s": setup_s,
        },
        "value": sm[1] / (mx[0] * 1e-3), "unit": "crossings/s", "ms_per_step": mx[0] / K, "n_gpus": world, "steps": K, "warmup": W,
        "scaling": "weak", "collective": "none on the data path (gather='owner': every rank keeps the records of its slice)",
        # per rank and stop: advect_kernel + finish_stop_kernel + pack_records_kernel (+ get_u_du_kernel at the slab swap)
        "gpu_launches": world * (3 * nsnap + 1),
    }
    if world > 1 and not a.no_gather:
        ms2, cross2, _ = run(W + K, K, "all")
        mx2, sm2 = ctx.reduce([ms2, float(cross2)])
        exchange.check()
        out["all_gather_variant"] = {"value": sm2[1] / (mx2[0] * 1e-3), "ms_per_step": mx2[0] / K,
                                     "collective": f"fused into the advection kernel of every stop ({exchange.mode}): {32 * n * world} B received per rank and stop"}
    # the kernel alone, on a copy of the state, one more stop ahead
    L = _lib.lib()
    keep = {k: (None if v is None else v.clone()) for k, v in pt._st.items()}
    t_now = float(pt._st["t"][0].item())
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    kern_ms, kern_cross = [], []
    for _ in range(3):
        for k, v in keep.items():
            if v is not None:
                pt._st[k].copy_(v)
        p = pt._cparticles()
        par = pt._params(pt.max_iteration)
        stats.zero_()
        par.stats = stats.data_ptr()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.sdk_advect(od.grid_handle(), C.byref(pt._vel), C.byref(p), t_now + dt, C.byref(par), None, 1, _lib.current_stream()), "sdk_advect")
        e1.record()
        torch.cuda.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
        kern_cross.append(int(stats[1].item()))
    out["clocks"] = sampler.finish()
    traffic, traffic_src = load_traffic("c4") if (grid, particles) == (4320, 12_500_000) else (None, None)
    out["roofline"] = roofline_of("advect_kernel<double,false,3>", float(np.mean(kern_ms)), float(np.mean(kern_cross)), n, "f64", False, traffic, traffic_src)
    return out


def run_interp(ctx, a):
    """configs[2]: the Eulerian sampling benchmark of bench_interp.py, in this process."""
    import bench_interp

    return bench_interp.measure(points=a.c3_points, steps=max(3, a.steps // 2), warmup=a.warmup, cpu_baseline=False,
                                clock_sampler=ClockSampler(ctx.local))


def variant(a, **kw):
    b = argparse.Namespace(**vars(a))
    for k, v in kw.items():
        setattr(b, k, v)
    return b


def run_cuda(a):
    ctx = Ctx(a)
    world, rank = ctx.world, ctx.rank
    main = run_batches(ctx, a, kind="c2", with_e2e=not a.no_e2e, with_orders=(world == 1))
    extras = []
    if a.extras == "auto":
        which = ["c4", "c5", "f32", "c3"] if world == 1 else ["c4"]
    elif a.extras == "none":
        which = []
    else:
        which = [w for w in a.extras.split(",") if w]
    for w in which:
        gc.collect()  # (the objects of the previous workload hold GBs of device memory through reference cycles)
        ctx.torch.cuda.empty_cache()
        try:
            if w == "c4":
                r = run_stops(ctx, a, a.c4_grid, a.c4_particles, a.c4_hours)
                r["name"] = "configs[3]"
            elif w == "c5":
                b = variant(a, save_raw=True, max_batches=4, steps=max(3, a.steps // 4), sort="off")
                r = run_batches(ctx, b, kind="c5")
                r = {"name": "configs[4]", "config": {"workload": workload_name(b), "batches": r.pop("batches"), "particle_order": r.pop("particle_order"),
                                                       "budget_postprocessing": r.pop("budget_postprocessing", None)},
                     "unit": "crossings/s", "n_gpus": world, "steps": b.steps, "warmup": b.warmup, **r}
            elif w == "f32":
                b = variant(a, dtype="f32")
                r = run_batches(ctx, b, kind="f32")
                r = {"name": "configs[1] with fp32 field storage (fp64 math)", "config": {"workload": workload_name(b), "batches": r.pop("batches"), "particle_order": r.pop("particle_order")},
                     "unit": "crossings/s", "dtype": "f32", "n_gpus": world, "steps": b.steps, "warmup": b.warmup, **r}
            elif w == "c3":
                r = run_interp(ctx, a)
                r["name"] = "configs[2]"
            else:
                continue
        except Exception as exc:  # noqa: BLE001  (an extra workload must not take the headline line with it)
            r = {"name": w, "failed": repr(exc)[:400]}
        extras.append(r)
    if rank == 0:
        line = {
            "metric": "particle cell-crossings/sec", "value": main["value"], "unit": "crossings/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {
                "workload": workload_name(a), "transport": True, "particles_per_gpu_per_step": a.particles, "days_per_step": a.days,
                "crossings_per_step": main["crossings_per_step"], "particle_order": main["particle_order"], "batches": main["batches"],
                "l2": "inputs larger than L2 (fields 168 MB + 220 MB of particle state per batch, a different batch every step); no flush",
                "collective": main["collective"], "nvlink": main.get("nvlink"), "step_ms_first10": main["step_ms_first10"],
                "host_wall_s": main["host_wall_s"],
            },
            "roofline": main["roofline"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
        }
        for k in ("orders", "budget_postprocessing"):
            if k in main:
                line["config"][k] = main[k]
        if "e2e" in main:
            line["e2e"] = main["e2e"]
        if extras:
            line["extra_workloads"] = extras
        if not a.no_cpu_baseline:
            tot, el = cpu_baseline(a, threads=1)
            line["cpu_baseline"] = {"value": tot / el, "unit": "crossings/s", "cores": 1, "kind": "port",
                                    "sample": f"{a.cpu_sample} particles x {a.days:g} d, one step, C port of the reference loop (oracle/sd_oracle.c), 1 thread"}
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cuda(a)


if __name__ == "__main__":
    main()
