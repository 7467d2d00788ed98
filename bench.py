#!/usr/bin/env python
"""Headline benchmark: particle cell-crossings per second (BASELINE.json `metric`).

Workload (BASELINE.json configs[1]): a 3-D LLC90-shaped synthetic field (13 faces x 90 x 90 x 50
levels, volume transports in fp64) and 10^6 particles per GPU seeded uniformly on the sphere, so
trajectories cross cell walls, vertical levels and LLC face edges.  One *step* is one call of
``Particle.to_next_stop`` (one launch of the fused advection kernel) that carries every particle
`--days` days further; consecutive steps continue the same trajectories, so K steps are K output
stops of ``Particle.to_list_of_time``.  A crossing is an analytic step that ended on a cell wall
(``tend < 6``), counted on the device exactly like the instrumented reference does.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (see the keys at the bottom).  `value` = crossings/s with the state resident
in HBM; `e2e` = the same metric through the C-ABI host-buffer call (``sdk_advect_host``: pinned
host state -> H2D -> kernel -> D2H inside the timed region).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DAY = 86400.0
ALGO_BYTES_PER_CROSSING = 88  # 6 wall transports + Vol (fp64) + 4 corners x (XG,YG) f32
STATE_BYTES_PER_PARTICLE = 440  # 220 B of state read + 220 B written once per launch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--days", type=float, default=30.0, help="simulated days per step")
    ap.add_argument("--grid", type=int, default=90)
    ap.add_argument("--levels", type=int, default=50)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="field storage (math is fp64)")
    ap.add_argument("--no-sort", action="store_true", help="keep particles in seeding order")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="particles of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return f"LLC{a.grid}-shaped 3D synthetic transports ({13}x{a.grid}x{a.grid}x{a.levels}), {a.particles} particles/GPU, {a.days:g} d per step"


KW = dict(uname="utrans", vname="vtrans", wname="wtrans", transport=True)


def make_dataset(a):
    from seaduck_b200 import synthetic

    dtype = np.float64 if a.dtype == "f64" else np.float32
    return synthetic.llc(n=a.grid, nz=a.levels, dtype=dtype)


def seeds(ds, n, seed):
    from seaduck_b200 import synthetic

    return synthetic.llc_seed_particles(ds, n, seed=seed)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(s) > 3 + k and s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_baseline(a, threads, steps=1, seed=99):
    """Time the CPU oracle (C port of the reference loop) on a bounded sample of the same workload."""
    from oracle import oracle
    from seaduck_b200.ocedata import OceData

    ds = make_dataset(a)
    od = OceData(ds)
    n = a.cpu_sample
    lon, lat, dep = seeds(ds, n, seed)
    op = oracle.oracle_particle_from_od(od, lon, lat, dep, np.zeros(n), nthreads=threads, **KW)
    total, elapsed, per_step = 0, 0.0, []
    for k in range(steps):
        t0 = time.perf_counter()
        c = op.to_next_stop((k + 1) * a.days * DAY)
        dt = time.perf_counter() - t0
        total += c
        elapsed += dt
        per_step.append(dt)
    return total, elapsed, per_step


def run_reference(a):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the Python reference cannot
    travel to the GPU box) on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # warm-up steps advance the same trajectories, untimed
    from oracle import oracle
    from seaduck_b200.ocedata import OceData

    ds = make_dataset(a)
    od = OceData(ds)
    n = a.cpu_sample
    lon, lat, dep = seeds(ds, n, 99)
    op = oracle.oracle_particle_from_od(od, lon, lat, dep, np.zeros(n), nthreads=threads, **KW)
    k = 0
    for _ in range(a.warmup):
        k += 1
        op.to_next_stop(k * a.days * DAY)
    total = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        k += 1
        total += op.to_next_stop(k * a.days * DAY)
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    sample = f"{n} particles x {a.days:g} d per step (same field, same seeding law), C port of the reference loop, {threads} threads"
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
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ctypes as C

    import seaduck_b200 as sd
    from seaduck_b200 import _lib

    ds = make_dataset(a)
    od = sd.OceData(ds)
    n = a.particles
    lon, lat, dep = seeds(ds, n, seed=1000 + rank)  # disjoint particle slice per rank
    if not a.no_sort:
        # cell-major order so that the lanes of a warp touch neighbouring memory (K5)
        loc = od.locate(x=lon, y=lat, z=dep)
        key = ((loc["izl"].long() * 13 + loc["face"].long()) * a.grid + loc["iy"].long()) * a.grid + loc["ix"].long()
        order = torch.argsort(key).cpu().numpy()
        lon, lat, dep = lon[order], lat[order], dep[order]
    pt = sd.Particle(x=lon, y=lat, z=dep, t=np.zeros(n), data=od, **KW)
    pt.tuning = dict(refill_threshold=a.refill, blocks_per_sm=a.blocks_per_sm, threads_per_block=a.threads)
    snap_bytes = n * (3 * 8 + 4 * 4)
    gather = None
    if world > 1:
        gather = torch.empty(world * n * 5, dtype=torch.float64, device=dev)

    def snapshot_allgather():
        # output trajectories of all ranks (lon, lat, dep f64 + face/iy/ix/izl packed) -- the only
        # collective of the path
        st = pt._st
        packed = torch.cat([st["lon"], st["lat"], st["dep"],
                            (st["face"].double() * 1e6 + st["iy"].double()), (st["ix"].double() * 1e3 + st["izl"].double())])
        dist.all_gather_into_tensor(gather, packed)

    def step(k):
        pt.to_next_stop(k * a.days * DAY)
        if world > 1:
            snapshot_allgather()

    k = 0
    for _ in range(a.warmup):
        k += 1
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    c0 = pt.crossings
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    wall0 = time.perf_counter()
    for i in range(a.steps):
        k += 1
        ev[i][0].record()
        step(k)
        ev[i][1].record()
    t_end.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    elapsed_ms = t_start.elapsed_time(t_end)
    crossings = pt.crossings - c0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]

    # ---- kernel-only timing of the dominant kernel (events right around the C-ABI launch) ------
    L = _lib.lib()
    handle = od.grid_handle()
    kern_ms, kern_cross = [], []
    for i in range(min(a.steps, 5)):
        k += 1
        p = pt._cparticles()
        par = pt._params(pt.max_iteration)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.sdk_advect(handle, C.byref(pt._vel), C.byref(p), float(k * a.days * DAY), C.byref(par), None, 1, _lib.current_stream()), "sdk_advect")
        e1.record()
        torch.cuda.synchronize()
        pt._st["t"].fill_(k * a.days * DAY)
        kern_ms.append(e0.elapsed_time(e1))
        kern_cross.append(int(pt._st["ncross"].sum().item()))
    kms = float(np.mean(kern_ms))
    kcr = float(np.mean(kern_cross))
    algo_bytes = kcr * ALGO_BYTES_PER_CROSSING + n * STATE_BYTES_PER_PARTICLE
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

    # ---- end-to-end through the C-ABI host-buffer call -------------------------------------------
    e2e = None
    if not a.no_e2e:
        names = _lib.PARTICLE_F64 + _lib.PARTICLE_I32 + _lib.PARTICLE_F64B + _lib.PARTICLE_F32 + _lib.PARTICLE_F64C + ["status", "niter", "ncross"]
        host = {nm: (None if pt._st.get(nm) is None else pt._st[nm].cpu().pin_memory()) for nm in names}
        hp = _lib.Particles()
        hp.n = n
        for nm in names:
            setattr(hp, nm, None if host[nm] is None else host[nm].data_ptr())
        scratch = torch.empty(int(L.sdk_particles_bytes(n)), dtype=torch.uint8, device=dev)
        in_bytes = sum(v.numel() * v.element_size() for nm, v in host.items() if v is not None and nm not in ("status", "niter", "ncross"))
        out_bytes = sum(v.numel() * v.element_size() for nm, v in host.items() if v is not None)
        par = pt._params(pt.max_iteration)
        kk = k
        e2e_cross, e2e_t = 0, 0.0
        for i in range(2 + a.steps):
            kk += 1
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(L.sdk_advect_host(handle, C.byref(pt._vel), C.byref(hp), float(kk * a.days * DAY), C.byref(par), scratch.data_ptr(), _lib.current_stream()), "sdk_advect_host")
            dt = time.perf_counter() - t0
            host["t"].fill_(kk * a.days * DAY)
            if i >= 2:
                e2e_t += dt
                e2e_cross += int(host["ncross"].sum().item())
        e2e = (e2e_cross, e2e_t, in_bytes, out_bytes)

    # ---- aggregate over ranks ------------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, float(crossings), e2e[1] if e2e else 0.0, float(e2e[0]) if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, crossings_all = float(mx[0]), float(sm[1])
        e2e_time, e2e_cross_all = float(mx[2]), float(sm[3])
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
                "workload": workload_name(a), "transport": True, "particles_per_gpu": n, "days_per_step": a.days,
                "crossings_per_step": crossings_all / a.steps, "particle_order": "seeding" if a.no_sort else "cell-sorted at start",
                "l2": "inputs larger than L2 (fields 168 MB + particle state 220 MB per 1e6 particles); no flush",
                "collective": "all_gather of the output snapshot per step" if world > 1 else "none",
                "step_ms": [round(x, 3) for x in step_ms], "host_wall_s": wall,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "advect_kernel",
                         "kernel_ms": kms, "crossings_per_launch": kcr,
                         "algorithmic_bytes_per_launch": algo_bytes},
            "gpu_launches": a.steps * world,
            "clocks": sampler.summary(),
        }
        if e2e:
            line["e2e"] = {"value": e2e_cross_all / e2e_time if e2e_time else None, "unit": "crossings/s",
                           "h2d_bytes_per_step": e2e[2], "d2h_bytes_per_step": e2e[3],
                           "api": "sdk_advect_host (C ABI, pinned host state in/out)"}
        if not a.no_cpu_baseline:
            tot, el, _ = cpu_baseline(a, threads=1)
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
