#!/usr/bin/env python
"""Eulerian sampling benchmark (BASELINE.json configs[2]): T, S and (u, v) at 10^7 scattered points on an
LLC270-shaped grid (13 x 270 x 270 x 50) with masked 9-point kernels for the tracers and the
particle kernels for the velocity -- ``Position.from_latlon`` + ``Position.interpolate``.

    python bench_interp.py [--points N] [--grid n] [--levels nz] [--steps K] [--warmup W]

One *step* = one pass over all points: three launches of the interpolation kernel (T, S, (u, v)).
Prints ONE JSON line in the layout of bench.py: `value` = points/s with the located points resident in
HBM, `e2e` = host lon/lat/dep in -> locate -> interpolate -> host arrays out, `roofline` from CUDA
events around the launches and the algorithmic bytes of SURVEY.md 8d (298 B/point), `cpu_baseline` =
the numpy/C oracle (oracle/interp_oracle.py) on a bounded sample.  The companion of bench.py (which
stays the driver's headline benchmark); not run by the driver.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# T, S with the masked 9-point kernel (2 x 81 B), (u, v) with the default kernels (48 B),
# inputs 24 B, corners 32 B, four outputs 32 B
ALGO_BYTES_PER_POINT = 298


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--grid", type=int, default=270)
    ap.add_argument("--levels", type=int, default=50)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=20_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sorted", action="store_true", help="cell-sort the points first (default: scattered, as configs[2] says)")
    return ap.parse_args()


def measure(points=10_000_000, grid=270, levels=50, steps=10, warmup=3, cpu_sample=20_000, cpu_baseline=True, e2e=True,
            presorted=False, clock_sampler=None):
    """Run the benchmark in this process and return the JSON line as a dict (bench.py calls this for its
    `extra_workloads`; `clock_sampler` is a started-on-demand NVML sampler thread of the caller)."""
    a = argparse.Namespace(points=points, grid=grid, levels=levels, steps=steps, warmup=warmup, cpu_sample=cpu_sample,
                           no_cpu_baseline=not cpu_baseline, no_e2e=not e2e, sorted=presorted)
    import torch

    import seaduck_b200 as sd
    from seaduck_b200 import interp, synthetic

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases_interp as ci

    t0 = time.perf_counter()
    ds = synthetic.llc(n=a.grid, nz=a.levels, land_blobs=True, tracers=True)
    od = sd.OceData(ds)
    n = a.points
    lon, lat, _ = synthetic.llc_seed_particles(ds, n, seed=3, lat_range=(-72.0, 88.0))
    dep = -np.random.default_rng(4).uniform(5.0, 3000.0, n)
    build_s = time.perf_counter() - t0
    knw_t = sd.KnW()
    uknw, vknw = ci.make_knw(sd.KnW, (ci.UKNW, ci.VKNW))
    var = ["T", "S", ("UVELMASS", "VVELMASS")]
    kernels = [knw_t, knw_t, (uknw, vknw)]

    pos = sd.Position().from_latlon(x=lon, y=lat, z=dep, data=od)
    if a.sorted:
        key = ((pos.iz * 13 + pos.face) * a.grid + pos.iy) * a.grid + pos.ix
        order = np.argsort(key, kind="stable")
        lon, lat, dep = lon[order], lat[order], dep[order]
        pos = sd.Position().from_latlon(x=lon, y=lat, z=dep, data=od)

    def step():
        return interp.interpolate_device(pos, list(var), list(kernels))

    # the first pass also builds the cell-sorted copy of the located points that the kernels run over (one argsort +
    # one gather per point array, cached with the Position): timed on its own, reported, and part of `e2e`
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record()
    out = step()
    f1.record()
    torch.cuda.synchronize()
    first_ms = f0.elapsed_time(f1)
    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    if clock_sampler is not None:
        clock_sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    clocks = clock_sampler.finish() if clock_sampler is not None else None
    value = n / (ms * 1e-3)
    nan_frac = float(torch.isnan(out[0]).float().mean().item())

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = n * ALGO_BYTES_PER_POINT / (ms * 1e-3) / 1e9
    field_mb = 4 * 13 * a.grid * a.grid * a.levels * 8 / 1e6
    line = {
        "metric": "interpolated points/sec", "value": value, "unit": "points/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"OceInterp-style sampling of T, S (masked 9-point KnW, one multi-field launch) and (u, v) (particle kernels) at {n} "
                        f"{'cell-sorted' if a.sorted else 'scattered'} points, LLC{a.grid}-shaped grid (13x{a.grid}x{a.grid}x{a.levels}), fp64 fields; "
                        f"a step is one pass over the located points",
            "point_order": "points come in scattered; the product keeps a cell-sorted copy of the located points (built during the first "
                           f"pass: {first_ms:.1f} ms for that pass against {ms:.2f} ms afterwards) and scatters the results back to the caller's order; "
                           "the timed steps reuse that copy, `e2e` pays for it every time",
            "first_pass_ms": first_ms,
            "l2": f"inputs larger than L2 (four fields {field_mb:.0f} MB + {n * 60 / 1e6:.0f} MB of point state); no flush",
            "nan_fraction_T": nan_frac, "dataset_build_s": build_s,
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "interp_scalar_multi_kernel (T, S) + interp_vector_kernel (u, v)",
                     "algorithmic_bytes_per_point": ALGO_BYTES_PER_POINT},
        "gpu_launches": 2 * a.steps,
    }
    if clocks is not None:
        line["clocks"] = clocks
    if not a.no_e2e:
        # host arrays in -> locate -> interpolate -> host arrays out: sd.OceInterp, the call a user makes (large batches
        # take the library's pipelined host entry, sdk_interp_host)
        def timed(call, repeat=4):
            tt = []
            for _ in range(repeat):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = call()
                tt.append(time.perf_counter() - t0)
                del res
            return min(tt[1:])

        px, py, pz = (sd.pinned(n) for _ in range(3))
        px[:], py[:], pz[:] = lon, lat, dep
        t_model = 0.0
        best = timed(lambda: sd.OceInterp(od, list(var), px, py, pz, t_model, kernel_list=list(kernels)))
        pageable = timed(lambda: sd.OceInterp(od, list(var), lon, lat, dep, t_model, kernel_list=list(kernels)))
        two_step = timed(lambda: sd.Position().from_latlon(x=lon, y=lat, z=dep, data=od).interpolate(list(var), list(kernels)), repeat=3)
        line["e2e"] = {"value": n / best, "unit": "points/s", "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 32 * n,
                       "api": "sd.OceInterp(od, [T, S, (u, v)], x, y, z, t, kernel_list=...) with the coordinates in pinned numpy arrays "
                              "(sd.pinned) and the results in pinned numpy arrays; wall clock around the call, best of 3 after one warm-up",
                       "ms": best * 1e3,
                       "pageable_input_points_per_s": n / pageable,
                       "two_step_points_per_s": n / two_step,
                       "two_step_api": "Position.from_latlon + Position.interpolate on ordinary numpy arrays (all rel-coords computed, "
                                       "cell-sorted copy built, results through pinned memory)"}
    if not a.no_cpu_baseline:
        from oracle import interp_oracle as io

        m = a.cpu_sample
        t0 = time.perf_counter()
        opos = io.OraclePosition(od, x=lon[:m], y=lat[:m], z=dep[:m])
        orc = io.InterpOracle(od, opos)
        t1 = time.perf_counter()
        ref = orc.interpolate(list(var), list(kernels))
        t2 = time.perf_counter()
        # the sample doubles as a parity check of this run
        got = interp.interpolate(sd.Position().from_latlon(x=lon[:m], y=lat[:m], z=dep[:m], data=od), list(var), list(kernels))
        worst = 0.0
        for g, r in zip(ci.flatten_result(got), ci.flatten_result(ref)):
            ci.assert_values_close("bench sample", g, r)
            ok = ~np.isnan(r)
            if ok.any():
                worst = max(worst, float(np.abs(g[ok] - r[ok]).max() / max(np.abs(r[ok]).max(), 1e-300)))
        line["cpu_baseline"] = {"value": m / (t2 - t1), "unit": "points/s", "cores": 1, "kind": "port",
                                "sample": f"{m} of the same points, interpolation only (oracle/interp_oracle.py, numpy + C topology walk); "
                                          f"locating them took {t1 - t0:.1f} s more",
                                "parity_on_sample_max_rel_err": worst}
    return line


def main():
    a = parse()
    import torch

    torch.cuda.set_device(0)
    sys.path.insert(0, ROOT)
    from bench import ClockSampler

    line = measure(points=a.points, grid=a.grid, levels=a.levels, steps=a.steps, warmup=a.warmup, cpu_sample=a.cpu_sample,
                   cpu_baseline=not a.no_cpu_baseline, e2e=not a.no_e2e, presorted=a.sorted, clock_sampler=ClockSampler(0))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
