"""CPU: the reference arm of bench.py (`--impl reference`) prints one JSON line with the keys the driver reads.
(The GPU arm needs a device; its line is checked by the driver itself.)"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
         "--cpu-sample", "3000", "--grid", "30", "--levels", "10", "--particles", "3000"],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
    )
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "particle cell-crossings/sec" and d["unit"] == "crossings/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_follows_the_survey_byte_model():
    """`roofline.achieved` = SURVEY.md 8d's algorithmic bytes (88 / 72 B per crossing for 3-D / surface fp64 fields + 256 B per
    particle and launch) over the kernel time; `traffic` comes from the committed ncu capture of that kernel, or is None."""
    import sys

    sys.path.insert(0, ROOT)
    import bench

    r = bench.roofline_of("advect_kernel<double,false,1>", 1.0, 11.2e6, 1_000_000, "f64", True, *bench.load_traffic("c2"))
    assert r["algorithmic_bytes_per_launch"] == 11.2e6 * 88 + 1_000_000 * 256
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / 1e-3 / 1e9) < 1e-9 and r["unit"] == "GB/s" and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1
    assert r["traffic"] and "ncu" in r["traffic_source"]
    s = bench.roofline_of("advect_kernel<double,false,3>", 3.7, 46.2e6, 12_500_000, "f64", False, *bench.load_traffic("c4"))
    assert s["algorithmic_bytes_per_launch"] == 46.2e6 * 72 + 12_500_000 * 256
    assert s["traffic"] > s["algorithmic_bytes_per_launch"]  # sector granularity on a grid no cache holds
    assert bench.load_traffic("f32") == (None, None)
