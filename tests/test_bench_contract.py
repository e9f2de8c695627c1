"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for both arms.

The reference arm (the CPU oracle port on the host cores) runs without a GPU; the product arm is a
`-m gpu` test at a handful of steps."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                       text=True, cwd=ROOT, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "V2E+voxel Gpixel-frames/s" and d["unit"] == "Gpixel-frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "cfg3" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line():
    d = run_bench("--steps", "6", "--warmup", "3", "--cpu-budget", "2")
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "V2E+voxel Gpixel-frames/s" and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert d["steps"] == 6 and d["warmup"] >= 3 and d["value"] > 1.0
    assert d["gpu_launches"] >= 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 720 * 1280 * 9 * 4 and e["d2h_bytes_per_step"] == 8
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["algorithmic_bytes_per_pack"] == 92 * 720 * 1280
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
