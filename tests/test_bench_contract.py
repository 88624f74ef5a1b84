"""bench.py contract (CPU): the JSON line the driver parses carries every required key -- checked on the
committed B200 lines under profiles/ and, live, on the CPU arm (`--impl reference`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].split(" (")[0] in baseline["metric"]


@pytest.mark.parametrize("name", ["r1_bench_tf32.json", "r1_bench_bf16.json"])
def test_committed_b200_lines_follow_the_contract(name):
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * 80 * 172 * 4 and d["e2e"]["d2h_bytes_per_step"] == 16 * 172 * 256 * 4
    assert abs(d["value"] - 16 * 172 * 256 / 22050 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    b = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("port", "reference")
    assert "l2" in d["config"]                           # how L2 is handled between timed iterations


def test_reference_arm_line_live():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
