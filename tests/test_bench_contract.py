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


@pytest.mark.parametrize("name", ["r1_bench_tf32.json", "r1_bench_bf16.json", "r2_bench_tf32.json", "r2_bench_bf16.json",
                                  "r2_bench_fp16.json"])
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


def test_round2_line_carries_the_baseline_configs():
    """The driver-style line of round 2 also measures BASELINE configs 3, 4 and 1 (sub-records), compares with the
    ORACLE, times the CPU arm on the full batch, and both arms print the same config dictionary."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_tf32.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_reference_arm.json")))
    assert d["config"] == ref["config"]
    assert "full batch" in d["cpu_baseline"]["sample"] and "full batch" in ref["cpu_baseline"]["sample"]
    for m in ("bf16", "fp16"):
        c3 = d["config3"][m]
        assert c3["value"] > 0 and c3["e2e"]["value"] > 0 and "max_abs" in c3["parity_vs_oracle"]
        assert c3["e2e"]["h2d_bytes_per_step"] == 256 * 80 * 172 * 4 and c3["e2e"]["d2h_bytes_per_step"] == 256 * 172 * 256 * 4
    assert d["config3"]["scaling"] == "strong"
    for m, r in d["config4"].items():
        if isinstance(r, dict):
            assert r["chunks_equal_unchunked"]["bit_equal"] is True and r["parity_vs_oracle"]["max_abs"] < 1e-3 + (m == "bf16")
    assert {"tf32", "fp16", "bf16"} <= set(d["config1"])
    q = d["quality"]
    assert q["tf32"]["max_abs_vs_oracle"] <= 1e-3 and q["fp16"]["max_abs_vs_oracle"] <= 1e-3
    assert d["roofline"]["frac"] == pytest.approx(d["roofline"]["achieved"] / d["roofline"]["peak"])
    n8 = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_n8.json")))
    assert n8["n_gpus"] == 8 and n8["config3"]["bf16"]["value"] >= 1e5 and n8["config3"]["bf16"]["e2e"]["value"] >= 1e5


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches both arms the same way; for N > 1 that is torchrun.  Rank 0 alone runs the CPU arm and
    prints the line, the other rank exits 0 without work (and without a GPU or a process group)."""
    port = 29500 + os.getpid() % 400
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["value"] == d["value"] > 0


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the product arm exits non-zero with a clear message instead of timing anything else."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
