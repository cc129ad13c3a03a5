"""The bench.py contract the driver relies on, checked without a GPU: the reference arm (`--impl reference`) runs on the
host cores and prints ONE JSON line with the agreed keys; the workload description is identical for both arms."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ)
    env.pop("RANK", None)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample-rows", "8192", "--gpus", "1"], capture_output=True, text=True, timeout=600, env=env,
                          cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # everything else (library banners, warnings) goes to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "queries/s"
    assert d["metric"].startswith("QPS, exact IP top-100 over 21000000x1024")
    assert d["config"]["rows"] == 21_000_000 and d["config"]["batch"] == 4096 and d["config"]["k"] == 100
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a step of this arm is one pass over the bounded sample: steps * ms_per_step is what the run really took
    assert 0 < d["ms_per_step"] < 120_000 and d["ms_per_full_step_extrapolated"] > d["ms_per_step"]
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample-rows", "4096", "--gpus", "2"], capture_output=True, text=True, timeout=300, env=env,
                          cwd=ROOT)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_both_arms_describe_the_same_workload():
    sys.path.insert(0, ROOT)
    import bench

    c = bench.workload_config(21_000_000, 4096, 100)
    assert set(c) == {"workload", "rows", "dim", "batch", "k", "l2"} and "configs[3]" in c["workload"]
    p = bench.load_peaks()
    assert p["hbm_gbs"] > 0 and p["bf16_tflops"] >= p["bf16_tflops_sustained"] > 0 and p["source"] in ("measured", "fallback")
