"""bench.py on a box without a GPU: the reference arm (the CPU oracle, rank 0 only) prints the contract's JSON line, the
product arm refuses to run -- there is no CPU fallback to time."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=300, env=e)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--workload", "C1", "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1      # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 20000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "20000 cells" in cb["sample"]
    assert set(cb["variants"]) == {"fast_build_1_rank", "parity_build_all_ranks", "parity_build_1_rank"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        assert d["metric"] == json.load(f)["metric"]


def test_reference_arm_other_ranks_do_nothing():
    r = run_bench("--impl", "reference", "--workload", "C1", "--steps", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_without_a_gpu(fcmod):
    if fcmod.lib.fc_device_count() >= 1:
        import pytest
        pytest.skip("a GPU is visible")
    r = run_bench("--workload", "C2", "--steps", "1", "--no-e2e", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())      # and no number is printed
