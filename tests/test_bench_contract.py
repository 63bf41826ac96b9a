"""bench.py contract checks that need no GPU: the reference arm (the reference's OpenMP kernels on the host cores) prints one JSON line with the
keys the driver reads, and our arm refuses to run without a CUDA device (the product has no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = _run("--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "cg_matvec_tflops" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("C1:")
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "rows" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    res = _run("--steps", "1", "--warmup", "0", "--workload", "C1", timeout=120)
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
