"""bench.py's contract on the CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints exactly ONE
line on stdout, a JSON object with the keys the driver reads; the GPU arm fails loudly where there is no GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["metric"].startswith("Mrays/s") and d["unit"] == "Mrays/s" and d["value"] > 0 and d["dtype"] == "f32"
    assert d["config"]["workload"] == "config2_spheres_capsules_1080p" and d["config"]["image"] == "1920x1080"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None  # BASELINE.md holds no published number


def test_reference_arm_of_an_extension_workload_says_unavailable():
    r = _run("--impl", "reference", "--workload", "config3")
    assert r.returncode == 0
    d = json.loads(r.stdout.strip())
    assert d["impl"] == "reference" and "unavailable" in d


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu", "--no-config5", timeout=120)
    assert r.returncode != 0 and r.stdout.strip() == ""
