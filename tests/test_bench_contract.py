"""bench.py contract checks that need no GPU: the reference arm (the oracle on the host cores) prints exactly one JSON
line with the keys of the contract, and the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "1024", "--replicas", "2")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "paths tracked/sec" and d["unit"] == "paths/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cyclic-7 polyhedral" in d["config"]["workload"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    p = run("--steps", "1", "--warmup", "1", "--replicas", "1", "--no-cpu-baseline")
    assert p.returncode != 0
    assert not any(l.lstrip().startswith("{") for l in p.stdout.splitlines())   # no bench line from a CPU path
