"""bench.py contract checks that run without a GPU: the reference arm (the reference's own CPU sampler through oracle/_ref, else
the C port) prints ONE JSON line with the keys the driver reads; the product arm refuses to run without the CUDA library's device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vertex_updates_per_sec" and d["unit"] == "vertex-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_product_arm_needs_the_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm runs (covered by the -m gpu tests and the bench itself)")
    r = run_bench("--workload", "small", "--steps", "1", "--warmup", "3", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0                          # fails loudly: no CPU fallback behind bench.py either
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
