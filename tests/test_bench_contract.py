"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU port of the hot path on the
host cores) runs here and prints ONE JSON line with the keys the driver reads; the GPU arm refuses to run without a device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-400:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("stereo pairs/s") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert abs(d["value"] * d["ms_per_step"] * 1e-3 - 1.0) < 1e-6               # one pair per step
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]      # no number without the CUDA path
