"""The reference arm of bench.py runs without a GPU (it times the CPU oracle): its JSON line must carry
the contract's keys, and the `ours` arm must refuse to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    rows = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(rows) == 1, "exactly one JSON line on stdout"
    d = json.loads(rows[0])
    assert d["impl"] == "reference" and d["unit"] == "shapes/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "metric" in d and "config" in d and "ms_per_step" in d


def test_ours_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "GPU" in (r.stderr + r.stdout)
