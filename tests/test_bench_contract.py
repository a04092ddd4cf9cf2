"""CPU: bench.py's reference arm (the CPU oracle) prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_b200_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--frames-in-flight", "2"]])
def test_b200_arm_json_line(extra):
    """GPU: the product arm on the reference's own small case (C1) prints one JSON line with the contract's keys, a green frame check,
    launches of its own kernels, a roofline from the production walk's counters and an e2e leg with host buffers."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "C1", "--steps", "4", "--warmup", "3", "--breakdown", "none",
                        "--min-seconds", "0", "--no-cpu-baseline"] + extra, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "frame_check", "mrays_traversed_per_s"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["value"] > 0 and d["gpu_launches"] > 0 and d["dtype"] == "f32"
    assert d["frame_check"]["status"] == "ok", d["frame_check"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 800 * 800 * 4 and d["e2e"]["value"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and 0 < rf["frac"] < 1.5 and rf["achieved"] > 0 and rf["walk_counters"]["rays"] > 0
    assert rf["node_fetch"]["peak"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    if extra:
        assert d["config"]["frames_in_flight"].startswith("2") and d["frame_latency_ms"] > 0
        assert d["frame_check"]["frames_in_flight"] == "the RGBA8 frames of all lanes are identical"
