"""bench.py's reference arm on CPU: the JSON line carries every key the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "simsmall",
                         "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert cp.returncode == 0, cp.stderr[-2000:]
    lines = [l for l in cp.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "options_priced_per_sec" and d["unit"] == "options/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "options/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_refuses_to_run_without_a_gpu():
    import ctypes
    sys.path.insert(0, ROOT)
    from p3arsec_b200 import host
    if host.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert cp.returncode != 0 and "no CUDA device" in (cp.stderr + cp.stdout)
