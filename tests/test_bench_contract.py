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


def test_reference_arm_follows_our_arm_to_the_strong_1b_workload_at_n_gt_1():
    # bench.py --gpus N > 1 defaults to the 1B-option strong-scaling set (BASELINE.json configs[3]); the reference arm must
    # report the same metric / config / scaling on a bounded sample of it (its driver cannot hold 1B options: int numOptions,
    # 36 GB of AoS), from rank 0 only
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert cp.returncode == 0, cp.stderr[-2000:]
    d = json.loads([l for l in cp.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong"
    assert "1B-option" in d["config"]["workload"] and d["config"]["sample_options"] < 10_000_000
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and "full_size" not in d
