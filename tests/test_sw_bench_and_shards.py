"""Host logic around the swaptions Map on CPU: the reference arm's JSON line, and the N>1 sharding rule of
tools/sw_bench.py (contiguous swaption ranges, swaption i keeps seed swaption_seed + i) checked with two gloo ranks
pricing their shards through the oracle."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

import sw_oracle_lib as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line_through_bench_py():
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "swaptions_simsmall",
                         "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert cp.returncode == 0, cp.stderr[-2000:]
    lines = [l for l in cp.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "swaption_trials_per_sec" and d["unit"] == "trials/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and "swaptions simsmall" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sw_bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                         "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert cp.returncode == 0 and cp.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    sys.path.insert(0, ROOT)
    from p3arsec_b200 import swaptions as sw
    if sw.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sw_bench.py"), "--steps", "1", "--workload", "simsmall"],
                        capture_output=True, text=True, timeout=600)
    assert cp.returncode != 0 and "no CUDA device" in (cp.stderr + cp.stdout)


WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import sw_oracle_lib as so
from p3arsec_b200 import swaptions as sw
from p3arsec_b200.dist import Ranks, shard_range
r = Ranks(backend="gloo")
ns, trials = 13, 512
seed, p, y, f = sw.make_portfolio(ns)
first, count = shard_range(ns, r.world, r.rank)
mean, err = so.price_map(p[first:first + count], y[first:first + count], f[first:first + count], seed + first, trials, nthreads=1)
r.barrier()
out = {"rank": r.rank, "first": first, "count": count, "mean": mean.tolist(), "err": err.tolist(),
       "trials_total": r.sum(count * trials), "max": r.max(1.0 + r.rank)}
r.close()
json.dump(out, open(os.path.join(%(out)r, "rank%%d.json" %% r.rank), "w"))
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_gloo_ranks_price_the_portfolio_in_shards(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    for attempt in range(3):
        cp = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                             "--master-port", str(_free_port()), str(script)], capture_output=True, text=True, timeout=300)
        if cp.returncode == 0:
            break
    assert cp.returncode == 0, cp.stdout[-2000:] + cp.stderr[-2000:]
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in (0, 1)]
    assert (res[0]["first"], res[0]["count"], res[1]["first"], res[1]["count"]) == (0, 7, 7, 6)
    assert all(d["trials_total"] == 13 * 512 and d["max"] == 2.0 for d in res)
    # shards priced with seed + first concatenate to exactly what one process computes for the whole portfolio
    from p3arsec_b200 import swaptions as sw
    seed, p, y, f = sw.make_portfolio(13)
    mean, err = so.price_map(p, y, f, seed, 512)
    got_mean = np.array(res[0]["mean"] + res[1]["mean"])
    got_err = np.array(res[0]["err"] + res[1]["err"])
    assert got_mean.tobytes() == mean.tobytes()
    assert np.array_equal(np.isnan(got_err), np.isnan(err)) and np.array_equal(got_err[~np.isnan(err)], err[~np.isnan(err)])
