"""The N>1 host logic of bench.py on CPU: two gloo ranks, contiguous shards, max-over-ranks timing.
No data-path collective exists (shards are independent); only the barrier / reductions are exercised."""
import os
import socket
import subprocess
import sys

import pytest

from p3arsec_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
from p3arsec_b200.dist import Ranks, shard_range
r = Ranks(backend="gloo")
n = 1000003
first, count = shard_range(n, r.world, r.rank)
r.barrier()
out = {"rank": r.rank, "world": r.world, "first": first, "count": count,
       "max": r.max(10.0 + r.rank), "sum": r.sum(count)}
r.barrier()
r.close()
json.dump(out, open(os.path.join(%(out)r, "rank%%d.json" %% r.rank), "w"))  # one file per rank: stdout lines of two ranks can interleave
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,world", [(10, 3), (1000003, 8), (7, 8), (0, 4), (2**31 - 1, 8)])
def test_shard_range_partitions_like_the_static_parallel_for(n, world):
    spans = [shard_range(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == n
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)  # first n % world ranks get +1


def test_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    for attempt in range(3):  # a probed-free port can be taken before torchrun binds it: retry on a new one
        port = _free_port()
        cp = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                             "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                            capture_output=True, text=True, timeout=300)
        if cp.returncode == 0:
            break
    assert cp.returncode == 0, cp.stdout[-2000:] + cp.stderr[-2000:]
    import json
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in (0, 1)]
    assert [d["rank"] for d in res] == [0, 1] and all(d["world"] == 2 for d in res)
    assert res[0]["first"] == 0 and res[0]["count"] == 500002 and res[1]["first"] == 500002 and res[1]["count"] == 500001
    assert all(d["max"] == 11.0 and d["sum"] == 1000003.0 for d in res)   # max over ranks / whole-job units


def test_reference_arm_runs_only_on_rank0(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=env, timeout=120)
    assert cp.returncode == 0 and cp.stdout.strip() == ""
