#!/usr/bin/env python3
"""simsmall (4,096 options x NUM_RUNS=100) is launch-bound: ~1.3 us per graph kernel node.  Does programmatic dependent launch
between the runs (BS_GPU_FLAG_PDL) shorten the gap for such tiny launches?  Prints us per run for a few sizes, with/without."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402

for n in (4096, 16384, 65536, 262144, 1048576):
    row = []
    for pdl in (False, True):
        for graph in (True, False):
            with host.BlackScholesGPU(n, host_staging=False, with_dgrefval=False, pdl=pdl, use_graph=graph) as bs:
                bs.fill_synthetic(0)
                for _ in range(5):
                    bs.run(100)
                t = []
                for _ in range(21):
                    bs.run(100)
                    t.append(bs.timing()["roi_ms"] * 10.0)  # us per run
                row.append("%s%s %.3f" % ("pdl" if pdl else "plain", "+graph" if graph else "", statistics.median(t)))
    print("n=%-8d us/run: %s" % (n, " | ".join(row)), flush=True)
