#!/usr/bin/env python3
"""Small ncu target: device-resident inputgen options, a few plain (non-graph) launches of the pricing kernel.

    ncu --set full --clock-control none --import-source on -k regex:bs_map -s 2 -c 3 -o gpurun_out/prof \
        python tools/profile_target.py --n 10000000 --fp 4 --math fast --runs 6
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--fp", type=int, default=4)
ap.add_argument("--math", default="default")
ap.add_argument("--runs", type=int, default=6)
ap.add_argument("--unroll", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--blocks-per-sm", type=int, default=0)
a = ap.parse_args()
m = {"default": host.MATH_DEFAULT, "ieee": host.MATH_IEEE, "fast": host.MATH_FAST}[a.math]
with host.BlackScholesGPU(a.n, fp_bytes=a.fp, host_staging=False, with_dgrefval=False, math=m, use_graph=False,
                          unroll=a.unroll, threads_per_block=a.threads, blocks_per_sm=a.blocks_per_sm) as bs:
    bs.fill_synthetic(0)
    bs.run(a.runs)
    print(bs.launch(), bs.timing()["roi_ms"] / a.runs * 1e3, "us/launch")
