#!/usr/bin/env python3
"""Launch-geometry / math sweep of the pricing kernel on one GPU (run under gpurun).

For every configuration: 10M device-resident inputgen options, 2 warm-up ROIs, then the best of 3 ROIs of
NUM_RUNS=100 launches, timed by the library's CUDA events.  Prints options/s and algorithmic GB/s.
"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402


def measure(n, fp_bytes, runs=100, reps=3, **kw):
    with host.BlackScholesGPU(n, fp_bytes=fp_bytes, host_staging=False, with_dgrefval=False, **kw) as bs:
        bs.fill_synthetic(0)
        bs.run(runs)
        bs.run(runs)
        best = min((bs.run(runs), bs.timing()["roi_ms"])[1] for _ in range(reps))
        launch = bs.launch()
    per_launch_us = best / runs * 1e3
    gbs = host.bytes_per_option(fp_bytes) * n / (per_launch_us * 1e-6) / 1e9
    return {"us_per_launch": per_launch_us, "gbs": gbs, "gopts": n / (per_launch_us * 1e-6) / 1e9, **launch}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rows = []
    maths = [("fast", host.MATH_FAST), ("ieee", host.MATH_IEEE)]
    sweep = list(itertools.product(maths, (1, 2, 4), (128, 256), (0, 2, 3, 4, 6, 8)))
    if args.quick:
        sweep = list(itertools.product(maths, (1, 2), (256,), (0,)))
    print("%-5s %-4s %6s %7s %8s %7s | %10s %9s %9s" % ("fp", "math", "unroll", "threads", "blk/SM", "blocks", "us/launch", "GB/s", "Gopt/s"))
    for (mname, m), unroll, threads, bps in sweep:
        try:
            r = measure(args.n, 4, math=m, unroll=unroll, threads_per_block=threads, blocks_per_sm=bps)
        except Exception as e:
            print("fp32 %s u%d t%d b%d failed: %s" % (mname, unroll, threads, bps, e))
            continue
        r.update(fp=32, unroll=unroll, blocks_per_sm=bps)
        rows.append(r)
        print("%-5s %-4s %6d %7d %8d %7d | %10.2f %9.1f %9.2f" % ("fp32", mname, unroll, threads, bps, r["blocks"], r["us_per_launch"], r["gbs"], r["gopts"]), flush=True)
    for unroll, threads, bps in itertools.product((1, 2, 4), (128, 256), (0, 2, 4)):
        if args.quick and (threads != 256 or bps != 0):
            continue
        try:
            r = measure(args.n, 8, runs=20, unroll=unroll, threads_per_block=threads, blocks_per_sm=bps)
        except Exception as e:
            print("fp64 u%d t%d b%d failed: %s" % (unroll, threads, bps, e))
            continue
        r.update(fp=64, unroll=unroll, blocks_per_sm=bps)
        rows.append(r)
        print("%-5s %-4s %6d %7d %8d %7d | %10.2f %9.1f %9.2f" % ("fp64", "ieee", unroll, threads, bps, r["blocks"], r["us_per_launch"], r["gbs"], r["gopts"]), flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
