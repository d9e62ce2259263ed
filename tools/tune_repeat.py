#!/usr/bin/env python3
"""Repeat-and-interleave comparison of a few launch configurations (run under gpurun).

Each round visits every configuration once (so drift and neighbours affect all alike); per configuration the
median and the minimum over the rounds are printed.  Times are the library's CUDA-event ROI times.
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402

M = {"fast": host.MATH_FAST, "ieee": host.MATH_IEEE, "reference": host.MATH_REFERENCE}
# (fp_bytes, math, unroll, threads, blocks_per_sm, variant)   variant: 1 = pipelined loads, 2 = traffic probe
CONFIGS = {
    "fp32": [(4, "fast", 1, 256, 4, 0), (4, "fast", 1, 128, 8, 0), (4, "fast", 2, 256, 0, 0), (4, "fast", 1, 256, 0, 1), (4, "fast", 1, 256, 3, 1),
             (4, "fast", 1, 128, 0, 1), (4, "fast", 1, 256, 2, 1), (4, "fast", 1, 256, 4, 2), (4, "fast", 1, 128, 8, 2), (4, "fast", 2, 256, 4, 2),
             (4, "fast", 1, 256, 6, 2), (4, "fast", 1, 256, 8, 2),
             (4, "ieee", 1, 256, 0, 0), (4, "ieee", 1, 256, 8, 0), (4, "ieee", 1, 256, 0, 1), (4, "ieee", 1, 128, 0, 1), (4, "ieee", 2, 256, 0, 0)],
    "pdl": [(4, "fast", 1, 256, 4, 0), (4, "fast", 1, 256, 4, 16), (4, "fast", 2, 256, 0, 0), (4, "fast", 2, 256, 0, 16),
            (4, "fast", 1, 256, 4, 2), (4, "fast", 1, 256, 4, 18),
            (8, "fast", 1, 256, 0, 1), (8, "fast", 1, 256, 0, 17), (4, "ieee", 1, 256, 0, 0), (4, "ieee", 1, 256, 0, 16)],
    "tma": [(4, "fast", 1, 256, 4, 0), (4, "fast", 1, 256, 0, 4), (4, "fast", 1, 256, 1, 4), (4, "fast", 1, 256, 4, 2), (4, "fast", 1, 256, 0, 6),
            (4, "ieee", 1, 256, 0, 0), (4, "ieee", 1, 256, 0, 4),
            (8, "fast", 1, 256, 0, 1), (8, "fast", 1, 256, 0, 4), (8, "fast", 1, 256, 1, 4), (8, "fast", 1, 256, 0, 6), (8, "ieee", 1, 256, 0, 4)],
    "fp64r2": [(8, "fast", 1, 256, 0, 1), (8, "fast", 1, 224, 0, 1), (8, "fast", 1, 192, 0, 1), (8, "fast", 1, 160, 0, 1), (8, "fast", 1, 128, 0, 1),
               (8, "fast", 1, 96, 0, 1), (8, "fast", 1, 64, 0, 1), (8, "fast", 1, 256, 0, 0), (8, "fast", 1, 192, 0, 0), (8, "fast", 1, 128, 0, 0),
               (8, "fast", 2, 128, 0, 0), (8, "fast", 2, 128, 0, 1), (8, "fast", 2, 256, 0, 1), (8, "fast", 1, 256, 0, 4), (8, "fast", 1, 256, 0, 2), (8, "fast", 2, 256, 0, 2)],
    "fp64ref": [(8, "reference", 1, 256, 0, 0), (8, "ieee", 1, 256, 0, 0), (8, "ieee", 1, 256, 0, 1), (4, "reference", 1, 256, 0, 0), (4, "ieee", 1, 256, 0, 0)],
    "fp64tma2": [(8, "fast", 1, 256, 0, 4), (8, "fast", 1, 256, 0, 68), (8, "fast", 1, 256, 0, 36), (8, "fast", 1, 256, 0, 1), (8, "fast", 1, 256, 0, 6),
                 (8, "fast", 1, 256, 0, 70), (8, "fast", 1, 256, 0, 2)],
    "fp64tma": [(8, "fast", 1, 256, 0, 1), (8, "fast", 1, 256, 0, 4), (8, "fast", 1, 256, 0, 36), (8, "fast", 1, 256, 0, 2), (8, "fast", 2, 256, 0, 2),
                (8, "fast", 1, 256, 0, 6), (8, "fast", 1, 256, 0, 38), (8, "ieee", 1, 256, 0, 4), (8, "ieee", 1, 256, 0, 36), (8, "ieee", 1, 256, 0, 1)],
    "fp64": [(8, "fast", 1, 256, 0, 0), (8, "fast", 2, 256, 0, 0), (8, "fast", 1, 256, 0, 1), (8, "fast", 1, 128, 0, 1), (8, "fast", 2, 128, 0, 1),
             (8, "fast", 2, 256, 0, 1), (8, "fast", 1, 64, 0, 1), (8, "fast", 1, 256, 0, 2), (8, "fast", 2, 256, 0, 2),
             (8, "ieee", 1, 256, 0, 0), (8, "ieee", 2, 256, 0, 0), (8, "ieee", 1, 256, 0, 1), (8, "ieee", 1, 128, 0, 1)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="fp32")
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--runs", type=int, default=100)
    a = ap.parse_args()
    ctxs = []
    for fp, m, u, t, b, var in CONFIGS[a.which]:
        os.environ["BS_GPU_TMA_WIDE"] = "2" if var & 64 else "1" if var & 32 else "0"   # read by bs_gpu_init_ex: shape of the fp64 TMA kernel
        bs = host.BlackScholesGPU(a.n, fp_bytes=fp, host_staging=False, with_dgrefval=False, math=M[m], unroll=u,
                                  threads_per_block=t, blocks_per_sm=b, variant=var & 7, pdl=bool(var & 16))
        bs.fill_synthetic(0)
        bs.run(a.runs)
        ctxs.append(((fp, m, u, t, b, var), bs, []))
    for _ in range(a.rounds):
        for cfg, bs, times in ctxs:
            bs.run(a.runs)
            times.append(bs.timing()["roi_ms"] / a.runs * 1e3)
    print("%-4s %-11s %6s %7s %6s %7s | %9s %9s | %9s %9s" % ("fp", "math", "unroll", "threads", "blk/SM", "blocks", "med us", "min us", "med GB/s", "Gopt/s"))
    for (fp, m, u, t, b, var), bs, times in sorted(ctxs, key=lambda c: statistics.median(c[2])):
        med, mn = statistics.median(times), min(times)
        m = ("PROBE" if var & 2 else m + ("+p" if var & 1 else "")) + ("/tma" if var & 4 else "") + ("w" if var & 32 else "") + ("x2" if var & 64 else "") + ("/pdl" if var & 16 else "")
        print("%-4d %-11s %6d %7d %6d %7d | %9.2f %9.2f | %9.1f %9.2f" % (fp * 8, m, u, t, b, bs.launch()["blocks"], med, mn,
              host.bytes_per_option(fp) * a.n / med / 1e3, a.n / med / 1e3))
        bs.close()


if __name__ == "__main__":
    main()
