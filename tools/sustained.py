#!/usr/bin/env python3
"""Seconds-long runs of the pricing kernel vs the traffic probe: does the SM work cost HBM bandwidth once the
board sits at its power cap?  Prints per-ROI GB/s with SM clock, power and throttle reasons sampled via NVML."""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not self.stop:
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                              pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)))
            time.sleep(0.01)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=250_000_000)
    ap.add_argument("--runs", type=int, default=1000)
    ap.add_argument("--rois", type=int, default=4)
    ap.add_argument("--fp", type=int, default=4)
    a = ap.parse_args()
    cases = [("fast", dict(math=host.MATH_FAST)), ("probe", dict(variant=2)), ("ieee", dict(math=host.MATH_IEEE)),
             ("fast u2", dict(math=host.MATH_FAST, unroll=2, blocks_per_sm=0, threads_per_block=256))]
    if a.fp == 8:  # fp64: the TMA ring kernel (default), software-pipelined LDG, plain LDG, and the two traffic probes
        cases = [("tma", dict(variant=4)), ("ldg+pipe", dict(variant=1)), ("ldg", dict(variant=0, threads_per_block=256)),
                 ("probe tma", dict(variant=6)), ("probe ldg", dict(variant=2)), ("tma", dict(variant=4))]
    bpo = host.bytes_per_option(a.fp)
    for name, kw in cases:
        with host.BlackScholesGPU(a.n, fp_bytes=a.fp, host_staging=False, with_dgrefval=False, **kw) as bs:
            bs.fill_synthetic(0)
            bs.run(10)
            for i in range(a.rois):
                smp = Sampler()
                smp.start()
                bs.run(a.runs)
                smp.stop = True
                smp.join()
                ms = bs.timing()["roi_ms"]
                sm = sorted(r[0] for r in smp.rows)
                pw = [r[1] for r in smp.rows]
                reasons = 0
                for r in smp.rows:
                    reasons |= r[2]
                print("%-8s ROI %d: %8.1f ms  %7.1f GB/s  %7.2f Gopt/s | SM MHz med %d min %d  mem MHz %d  power max %.0f W avg %.0f W  reasons 0x%x" % (
                    name, i, ms, bpo * a.n * a.runs / ms / 1e6, a.n * a.runs / ms / 1e6, sm[len(sm) // 2], sm[0], smp.rows[-1][3], max(pw),
                    sum(pw) / len(pw), reasons), flush=True)
        time.sleep(2.0)


if __name__ == "__main__":
    main()
