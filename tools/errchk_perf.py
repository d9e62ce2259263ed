#!/usr/bin/env python3
"""ERR_CHK variant of the kernel (32 B/option fp32, 60 B/option fp64): time per 10M-option launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3arsec_b200 import host  # noqa: E402

for fp in (4, 8):
    with host.BlackScholesGPU(10_000_000, fp_bytes=fp, host_staging=False, with_dgrefval=True) as bs:
        bs.fill_synthetic(0)
        for chk in (False, True):
            bs.run(100, err_chk=chk)
            best = 1e9
            for _ in range(5):
                errs = bs.run(100, err_chk=chk)
                best = min(best, bs.timing()["roi_ms"])
            us = best * 10
            b = host.bytes_per_option(fp, chk)
            print("fp%d err_chk=%d: %.2f us/launch, %.0f GB/s (%d B/option), %.1f G options/s, Num Errors %d" % (
                fp * 8, chk, us, b * 10_000_000 / us / 1e3, b, 10_000_000 / us / 1e3, errs))
