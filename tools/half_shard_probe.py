import os, sys
sys.path.insert(0, os.getcwd())
from p3arsec_b200 import host
for n in (10_000_000, 5_000_000, 4_500_000, 2_500_000):
    with host.BlackScholesGPU(n, fp_bytes=4, host_staging=False, with_dgrefval=False) as bs:
        bs.fill_synthetic(0)
        bs.run(100); bs.run(100)
        best = min((bs.run(100), bs.timing()["roi_ms"])[1] for _ in range(5))
        print("n=%d: %.2f us/launch, %.0f GB/s algorithmic" % (n, best * 10, 28 * n / (best * 10) / 1e3))
