"""Shared helpers for the -m gpu parity tests (all GPU work goes through the C ABI via p3arsec_b200.host)."""
import numpy as np

import oracle_lib
from p3arsec_b200 import host

FP32_ABS_TOL = 1e-4      # BASELINE.json north_star: max |delta| <= 1e-4 for fp32 (the reference's ERR_CHK threshold)
FP64_REL_TOL = 1e-9      # BASELINE.json north_star: <= 1e-9 relative for fp64
# Cancellation floor of the fp64 comparison: s*N(d1) - fv*N(d2) is formed from terms of size ~spot, so two
# correct evaluations whose exp()/log() differ in the last ulp may differ by a few ulp(spot) ~ 1e-13 absolute
# however small the price is.  The relative bound is therefore applied as |d| <= 1e-9*|ref| + 1e-12.
FP64_ABS_FLOOR = 1e-12


def inputgen_like(n, seed, dtype=np.float32):
    """Random options from the distribution of the synthetic inputgen table."""
    rng = np.random.RandomState(seed)
    s = np.round(rng.uniform(20.0, 120.0, n), 2)
    k = np.round(s * rng.uniform(0.7, 1.3, n), 2)
    r = rng.choice([0.0250, 0.0275, 0.0500, 0.0750, 0.1000], n)
    v = np.round(rng.uniform(0.05, 0.65, n), 2)
    t = np.round(rng.uniform(0.05, 1.00, n), 2)
    o = (rng.uniform(size=n) < 0.5).astype(np.int32)
    return tuple(a.astype(dtype) for a in (s, k, r, v, t)) + (o,)


def gpu_prices(inputs, fp_bytes=4, num_runs=1, dgrefval=None, err_chk=False, **kw):
    s, k, r, v, t, o = inputs
    with host.BlackScholesGPU(len(s), fp_bytes=fp_bytes, **kw) as bs:
        bs.set_inputs(s, k, r, v, t, o, dgrefval)
        errs = bs.price(num_runs, err_chk)
        out = bs.prices.copy()
        bad = bs.errors() if err_chk else None
    return out, errs, bad


def oracle_prices(inputs, fp_bytes=4):
    s, k, r, v, t, o = inputs
    return oracle_lib.price_map(s, k, r, v, t, o, fp_bytes)


def magnitude_scale(spot, strike):
    """1 inside the inputgen range (spot, strike <= 128); beyond it the fp32 bound grows with the size of
    the operands, i.e. stays the same number of ulps: a price near 1000 is itself quantised to 6e-5, and
    the reference's own fp32 build misses its fp64 build by 1.7e-4 on tests/golden/edge2k (3.2e-5 on the
    in-range table1k)."""
    mag = np.maximum(np.abs(np.asarray(spot, np.float64)), np.abs(np.asarray(strike, np.float64)))
    return np.maximum(1.0, mag / 128.0)


def assert_parity(got, ref, fp_bytes, what="", scale=None):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.all(np.isfinite(got) == np.isfinite(ref)), what
    m = np.isfinite(ref)
    d = np.abs(got[m] - ref[m])
    if fp_bytes == 4:
        if scale is not None:
            d = d / np.asarray(scale, np.float64)[m]
        worst = float(d.max()) if d.size else 0.0
        assert worst <= FP32_ABS_TOL, "%s fp32 max|delta| = %.3e > 1e-4 at %d" % (what, worst, int(d.argmax()))
        return worst
    bound = FP64_REL_TOL * np.abs(ref[m]) + FP64_ABS_FLOOR
    over = d - bound
    assert not (over > 0).any(), "%s fp64: |delta|=%.3e vs ref=%.6e at %d" % (
        what, float(d[over.argmax()]), float(ref[m][over.argmax()]), int(over.argmax()))
    rel = d / np.maximum(np.abs(ref[m]), 1e-300)
    return float(d.max()) if d.size else 0.0
