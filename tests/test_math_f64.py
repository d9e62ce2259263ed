"""CPU check of the fp64 fast-math building blocks (p3arsec_b200/csrc/bs_math_f64.h) before any GPU time is
spent: the header is host+device code; on the host the MUFU seeds are emulated at the hardware's ~2^-20 accuracy.
The GPU result itself is checked against the oracle in tests/test_gpu_parity.py."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import golden_path
from gpu_util import FP64_ABS_FLOOR, FP64_REL_TOL, inputgen_like

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("m64") / "math_f64_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "math_f64_host_check.cpp"), "-lm"], check=True)
    return exe


def test_building_blocks_within_a_few_ulp(checker):
    out = dict(l.split() for l in subprocess.run([checker], capture_output=True, text=True, check=True).stdout.splitlines())
    assert float(out["rcp"]) <= 1.0 and float(out["rsqrt"]) <= 1.5
    assert float(out["exp"]) <= 1.5 and float(out["exp_small"]) <= 1.0
    assert float(out["exp_pair_pos"]) <= 1.5 and float(out["exp_pair_neg"]) <= 1.5
    # the 256-entry blocks the blackscholes fp64 kernel uses (one / two polynomial degrees less)
    assert float(out["exp256"]) <= 1.5 and float(out["log256"]) <= 1.5 and float(out["log256_near_1_abs"]) <= 0.1
    assert float(out["exp256_0"]) == 1.0 and abs(float(out["log256_1"])) < 1e-17
    # log feeds d1 = (drift*t + log(s/k)) / den: its ABSOLUTE error matters, measured in ulps of max(1, |log x|)
    assert float(out["log"]) <= 1.5 and float(out["log_near_1_abs"]) <= 0.1
    assert float(out["exp_below_-708"]) == 0.0 and float(out["exp_0"]) == 1.0 and abs(float(out["log_1"])) < 1e-17


def _fast_prices(checker, s, k, r, v, t, o):
    rows = "\n".join("%.17g %.17g %.17g %.17g %.17g %d" % z for z in zip(s, k, r, v, t, o))
    out = subprocess.run([checker, "price"], input=rows, capture_output=True, text=True, check=True).stdout.split()
    return np.array(out[0::2], dtype=np.float64), np.array(out[1::2], dtype=np.int32)


@pytest.mark.parametrize("name", ["hull4", "table1k", "edge2k"])
def test_fast_fp64_price_matches_oracle_on_goldens(checker, name):
    d = oracle_lib.load(golden_path(name, "in.txt"), 8)
    ref = oracle_lib.price_map(d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"], 8)
    got, ok = _fast_prices(checker, d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"])
    ok = ok.astype(bool)
    # rows the fast path declines (|d1| >= 37 in the stress file: exp(-d1^2/2) has underflowed) are priced by the
    # IEEE-order path on the device; inside the inputgen range every row stays on the fast path
    den = d["volatility"] * np.sqrt(d["otime"])
    d1 = ((d["rate"] + 0.5 * d["volatility"] ** 2) * d["otime"] + np.log(d["sptprice"] / d["strike"])) / den
    assert (ok | (np.abs(d1) > 36.9)).all()
    assert ok.all() if name != "edge2k" else ok.mean() > 0.8
    delta = np.abs(got - ref)[ok]
    assert (delta <= FP64_REL_TOL * np.abs(ref[ok]) + FP64_ABS_FLOOR).all(), (delta.max(), int(delta.argmax()))
    print("%s: fast fp64 vs oracle max|delta| = %.3e (%d of %d rows on the fast path)" % (name, delta.max(), ok.sum(), len(ok)))


def test_fast_fp64_price_matches_oracle_on_random_inputs(checker):
    s, k, r, v, t, o = inputgen_like(200000, seed=5, dtype=np.float64)
    ref = oracle_lib.price_map(s, k, r, v, t, o, 8)
    got, ok = _fast_prices(checker, s, k, r, v, t, o)
    assert ok.all()
    delta = np.abs(got - ref)
    assert (delta <= FP64_REL_TOL * np.abs(ref) + FP64_ABS_FLOOR).all()
    assert delta.max() < 1e-12


def test_degenerate_inputs_are_flagged_for_the_ieee_path(checker):
    # t = 0; v = 0; s = 0; s denormal-ish; v = 1e-160 (shared reciprocal overflows); r t = -800 / +800 (exp range);
    # |d1| > 37 (exp(-d1^2/2) underflows: the identity for the second exponential must not be used)
    s = [100.0, 100.0, 0.0, 1e-300, 100.0, 100.0, 100.0, 100.0]
    k = [100.0, 100.0, 90.0, 100.0, 90.0, 90.0, 90.0, 10.0]
    r = [0.05, 0.05, 0.05, 0.05, 0.05, -800.0, 800.0, 0.05]
    v = [0.2, 0.0, 0.2, 0.2, 1e-160, 0.2, 0.2, 0.05]
    t = [0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    got, ok = _fast_prices(checker, s, k, r, v, t, [0] * 8)
    assert ok.tolist() == [0] * 8


def test_inputgen_range_never_leaves_the_fast_path(checker):
    # corners of the PARSEC inputgen domain (largest |d1| = 32 at strike/spot 0.7, v = t = 0.05): all stay on the fast path
    s = [120.0, 120.0, 20.0, 20.0, 120.0, 20.0]
    k = [84.0, 156.0, 14.0, 26.0, 120.0, 20.0]
    v = [0.05, 0.05, 0.05, 0.05, 0.65, 0.65]
    t = [0.05, 0.05, 0.05, 0.05, 1.0, 1.0]
    for o in (0, 1):
        got, ok = _fast_prices(checker, s, k, [0.1] * 6, v, t, [o] * 6)
        assert ok.all()
        ref = oracle_lib.price_map(*[np.array(a, np.float64) for a in (s, k, [0.1] * 6, v, t)], np.full(6, o, np.int32), 8)
        assert (np.abs(got - ref) <= FP64_REL_TOL * np.abs(ref) + FP64_ABS_FLOOR).all(), (got, ref)
