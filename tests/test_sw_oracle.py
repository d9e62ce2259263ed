"""The swaptions oracle (oracle/sw_oracle.c) against outputs of the reference itself.

Bit-exact at the precision the reference prints: the "%.10lf" lines the restatement produces must equal, byte for
byte, what the reference's own HJM_Securities.cpp + HJM_Swaption_Blocking.cpp (compiled unmodified into
oracle/_ref/sw_ref_*) printed for the same command line (tests/golden/sw_*.json, made by
tests/golden/make_sw_golden.py) and -- where oracle/_ref is present -- what they print right now.
The PARSEC-owned leaves (RanUnif, CumNormalInv, ...) are absent from the reference tree and restated in
oracle/sw_absent/: for them parity is UNPINNED (oracle/sw_absent/HJM_type.h); the known-answer tests below pin
this repo's three restatements of them (C leaf, Python mirror, device code) to each other, not to PARSEC.
"""
import glob
import json
import math
import os

import numpy as np
import pytest

import sw_oracle_lib as so
from conftest import GOLDEN
from p3arsec_b200 import swaptions as sw

SW_CASES = sorted(os.path.basename(p)[3:-5] for p in glob.glob(os.path.join(GOLDEN, "sw_*.json")))


def _oracle_lines(a):
    seed = 1979 if a["sd"] is None else a["sd"]
    s, p, y, f = so.portfolio(a["ns"], seed)
    m, e = so.price_map(p, y, f, s, a["sm"])
    return so.format_lines(m, e)


def test_goldens_present():
    assert set(SW_CASES) >= {"simsmall16", "ragged7", "single1", "two_trials", "seeds5", "medium32"}


@pytest.mark.parametrize("name", SW_CASES)
def test_oracle_matches_reference_output_byte_for_byte(name):
    gold = json.load(open(os.path.join(GOLDEN, "sw_%s.json" % name)))
    assert _oracle_lines(gold["args"]) == gold["lines"]


@pytest.mark.skipif(not so.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
@pytest.mark.parametrize("binary,nt", [("sw_ref_serial", 1), ("sw_ref_ff", 4)])
def test_oracle_matches_reference_run_now(binary, nt):
    if not so.have_ref(binary):
        pytest.skip(binary + " not built")
    rng = np.random.RandomState(20261017)
    for _ in range(3):
        a = dict(ns=int(rng.randint(4, 12)), sm=int(rng.randint(1, 3000)), sd=int(rng.randint(1, 2**31 - 2)))
        _, stderr, roi = so.run_ref(a["ns"], a["sm"], nt, a["sd"], binary)
        assert roi is not None and roi >= 0
        assert [l for l in stderr.splitlines() if l.startswith("Swaption")] == _oracle_lines(a)


def test_portfolio_python_mirror_equals_oracle():
    # p3arsec_b200.swaptions.make_portfolio restates HJM_Securities.cpp:198,276-296 for the Python callers
    for n, seed in ((1, 1979), (16, 1979), (128, 1979), (9, 42), (33, 2147483646)):
        s1, p1, y1, f1 = so.portfolio(n, seed)
        s2, p2, y2, f2 = sw.make_portfolio(n, seed)
        assert s1 == s2
        assert p1.tobytes() == p2.tobytes() and y1.tobytes() == y2.tobytes() and f1.tobytes() == f2.tobytes()
    s, p, y, f = so.portfolio(128, 1979)
    assert s == 2004984073                                  # (long)(2147483647 * RanUnif(1979))
    assert 5.0 <= p["dYears"].min() and p["dYears"].max() <= 19.75      # "5 to 20 years in 3 month intervals"
    assert 0.1 - 1e-12 <= p["dStrike"].min() and p["dStrike"].max() <= 4.9 + 1e-9
    assert np.all(p["dCompounding"] == 0) and np.all(p["dMaturity"] == 1.0) and np.all(p["dTenor"] == 2.0)


def test_ranunif_counter_semantics():
    # state advances by one per call; the draw is a pure function of the counter
    u1, c1 = so.ran_unif(1979)
    assert c1 == 1980
    assert u1 == sw.ran_unif([1979])
    for c in (0, 1, 2, 1418, 127773, 2147483646, 2147483647, 2147483648, 10**12 + 7):
        st = [c]
        assert so.ran_unif(c)[0] == sw.ran_unif(st) and st[0] == c + 1
        assert 0.0 <= so.ran_unif(c)[0] < 1.0
    assert so.ran_unif(0)[0] == 0.0 and so.ran_unif(2147483647)[0] == 0.0   # counters that are multiples of 2^31 - 1 draw 0


def test_cumnormalinv_shape():
    assert so.cum_normal_inv(0.5) == 0.0
    for u in (0.08, 0.1, 0.3, 0.45, 0.7, 0.92, 0.999, 1e-6):
        z = so.cum_normal_inv(u)
        assert abs(z + so.cum_normal_inv(1.0 - u)) < 1e-9   # antisymmetric up to the rounding of 1 - u
        # Moro's approximation is good to ~3e-9 against the true quantile
        assert abs(0.5 * math.erfc(-z / math.sqrt(2)) - u) < 1e-8 * max(u, 1e-3)
    assert so.cum_normal_inv(0.0) == -math.inf              # log(-log(0)) = +inf, negated on the lower side


def test_block_rounding_of_trials():
    # HJM_Swaption_Blocking.cpp:156 simulates whole blocks: 1..16 trials all simulate 16 and differ only in the divisor
    s, p, y, f = so.portfolio(3, 1979)
    m16, _ = so.price_map(p, y, f, s, 16)
    for t in (1, 5, 15):
        m, _ = so.price_map(p, y, f, s, t)
        np.testing.assert_allclose(m * t, m16 * 16, rtol=1e-15)
    # and the block size is pure blocking: 16 trials in blocks of 1, 2, 4, 8, 16 give bit-identical sums
    for bs in (1, 2, 4, 8):
        m, e = so.price_map(p, y, f, s, 16, block_size=bs)
        assert m.tobytes() == m16.tobytes()


def test_invalid_time_indices_are_refused():
    s, p, y, f = so.portfolio(1, 1979)
    p["dMaturity"] = 50.0      # swap start beyond the path
    with pytest.raises(ValueError):
        so.price_map(p, y, f, s, 16)
