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
@pytest.mark.parametrize("binary,nt", [("sw_ref_serial", 1), ("sw_ref_ff", 4), ("sw_ref_skepu", 2)])
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


# ---- host-checkable pieces of the device arithmetic (p3arsec_b200/csrc/sw_kernels.cuh) -------------------------------
def _kernel_constants():
    import re
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "p3arsec_b200", "csrc", "sw_kernels.cuh")).read()
    m = re.search(r"constexpr uint32_t S_LO = (\d+)u, S_HI = (\d+)u;", src)
    return int(m.group(1)), int(m.group(2))


def test_central_branch_integer_range_equals_the_double_test():
    """The kernel decides CumNormalInv's branch with S_LO <= s <= S_HI on the 31-bit draw s; the reference decides it
    with fabs(s * 4.656612875e-10 - 0.5) < 0.42 in doubles (two roundings).  Same decision for every s."""
    s_lo, s_hi = _kernel_constants()
    c = 4.656612875e-10

    def central(s):
        return abs(s * c - 0.5) < 0.42

    for edge in (s_lo, s_hi):
        for s in range(edge - 5000, edge + 5000):
            assert central(s) == (s_lo <= s <= s_hi), s
    rng = np.random.RandomState(3)
    for s in rng.randint(0, 2**31 - 1, 20000):
        assert central(int(s)) == (s_lo <= int(s) <= s_hi)
    assert not central(0) and not central(2**31 - 2) and central(2**30)


def test_residue_arithmetic_equals_ranunif():
    """The kernel keeps x = ctr * 1513517 mod (2^31 - 1) as a 32-bit residue and computes 16807 x mod (2^31 - 1) with a
    Mersenne fold instead of Schrage's split; phase A even leaves x and the folded sum unreduced.  Same draws."""
    M = 2**31 - 1

    def fold(p):
        return (p & M) + (p >> 31)

    rng = np.random.RandomState(5)
    ctrs = [0, 1, M - 1, M, M + 1, 2 * M, 2**40 - 1] + [int(v) for v in rng.randint(0, 2**40, 3000, dtype=np.int64)]
    for ctr in ctrs:
        x = fold(fold(ctr * 1513517))
        x = x - M if x >= M else x
        assert x == (ctr * 1513517) % M
        for k in (0, 1, 29):
            xu = x + k * 1513517                       # unreduced, as in phase A
            assert xu < 2**32
            s = fold(xu * 16807)
            true_s = int(round(so.ran_unif(ctr + k)[0] / 4.656612875e-10))
            assert so.ran_unif(ctr + k)[0] == true_s * 4.656612875e-10
            assert s % M == true_s and s < M + 2**16
            if s != true_s:                            # unreduced sum: stands for a tiny draw and must fail the range check
                s_lo, s_hi = _kernel_constants()
                assert s > s_hi and true_s < s_lo


def test_tail_table_matches_long_double_libm(tmp_path):
    """p3arsec_b200/csrc/sw_tail.h (host-or-device): one logarithm + the composite table against P8(log(-log r)) in
    long double, over 4 M random tail draws and the 200 000 smallest draws exhaustively."""
    import subprocess
    root = os.path.join(os.path.dirname(GOLDEN), "..")
    exe = str(tmp_path / "sw_tail_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tools", "sw_tail_host_check.cpp"), "-lm"], check=True)
    out = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    assert float(out["max_rel"]) < 5e-16 and float(out["max_ulp"]) <= 2.5


# ---- the fast kernels' per-trial arithmetic, built for the host from the SAME source (sw_kernels.cuh is host-or-device) ----
@pytest.fixture(scope="module")
def fast_host(tmp_path_factory):
    import subprocess
    root = os.path.join(os.path.dirname(GOLDEN), "..")
    exe = str(tmp_path_factory.mktemp("swfast") / "sw_fast_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-I/usr/local/cuda/include", "-o", exe,
                    os.path.join(root, "tools", "sw_fast_host_check.cpp"), "-lm"], check=True)

    def run(p, y, f, seed, trials, lean, source, block_size=16):
        lines = ["%d %d %d %d %d %d" % (len(p), trials, block_size, seed, lean, source)]
        for i in range(len(p)):
            lines.append(" ".join("%.17g" % p[k][i] for k in ("dStrike", "dCompounding", "dMaturity", "dTenor", "dPaymentInterval", "dYears")))
            lines.append(" ".join("%.17g" % v for v in np.asarray(y[i]).ravel()))
            lines.append(" ".join("%.17g" % v for v in np.asarray(f[i]).ravel()))
        out = subprocess.run([exe], input="\n".join(lines), capture_output=True, text=True, check=True).stdout
        rows = np.array([[float(x) for x in l.split()] for l in out.splitlines() if l.strip()])
        return rows[:, 0], rows[:, 1], rows[:, 2].astype(int)
    return run


def _close(mean, omean, rtol=1e-12):
    assert np.array_equal(np.isnan(mean), np.isnan(omean))
    m = np.isfinite(omean)
    assert np.array_equal(mean[~m & ~np.isnan(omean)], omean[~m & ~np.isnan(omean)])
    assert (np.abs(mean[m] - omean[m]) <= rtol * np.abs(omean[m]) + 1e-15).all(), np.abs(mean[m] - omean[m]).max()


@pytest.mark.parametrize("lean", [0, 1])
@pytest.mark.parametrize("source", [0, 1])
def test_fast_arithmetic_on_the_host_matches_the_oracle(fast_host, lean, source):
    """normals() + path_and_payoff() + exp_core + the tail table exactly as the GPU kernels compile them (tables from the
    shared-memory block or from the constant-bank record), summed in trial order: within 1e-12 of the oracle's price."""
    seed, p, y, f = sw.make_portfolio(12)
    for trials in (4096, 1003):
        omean, _ = so.price_map(p, y, f, seed, trials)
        s, s2, fb = fast_host(p, y, f, seed, trials, lean, source)
        assert fb.sum() == 0
        _close(s / trials, omean)


def test_fast_arithmetic_every_start_index_on_the_host(fast_host):
    n = 11
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"], p["dStrike"], p["dPaymentInterval"] = 11.0, 0.09, 1.0
    p["dMaturity"] = np.arange(n, dtype=np.float64)
    p["dTenor"] = np.minimum(3.0, 10.0 - np.arange(n))
    y = np.tile(0.05 + 0.004 * np.arange(11), (n, 1))
    f = np.tile(sw.FACTOR_TABLE[None] * 1.5, (n, 1, 1))
    omean, _ = so.price_map(p, y, f, 4242, 2000)
    for lean in (0, 1):
        s, _, fb = fast_host(p, y, f, 4242, 2000, lean, 1)
        assert fb.sum() == 0
        _close(s / 2000, omean)


def test_fast_arithmetic_falls_back_when_a_draw_is_zero(fast_host):
    """A counter that is a multiple of 2^31 - 1 draws 0 -> z = -inf -> an exponential leaves the fast range: the trial is
    redone by generic_trial() (the reference's operation order), and the non-finite sums come out like the oracle's."""
    _, p, y, f = sw.make_portfolio(4)
    seed = 2147483647 - 40
    omean, _ = so.price_map(p, y, f, seed, 64)
    s, _, fb = fast_host(p, y, f, seed, 64, 0, 0)
    assert fb.sum() >= 1
    _close(s / 64, omean)


def test_fast_arithmetic_hands_very_high_rates_to_the_generic_trial(fast_host):
    """exp_core is used for |rate * dt| < 2 (one range-reduction step, sw_kernels.cuh EXP_HI_LIMIT); a trial with a larger
    argument is redone by generic_trial().  Yield curves of 60-260 % with dt = 1 year put the portfolio on both sides of the
    limit: some swaptions never fall back, some always do, and every price still equals the oracle's."""
    n = 6
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"], p["dStrike"], p["dPaymentInterval"], p["dMaturity"], p["dTenor"] = 11.0, 0.9, 1.0, 1.0, 3.0
    y = np.tile(np.array([0.6, 1.0, 1.4, 1.8, 2.2, 2.6])[:, None], (1, 11)) + 0.01 * np.arange(11)[None]
    f = np.tile(sw.FACTOR_TABLE[None] * 4.0, (n, 1, 1))
    trials = 512
    omean, _ = so.price_map(p, y, f, 99, trials)
    for lean, source in ((0, 0), (0, 1), (1, 1)):
        s, _, fb = fast_host(p, y, f, 99, trials, lean, source)
        assert fb[0] == 0 and fb[-1] == trials, fb
        _close(s / trials, omean)


@pytest.mark.parametrize("rng_seed", [1, 2, 3])
def test_fast_arithmetic_random_portfolios_on_the_host(fast_host, rng_seed):
    """Random contracts beyond what the reference driver creates (compounding conventions, maturities, tenors, payment
    intervals, per-swaption yield curves and factor tables): the host build of the fast arithmetic against the oracle,
    full and lean.  Contracts whose time indices leave the path are refused by both (the checker prints 'invalid')."""
    rng = np.random.RandomState(rng_seed)
    n = 16
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"] = 5.0 + rng.randint(0, 60, n) * 0.25
    p["dStrike"] = 0.01 + rng.randint(0, 40, n) * 0.005
    p["dCompounding"] = rng.choice([0.0, 0.5, 1.0], n)
    p["dMaturity"] = rng.choice([0.5, 1.0, 2.0, 3.0], n)
    p["dTenor"] = rng.choice([1.0, 2.0, 3.0, 5.0], n)
    p["dPaymentInterval"] = rng.choice([0.5, 1.0, 2.0], n)
    y = 0.02 + 0.1 * rng.rand(n, 1) + np.cumsum(0.004 * rng.rand(n, 11), axis=1)
    f = sw.FACTOR_TABLE[None] * (0.25 + 2.0 * rng.rand(n, 3, 1)) * np.where(rng.rand(n, 3, 10) < 0.1, -1.0, 1.0)
    keep = []
    for i in range(n):
        try:
            so.price_map(p[i:i + 1], y[i:i + 1], f[i:i + 1], 1, 16)
            keep.append(i)
        except ValueError:
            pass
    assert len(keep) >= 6
    p, y, f = p[keep], y[keep], f[keep]
    seed = int(rng.randint(0, 2**31 - 1))
    omean, _ = so.price_map(p, y, f, seed, 1500)
    for lean in (0, 1):
        s, _, fb = fast_host(p, y, f, seed, 1500, lean, lean)
        assert fb.sum() == 0
        _close(s / 1500, omean)
