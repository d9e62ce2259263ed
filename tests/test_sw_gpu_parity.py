"""-m gpu: the swaptions Map on a B200 (through the C ABI, include/sw_gpu.h) against the oracle and the committed
outputs of the reference.

Tolerance (BASELINE.json north_star: <= 1e-9 relative for fp64), written out:
  mean price   |d| <= 1e-9 |ref| + 1e-12
  std error    compared as the variance it is the root of: var = n stderr^2, |d var| <= 1e-9 (mean^2 + var) + 1e-24.
               The reference forms var as sumsq - sum^2/n (HJM_Swaption_Blocking.cpp:213): a cancellation that
               magnifies any difference in the two sums (summation order included) by mean^2/var, and that goes
               negative -- std error NaN -- by rounding alone when every trial pays the same.
NaN / inf prices must sit where the reference has them.  The bound is loose on purpose: what the kernels actually reach is
printed by test_report_measured_distance and recorded in DESIGN.md.
"""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

import sw_oracle_lib as so
from conftest import GOLDEN
from p3arsec_b200 import swaptions as sw

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p3arsec_b200", "bin", "swaptions_gpu")
SW_CASES = sorted(os.path.basename(p)[3:-5] for p in glob.glob(os.path.join(GOLDEN, "sw_*.json")))
MODES = [("fast", 0), ("lean", sw.FLAG_LEAN), ("ieee", sw.FLAG_IEEE), ("ieee_lean", sw.FLAG_IEEE | sw.FLAG_LEAN)]


def assert_parity(mean, err, rmean, rerr, trials, what=""):
    mean, err, rmean, rerr = (np.asarray(a, np.float64) for a in (mean, err, rmean, rerr))
    assert np.array_equal(np.isnan(rmean), np.isnan(mean)), what + ": NaN prices differ"
    m = np.isfinite(rmean)
    assert np.array_equal(mean[~m & ~np.isnan(rmean)], rmean[~m & ~np.isnan(rmean)]), what + ": infinities differ"
    d = np.abs(mean[m] - rmean[m])
    assert (d <= 1e-9 * np.abs(rmean[m]) + 1e-12).all(), "%s: price off by %.3e (rel %.3e)" % (what, d.max(), (d / np.maximum(np.abs(rmean[m]), 1e-300)).max())
    # std error, compared where it is formed: var = n stderr^2 = (sumsq - sum^2/n)/(n-1).  A NaN std error is the square
    # root of a negative variance estimate (extra trials of a ragged last block, or plain rounding when every trial pays
    # the same); it counts as var <= 0, so "NaN vs 0" and "NaN vs 1e-12" are agreements when the price is large.
    n = max(trials, 1)
    with np.errstate(invalid="ignore"):
        var_g = np.where(np.isnan(err[m]), 0.0, n * err[m] ** 2)
        var_r = np.where(np.isnan(rerr[m]), 0.0, n * rerr[m] ** 2)
    fin = np.isfinite(var_r)
    assert np.array_equal(np.isfinite(var_g), fin), what + ": infinite std errors differ"
    dv = np.abs(var_g[fin] - var_r[fin])
    assert (dv <= 1e-9 * (rmean[m][fin] ** 2 + var_r[fin]) + 1e-24).all(), "%s: variance off by %.3e" % (what, dv.max())
    return float((d / np.maximum(np.abs(rmean[m]), 1e-300)).max()) if d.size else 0.0


def gpu_price(p, y, f, seed, trials, flags=0, block_size=16, num_gpus=1, iN=11, iFactors=3, geometry=None):
    with sw.SwaptionsGPU(max(len(p), 1), num_gpus=num_gpus, iN=iN, iFactors=iFactors) as ctx:
        if geometry:
            ctx.set_geometry(*geometry)
        mean, err = ctx.price(p, y, f, seed, trials, block_size, flags)
        return mean, err, ctx.timing(), ctx.shards()


@pytest.mark.parametrize("mode,flags", MODES)
@pytest.mark.parametrize("name", SW_CASES)
def test_reference_command_lines_match_committed_reference_output(name, mode, flags):
    """The reference's own answers (tests/golden/sw_*.json) for its own command lines."""
    gold = json.load(open(os.path.join(GOLDEN, "sw_%s.json" % name)))
    a = gold["args"]
    seed, p, y, f = sw.make_portfolio(a["ns"], 1979 if a["sd"] is None else a["sd"])
    mean, err, tm, _ = gpu_price(p, y, f, seed, a["sm"], flags)
    ref = so.parse_ref_output("\n".join(gold["lines"]))
    rmean = np.array([float(x[1]) for x in ref])
    rerr = np.array([float(x[2]) for x in ref])
    # the committed lines carry 10 decimals: compare at that resolution, then tighter against the oracle
    assert np.nanmax(np.abs(mean - rmean)) <= 0.5e-10 + 1e-9 * np.abs(rmean).max()
    ok = ~np.isnan(rerr) & ~np.isnan(err)
    if ok.any():
        assert np.abs(err[ok] - rerr[ok]).max() <= 1e-9
    omean, oerr = so.price_map(p, y, f, seed, a["sm"])
    assert_parity(mean, err, omean, oerr, a["sm"], "%s/%s" % (name, mode))
    assert tm["kernel_launches"] == 2 and tm["trials_simulated"] == a["ns"] * ((a["sm"] + 15) // 16) * 16   # few trials: one batched launch + finalize


@pytest.mark.parametrize("mode,flags", MODES)
def test_random_portfolios_against_oracle(mode, flags):
    """Seeded portfolios beyond what the reference driver creates: compounding conventions, maturities and tenors that
    move the swap start / length, yields and factor tables of their own per swaption."""
    rng = np.random.RandomState(7)
    n = 24
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"] = 5.0 + rng.randint(0, 60, n) * 0.25
    p["dStrike"] = 0.01 + rng.randint(0, 30, n) * 0.005
    p["dCompounding"] = rng.choice([0.0, 0.5, 1.0], n)
    p["dMaturity"] = rng.choice([0.5, 1.0, 2.0], n)
    p["dTenor"] = rng.choice([1.0, 2.0, 3.0], n)
    p["dPaymentInterval"] = rng.choice([0.5, 1.0], n)
    y = 0.02 + 0.1 * rng.rand(n, 1) + np.cumsum(0.004 * rng.rand(n, 11), axis=1)
    f = sw.FACTOR_TABLE[None] * (0.5 + rng.rand(n, 3, 1))
    # keep only what the reference could index (the oracle refuses the rest; the ABI returns SW_GPU_ERR_INVALID for them)
    keep = []
    for i in range(n):
        try:
            so.price_map(p[i:i + 1], y[i:i + 1], f[i:i + 1], 1, 16)
            keep.append(i)
        except ValueError:
            pass
    assert len(keep) >= 8
    p, y, f = p[keep], y[keep], f[keep]
    for trials, seed in ((4096, 123), (1000, 2004984073)):
        mean, err, _, _ = gpu_price(p, y, f, seed, trials, flags)
        omean, oerr = so.price_map(p, y, f, seed, trials)
        assert (omean > 0).any()
        assert_parity(mean, err, omean, oerr, trials, "random/%s/%d" % (mode, trials))


@pytest.mark.parametrize("iN,iFactors", [(6, 2), (16, 4), (32, 8), (2, 1), (11, 1)])
def test_other_path_shapes_use_the_generic_kernel(iN, iFactors):
    rng = np.random.RandomState(iN * 10 + iFactors)
    n = 5
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"] = iN * 0.5                      # ddelt = 0.5
    p["dStrike"] = 0.02 + 0.01 * rng.rand(n)
    p["dMaturity"] = 0.5 * min(2, iN - 2) if iN > 2 else 0.0
    p["dTenor"] = 0.5 * max(0, min(4, iN - 2 - min(2, iN - 2)))
    p["dPaymentInterval"] = 0.5
    y = 0.03 + np.cumsum(0.003 * rng.rand(n, iN), axis=1)
    f = 0.01 * rng.rand(n, iFactors, iN - 1)
    omean, oerr = so.price_map(p, y, f, 99, 2048, iN=iN, iFactors=iFactors)
    for flags in (0, sw.FLAG_IEEE):
        mean, err, _, _ = gpu_price(p, y, f, 99, 2048, flags, iN=iN, iFactors=iFactors)
        assert_parity(mean, err, omean, oerr, 2048, "shape %dx%d" % (iN, iFactors))


def test_block_size_and_ragged_trials():
    seed, p, y, f = sw.make_portfolio(6)
    for trials, bs in ((0, 16), (1, 16), (15, 16), (17, 16), (100, 7), (100, 1), (4097, 64)):  # 0 trials: 0/0 = NaN, like the reference
        mean, err, tm, _ = gpu_price(p, y, f, seed, trials, 0, block_size=bs)
        omean, oerr = so.price_map(p, y, f, seed, trials, block_size=bs)
        assert tm["trials_simulated"] == 6 * ((trials + bs - 1) // bs) * bs
        assert_parity(mean, err, omean, oerr, trials, "trials=%d bs=%d" % (trials, bs))


def test_counters_crossing_the_generator_modulus():
    """Counters that are multiples of 2^31 - 1 draw exactly 0, whose normal is -inf (log(-log(0))): the reference then
    carries inf / NaN through the path.  Seeds just below the modulus make a trial hit it."""
    _, p, y, f = sw.make_portfolio(4)
    for seed in (2147483647 - 40, 2147483647 - 1, 2147483647, 2 * 2147483647 - 100):
        omean, oerr = so.price_map(p, y, f, seed, 64)
        # (not the lean flavour: it skips the draws and discount factors the price does not depend on, which is only
        # equivalent while every intermediate value is finite -- include/sw_gpu.h)
        for mode, flags in (("fast", 0), ("ieee", sw.FLAG_IEEE)):
            mean, err, _, _ = gpu_price(p, y, f, seed, 64, flags)
            assert_parity(mean, err, omean, oerr, 64, "seed %d/%s" % (seed, mode))


def test_very_high_rates_fall_back_to_the_generic_trial():
    """The fast kernels' exponential is used for |rate * dt| < 2 (sw_kernels.cuh EXP_HI_LIMIT); trials beyond are redone by
    generic_trial() inside the same launch.  Yield curves of 60-260 % put the portfolio on both sides of the limit
    (tests/test_sw_oracle.py checks on the host build that some swaptions never and some always fall back)."""
    n = 6
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"], p["dStrike"], p["dPaymentInterval"], p["dMaturity"], p["dTenor"] = 11.0, 0.9, 1.0, 1.0, 3.0
    y = np.tile(np.array([0.6, 1.0, 1.4, 1.8, 2.2, 2.6])[:, None], (1, 11)) + 0.01 * np.arange(11)[None]
    f = np.tile(sw.FACTOR_TABLE[None] * 4.0, (n, 1, 1))
    omean, oerr = so.price_map(p, y, f, 99, 2048)
    for mode, flags in MODES[:2]:
        mean, err, _, _ = gpu_price(p, y, f, 99, 2048, flags)
        assert_parity(mean, err, omean, oerr, 2048, "high rates/" + mode)
    # and through the one-launch-per-swaption kernel
    mean, err, _, _ = gpu_price(p[:2], y[:2], f[:2], 99, 300000, 0)
    bmean, berr, _, _ = gpu_price(p[:2], y[:2], f[:2], 99, 300000, sw.FLAG_BATCHED)
    assert_parity(mean, err, bmean, berr, 300000, "high rates/one-swaption vs batched")


def test_negative_and_huge_seeds_fall_back_to_literal_arithmetic():
    _, p, y, f = sw.make_portfolio(3)
    for seed in (-5, -2147483647 * 3, 2**41 + 12345):
        omean, oerr = so.price_map(p, y, f, seed, 256)
        mean, err, _, _ = gpu_price(p, y, f, seed, 256, 0)
        assert_parity(mean, err, omean, oerr, 256, "seed %d" % seed)


def test_empty_and_invalid_inputs():
    seed, p, y, f = sw.make_portfolio(2)
    with sw.SwaptionsGPU(2) as ctx:
        mean, err = ctx.price(p[:0], y[:0], f[:0], seed, 128)
        assert mean.shape == (0,) and ctx.shards() == []
        bad = p.copy()
        bad["dMaturity"][1] = 50.0            # swap start beyond the 11-point path: the reference would index off its vectors
        with pytest.raises(sw.SwGpuError) as ei:
            ctx.price(bad, y, f, seed, 128)
        assert ei.value.status == -1 and "swaption 1" in str(ei.value)
        if sw.device_count() >= 2:            # an invalid swaption in the second device's shard: the first one has already started
            with sw.SwaptionsGPU(2, num_gpus=2) as two:
                with pytest.raises(sw.SwGpuError):
                    two.price(bad, y, f, seed, 128)
                m2, e2 = two.price(p, y, f, seed, 128)
                m1, e1 = ctx.price(p, y, f, seed, 128)
                assert m2.tobytes() == m1.tobytes()
        with pytest.raises(sw.SwGpuError):
            ctx.price(p, y, f, seed, 128, block_size=0)
        with pytest.raises(sw.SwGpuError):
            ctx.price(p, y, f, seed, 128, flags=64)
        with pytest.raises(sw.SwGpuError):   # more swaptions than the context was made for
            ctx.price(np.concatenate([p, p]), np.concatenate([y, y]), np.concatenate([f, f]), seed, 128)
        # the context still works after refused calls
        mean, err = ctx.price(p, y, f, seed, 128)
        omean, oerr = so.price_map(p, y, f, seed, 128)
        assert_parity(mean, err, omean, oerr, 128, "after errors")


def test_deterministic_and_geometry_independent():
    seed, p, y, f = sw.make_portfolio(8)
    a = gpu_price(p, y, f, seed, 20000, 0)
    b = gpu_price(p, y, f, seed, 20000, 0)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()   # no atomics on the data path
    for geo in ((1, 1), (2, 3), (4, 64)):
        c = gpu_price(p, y, f, seed, 20000, 0, geometry=geo)
        np.testing.assert_allclose(c[0], a[0], rtol=1e-13)
        np.testing.assert_allclose(c[1], a[1], rtol=1e-9)


def test_native_size_properties():
    """PARSEC native (-ns 128 -sm 1000000) is far beyond what the oracle finishes in seconds; size-independent checks:
    the lean kernel (dead work removed) equals the full one, a prefix of the portfolio prices identically, the 1M-trial
    price sits inside the oracle's 40k-trial confidence interval, and the standard error shrinks like 1/sqrt(n)."""
    seed, p, y, f = sw.make_portfolio(128)
    full = gpu_price(p, y, f, seed, 1000000, 0)
    lean = gpu_price(p, y, f, seed, 1000000, sw.FLAG_LEAN)
    np.testing.assert_allclose(lean[0], full[0], rtol=1e-11)
    assert full[2]["trials_simulated"] == 128 * 1000000
    head = gpu_price(p[:5], y[:5], f[:5], seed, 1000000, 0)
    np.testing.assert_allclose(head[0], full[0][:5], rtol=1e-13)
    omean, oerr = so.price_map(p[:16], y[:16], f[:16], seed, 40000)
    z = np.abs(full[0][:16] - omean) / np.maximum(oerr, 1e-300)
    assert (z[oerr > 0] < 6).all()
    small = gpu_price(p[:16], y[:16], f[:16], seed, 40000, 0)
    ok = (small[1] > 0) & np.isfinite(small[1])
    ratio = full[1][:16][ok] / small[1][ok]
    np.testing.assert_allclose(ratio, np.sqrt(40000 / 1000000), rtol=0.05)


@pytest.mark.multigpu
def test_shards_over_devices():
    n_dev = sw.device_count()
    if n_dev < 2:
        msg = "SKIPPED (LOUD): test_shards_over_devices needs >= 2 GPUs in ONE process; this box shows %d" % n_dev
        print("\n" + "!" * 100 + "\n" + msg + "\n" + "!" * 100)
        pytest.skip(msg)
    seed, p, y, f = sw.make_portfolio(13)
    one = gpu_price(p, y, f, seed, 8192, 0)
    for g in range(2, min(n_dev, 8) + 1):
        many = gpu_price(p, y, f, seed, 8192, 0, num_gpus=g)
        shards = many[3]
        assert len(shards) == g and sum(c for _, _, c in shards) == 13
        assert [f0 for _, f0, _ in shards] == list(np.cumsum([0] + [c for _, _, c in shards])[:-1])
        assert many[0].tobytes() == one[0].tobytes() and many[1].tobytes() == one[1].tobytes()


def test_driver_binary_prints_the_reference_lines():
    gold = json.load(open(os.path.join(GOLDEN, "sw_simsmall16.json")))
    cp = subprocess.run([EXE, "-ns", "16", "-sm", "10000", "-nt", "1"], capture_output=True, text=True)
    assert cp.returncode == 0, cp.stderr
    assert "Number of Simulations: 10000,  Number of threads: 1 Number of swaptions: 16" in cp.stdout
    assert "roi.time|" in cp.stdout
    got = so.parse_ref_output(cp.stderr)
    ref = so.parse_ref_output("\n".join(gold["lines"]))
    assert [g[0] for g in got] == list(range(16))
    for g, r in zip(got, ref):
        assert abs(float(g[1]) - float(r[1])) <= 1e-9 * max(1.0, abs(float(r[1])))
        assert abs(float(g[2]) - float(r[2])) <= 1e-9
    # at the printed resolution the lines are expected to be identical
    same = sum(1 for g, r in zip(got, ref) if g == r)
    print("driver lines identical to the reference's at 10 decimals: %d / 16" % same)


def test_report_measured_distance(capsys):
    """Not a bound: prints what each kernel flavour reaches against the oracle (recorded in DESIGN.md)."""
    seed, p, y, f = sw.make_portfolio(32)
    omean, oerr = so.price_map(p, y, f, seed, 20000)
    with capsys.disabled():
        for mode, flags in MODES:
            mean, err, tm, _ = gpu_price(p, y, f, seed, 20000, flags)
            rel = assert_parity(mean, err, omean, oerr, 20000, mode)
            ok = (oerr > 0) & np.isfinite(err)
            print("\n[sw parity] %-9s 32 x 20000: max rel |d price| = %.3e, max rel |d stderr| = %.3e, bit-identical prices %d/32, kernels %.3f ms"
                  % (mode, rel, float((np.abs(err[ok] - oerr[ok]) / oerr[ok]).max()), int((mean == omean).sum()), tm["roi_ms"]), end="")
        print()


def test_reference_driver_with_cuda_map():
    """oracle/_ref/sw_ref_cuda = the reference's own HJM_Securities.cpp with integration/HJM_Securities.cpp.enable_cuda.patch
    (its argument handling, RanUnif-driven portfolio set-up, ROI markers and result printing; only the Map is
    sw_gpu_price) against the committed output of the unmodified CPU builds."""
    exe = os.path.join(ROOT, "oracle", "_ref", "sw_ref_cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sw_ref_cuda not built (needs /root/reference at build time)")
    for name in ("simsmall16", "ragged7", "seeds5"):
        gold = json.load(open(os.path.join(GOLDEN, "sw_%s.json" % name)))
        a = gold["args"]
        cmd = [exe, "-ns", str(a["ns"]), "-sm", str(a["sm"]), "-nt", "1"] + (["-sd", str(a["sd"])] if a["sd"] is not None else [])
        cp = subprocess.run(cmd, capture_output=True, text=True)
        assert cp.returncode == 0, cp.stderr
        assert "roi.time|" in cp.stdout  # the hooks shim, at the reference's ROI markers
        got = so.parse_ref_output(cp.stderr)
        ref = so.parse_ref_output("\n".join(gold["lines"]))
        assert len(got) == len(ref) == a["ns"]
        for g, r in zip(got, ref):
            assert abs(float(g[1]) - float(r[1])) <= 1e-9 * max(1.0, abs(float(r[1])))
            if "nan" not in r[2] and "nan" not in g[2]:
                assert abs(float(g[2]) - float(r[2])) <= 1e-9


@pytest.mark.parametrize("mode,flags", MODES[:3])
def test_every_swap_start_index(mode, flags):
    """The fast kernel is specialised on iSwapStartTimeIndex = 1, 2, 3 and has a runtime-index version for the rest:
    one swaption per start index 0..10 (ddelt = 1 year, maturity = start years, tenor up to 3 years)."""
    n = 11
    p = np.zeros(n, dtype=sw.SWAPTION_DTYPE)
    p["dYears"] = 11.0
    p["dStrike"] = 0.09
    p["dMaturity"] = np.arange(n, dtype=np.float64)
    p["dTenor"] = np.minimum(3.0, 10.0 - np.arange(n))
    p["dPaymentInterval"] = 1.0
    y = np.tile(0.05 + 0.004 * np.arange(11), (n, 1))
    f = np.tile(sw.FACTOR_TABLE[None] * 1.5, (n, 1, 1))
    omean, oerr = so.price_map(p, y, f, 4242, 3000)
    assert (omean[:10] > 0).all() and omean[10] == 0.0   # start 10: nothing left of the path to pay on
    mean, err, _, _ = gpu_price(p, y, f, 4242, 3000, flags)
    assert_parity(mean, err, omean, oerr, 3000, "start index/" + mode)


def test_one_swaption_per_launch_path_against_oracle_and_batched_kernel():
    """From 262144 trials per swaption on, every swaption gets a launch of its own with its tables in the constant
    bank (sw_sim_one); SW_GPU_FLAG_BATCHED forces the single-launch kernel.  Both against the oracle, and each other."""
    seed, p, y, f = sw.make_portfolio(5)
    trials = 262144 + 48
    omean, oerr = so.price_map(p, y, f, seed, trials)
    for mode, flags in (("fast", 0), ("lean", sw.FLAG_LEAN)):
        one = gpu_price(p, y, f, seed, trials, flags)
        bat = gpu_price(p, y, f, seed, trials, flags | sw.FLAG_BATCHED)
        assert one[2]["kernel_launches"] == (5 + 1 if mode == "fast" else 2) and bat[2]["kernel_launches"] == 2  # lean always batches
        assert_parity(one[0], one[1], omean, oerr, trials, "one/" + mode)
        assert_parity(bat[0], bat[1], omean, oerr, trials, "batched/" + mode)
        np.testing.assert_allclose(one[0], bat[0], rtol=1e-13)
        if sw.device_count() >= 2:   # the per-swaption launches of two devices run at the same time
            two = gpu_price(p, y, f, seed, trials, flags, num_gpus=2)
            assert two[0].tobytes() == one[0].tobytes() and two[1].tobytes() == one[1].tobytes()
            assert two[2]["kernel_launches"] == (5 + 2 if mode == "fast" else 4)
    # a geometry that gives a launch several chunks per CTA, and a seed that makes a trial hit the generator's modulus
    hit = 2147483647 - 30 * 1000
    omean, oerr = so.price_map(p[:2], y[:2], f[:2], hit, trials)
    got = gpu_price(p[:2], y[:2], f[:2], hit, trials, 0, geometry=(2, 3))
    assert_parity(got[0], got[1], omean, oerr, trials, "one/modulus")
