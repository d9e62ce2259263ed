"""Parity of the CUDA Map with the reference (-m gpu; every call goes through the C ABI of libbs_gpu.so).

Three kinds of evidence:
  1. committed golden outputs of the unmodified reference binaries (tests/golden/*.ref_f32/f64.txt);
  2. the oracle restatement (bit-identical to the reference, tests/test_oracle.py) on seeded inputs of
     many sizes, including ragged ones;
  3. size-independent properties at BASELINE.json's full sizes (10M options): periodicity of the
     cyclic inputgen set, put-call parity, idempotence over NUM_RUNS, shard-independence.
Tolerances: fp32 max |delta| <= 1e-4 absolute; fp64 |delta| <= 1e-9*|ref| + 1e-12 (see gpu_util.py).
"""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN_CASES, golden_path
from gpu_util import FP32_ABS_TOL, assert_parity, gpu_prices, inputgen_like, magnitude_scale, oracle_prices
from p3arsec_b200 import host

pytestmark = pytest.mark.gpu

MATHS = [(host.MATH_IEEE, "ieee"), (host.MATH_FAST, "fast")]


def _golden_inputs(name, fp_bytes):
    d = host.load_options(golden_path(name, "in.txt"), fp_bytes)
    return (d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"]), d


def _golden_prices(name, sfx):
    n, toks = oracle_lib.read_prices_text(golden_path(name, "ref_%s.txt" % sfx))
    return np.array([float(t) for t in toks])


def test_library_reports_a_device():
    assert host.device_count() >= 1


@pytest.mark.parametrize("math,mname", MATHS)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp32_against_reference_golden(name, math, mname):
    inputs, d = _golden_inputs(name, 4)
    got, _, _ = gpu_prices(inputs, 4, num_runs=1, math=math)
    # flat 1e-4 for every row inside the inputgen range; only edge2k holds larger operands (see magnitude_scale)
    scale = magnitude_scale(inputs[0], inputs[1])
    assert name == "edge2k" or (scale <= 156.0 / 128.0).all()
    worst = assert_parity(got, _golden_prices(name, "f32"), 4, "%s/%s" % (name, mname), scale if name == "edge2k" else None)
    print("fp32 %-9s %-4s max|delta| = %.3e%s" % (name, mname, worst, " (per 128 of magnitude)" if name == "edge2k" else ""))


@pytest.mark.parametrize("math,mname", MATHS)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp64_against_reference_golden(name, math, mname):
    inputs, d = _golden_inputs(name, 8)
    got, _, _ = gpu_prices(inputs, 8, num_runs=1, math=math)
    ref = _golden_prices(name, "f64")
    worst = assert_parity(got, ref, 8, name)
    exact = float(np.mean(got == ref))
    print("fp64 %-9s %-4s max|delta| = %.3e, bit-identical on %.1f%% of rows" % (name, mname, worst, 100 * exact))


@pytest.mark.parametrize("math,mname", MATHS)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 37, 255, 256, 257, 1000, 4096, 65537, 1000003])
def test_fp32_against_oracle_sizes(n, math, mname):
    inputs = inputgen_like(n, seed=n)
    got, _, _ = gpu_prices(inputs, 4, num_runs=2, math=math)
    assert_parity(got, oracle_prices(inputs, 4), 4, "n=%d/%s" % (n, mname))


@pytest.mark.parametrize("math,mname", MATHS)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 37, 257, 4096, 65537, 1000003])
def test_fp64_against_oracle_sizes(n, math, mname):
    inputs = inputgen_like(n, seed=n, dtype=np.float64)
    got, _, _ = gpu_prices(inputs, 8, num_runs=2, math=math)
    assert_parity(got, oracle_prices(inputs, 8), 8, "n=%d/%s" % (n, mname))


def test_fp64_degenerate_inputs_follow_the_reference():
    # t = 0, v = 0, s = k with t = 0 ...: the fast path hands these to the IEEE-order path, which reproduces the
    # reference's inf/NaN arithmetic (intrinsic value, or NaN where the reference gives NaN).  Rows 7-9:
    # v = 1e-160 (the shared CNDF reciprocal's argument overflows), r t = -800 and +800 (exp's exponent range).
    s = np.array([100.0, 90.0, 100.0, 100.0, 100.0, 1e-310, 100.0, 100.0, 100.0, 100.0])
    k = np.array([90.0, 100.0, 100.0, 90.0, 110.0, 100.0, 1e308, 90.0, 90.0, 90.0])
    r = np.array([0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, -800.0, 800.0])
    v = np.array([0.2, 0.2, 0.2, 0.0, 0.0, 0.2, 0.2, 1e-160, 0.2, 0.2])
    t = np.array([0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    n = len(s)
    for o in (0, 1):
        inputs = (s, k, r, v, t, np.full(n, o, np.int32))
        ref = oracle_prices(inputs, 8)
        for math in (host.MATH_IEEE, host.MATH_FAST):
            got, _, _ = gpu_prices(inputs, 8, math=math)
            assert (np.isnan(got) == np.isnan(ref)).all(), (o, math, got, ref)
            m = ~np.isnan(ref)
            assert np.allclose(got[m], ref[m], rtol=1e-9, atol=1e-12), (o, math, got, ref)


@pytest.mark.parametrize("fp_bytes", [4, 8])
def test_empty_input(fp_bytes):
    with host.BlackScholesGPU(0, fp_bytes=fp_bytes) as bs:
        assert bs.price(3) == 0
        assert bs.prices.shape == (0,)


@pytest.mark.parametrize("fp_bytes", [4, 8])
def test_launch_geometry_does_not_change_results(fp_bytes):
    inputs = inputgen_like(300007, seed=9, dtype=np.float32 if fp_bytes == 4 else np.float64)
    base, _, _ = gpu_prices(inputs, fp_bytes)
    for kw in (dict(unroll=1), dict(unroll=2), dict(unroll=4), dict(threads_per_block=128), dict(threads_per_block=64, blocks_per_sm=3),
               dict(blocks_per_sm=1, unroll=1), dict(use_graph=False), dict(pdl=True), dict(pdl=True, use_graph=False), dict(pdl=True, unroll=2), dict(variant=1), dict(variant=1, unroll=2, threads_per_block=128),
               dict(variant=1, unroll=1, blocks_per_sm=1, threads_per_block=32), dict(variant=4), dict(variant=4, blocks_per_sm=1)):
        got, _, _ = gpu_prices(inputs, fp_bytes, **kw)
        assert got.tobytes() == base.tobytes(), kw


@pytest.mark.parametrize("fp_bytes", [4, 8])
@pytest.mark.parametrize("n", [1, 1023, 1024, 1025, 4096, 300007, 5_000_123])
def test_tma_variant_sizes(n, fp_bytes):
    # bulk-copy (TMA) variant: whole tiles through the shared-memory ring, the remainder by plain loads
    inputs = inputgen_like(n, seed=n + 1, dtype=np.float32 if fp_bytes == 4 else np.float64)
    base, _, _ = gpu_prices(inputs, fp_bytes, num_runs=2)
    got, _, _ = gpu_prices(inputs, fp_bytes, num_runs=2, variant=4)
    assert got.tobytes() == base.tobytes()


@pytest.mark.parametrize("shape", ["0", "1", "2"])
@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 300007, 3_000_001])
def test_tma_shapes_fp64(n, shape, monkeypatch):
    # the three shapes of the fp64 ring kernel (tile 1024 / 1536 / 2048 options; the library picks 0, BS_GPU_TMA_WIDE forces
    # one): each bit-identical to the LDG kernel, whole tiles and remainder
    inputs = inputgen_like(n, seed=n + 7, dtype=np.float64)
    base, _, _ = gpu_prices(inputs, 8, num_runs=2, variant=1)
    monkeypatch.setenv("BS_GPU_TMA_WIDE", shape)
    got, _, _ = gpu_prices(inputs, 8, num_runs=2, variant=4)
    assert got.tobytes() == base.tobytes()
    monkeypatch.delenv("BS_GPU_TMA_WIDE")
    dflt, _, _ = gpu_prices(inputs, 8, num_runs=2)
    assert dflt.tobytes() == base.tobytes()


def test_num_runs_is_idempotent_and_counts_launches():
    inputs = inputgen_like(100003, seed=4)
    with host.BlackScholesGPU(100003) as bs:
        bs.set_inputs(*inputs)
        bs.price(1)
        one = bs.prices.copy()
        assert bs.timing()["kernel_launches"] == 1
        bs.prices[:] = -1.0
        bs.price(host.NUM_RUNS)
        assert bs.timing()["kernel_launches"] == host.NUM_RUNS
        assert bs.prices.tobytes() == one.tobytes()
        assert bs.timing()["h2d_bytes"] == 0            # inputs not dirty: no second upload
        bs.mark_dirty()
        bs.price(2)
        assert bs.timing()["h2d_bytes"] == 100003 * 24  # 5 fp32 + 1 int32 per option
        assert bs.timing()["d2h_bytes"] == 100003 * 4


@pytest.mark.parametrize("fp_bytes", [4, 8])
def test_pipelined_price_any_run_count(fp_bytes):
    # bs_gpu_price pipelines run 0 with the H2D chunks and the last run with the D2H chunks: every run
    # count (0 = copy only, 1 = first is last, 2 = no whole-shard run in between, ...) gives the same prices
    dt = np.float32 if fp_bytes == 4 else np.float64
    n = 300007
    inputs = inputgen_like(n, seed=31, dtype=dt)
    ref = oracle_prices(inputs, fp_bytes)
    with host.BlackScholesGPU(n, fp_bytes=fp_bytes) as bs:
        bs.set_inputs(*inputs)
        bs.price(0)                                   # uploads, launches nothing
        assert bs.timing()["kernel_launches"] == 0
        first = None
        for runs in (1, 2, 3, 7):
            for dirty in (True, False):
                bs.prices[:] = -7.0
                if dirty:
                    bs.mark_dirty()
                bs.price(runs)
                tm = bs.timing()
                assert tm["h2d_bytes"] == (n * (5 * fp_bytes + 4) if dirty else 0) and tm["d2h_bytes"] == n * fp_bytes
                assert tm["pipeline_ms"] > 0 and tm["kernel_launches"] == runs
                assert_parity(bs.prices, ref, fp_bytes, "runs=%d" % runs)
                if first is None:
                    first = bs.prices.copy()
                assert bs.prices.tobytes() == first.tobytes()


def test_phases_equal_price():
    inputs = inputgen_like(50001, seed=6)
    a, _, _ = gpu_prices(inputs, 4, num_runs=3)
    with host.BlackScholesGPU(50001) as bs:
        bs.set_inputs(*inputs)
        bs.upload()
        bs.run(3)
        bs.download()
        assert bs.prices.tobytes() == a.tobytes()
        with pytest.raises(host.BsGpuError):
            host.BlackScholesGPU(10).run(1)  # nothing uploaded yet -> BS_GPU_ERR_STATE


# ---- ERR_CHK (blackscholes.c:333-340, :949-951) --------------------------------------------------
@pytest.mark.parametrize("fp_bytes,sfx", [(4, "f32"), (8, "f64")])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_err_chk_against_reference(name, fp_bytes, sfx):
    inputs, d = _golden_inputs(name, fp_bytes)
    runs = 5
    got, errs, bad = gpu_prices(inputs, fp_bytes, num_runs=runs, dgrefval=d["dgrefval"], err_chk=True)
    # (a) the checker itself is exact: same verdicts as the oracle's ERR_CHK applied to the GPU's prices
    cnt, idx = oracle_lib.errchk(got, d["dgrefval"], fp_bytes, cap=65536)
    assert errs == cnt * runs          # the reference counts once per run
    assert bad.tolist() == idx.tolist()
    # (b) and the verdicts are the reference's wherever a row is not within rounding of the 1e-4 threshold
    gold = json.load(open(golden_path(name, "errchk.json")))[sfx]
    ref_bad = {int(l.split()[2].rstrip(".")) for l in gold["errors_one_run"]}
    ref_prices = _golden_prices(name, sfx)
    margin = np.abs(np.abs(d["dgrefval"].astype(np.float64) - ref_prices) - 1e-4)
    decided = margin > (1e-4 * magnitude_scale(inputs[0], inputs[1]) if fp_bytes == 4 else 1e-9)
    mine = set(bad.tolist())
    for i in np.nonzero(decided)[0].tolist():
        assert (i in mine) == (i in ref_bad), i
    if name != "edge2k":
        assert errs == 0 and gold["num_errors_line"] == "Num Errors: 0"


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_err_chk_reference_mode_gives_exactly_the_reference_verdicts(name):
    # With BS_MATH_REFERENCE the prices ARE the reference's, so its ERR_CHK output must be reproduced with no
    # "within rounding of the threshold" exception at all: the same offender set, the same "Num Errors" total --
    # on edge2k too (hundreds of offenders), where the fast modes need the magnitude-scaled reading
    inputs, d = _golden_inputs(name, 4)
    runs = 100   # NUM_RUNS: the reference counts every offender once per run (blackscholes.c:338)
    got, errs, bad = gpu_prices(inputs, 4, num_runs=runs, dgrefval=d["dgrefval"], err_chk=True, math=host.MATH_REFERENCE)
    gold = json.load(open(golden_path(name, "errchk.json")))["f32"]
    ref_bad = sorted({int(l.split()[2].rstrip(".")) for l in gold["errors_one_run"]})
    assert bad.tolist() == ref_bad
    assert "Num Errors: %d" % errs == gold["num_errors_line"]


def test_err_chk_all_rows_bad_list_is_capped():
    n = 200000
    inputs = inputgen_like(n, seed=2)
    got, errs, bad = gpu_prices(inputs, 4, num_runs=2, dgrefval=np.full(n, -5.0, np.float32), err_chk=True)
    assert errs == 2 * n
    assert len(bad) == 65536 and len(set(bad.tolist())) == 65536


# ---- full-size properties (BASELINE.json configs[1]/[2]: native = 10M options) -----------------------
@pytest.mark.parametrize("fp_bytes", [4, 8])
def test_native_10m_periodicity_and_table_parity(fp_bytes):
    n = 10_000_000
    with host.BlackScholesGPU(n, fp_bytes=fp_bytes) as bs:
        bs.fill_synthetic(0)
        bs.run(2)
        bs.download()
        p = bs.prices
        # the inputgen set repeats every 1000 rows, so must the prices -- bit for bit, over all 10M
        assert (p.reshape(-1, 1000) == p[:1000]).all()
        # and one period equals the reference's answer for the table
        tab = host.load_options(golden_path("table1k", "in.txt"), fp_bytes)
        for k in ("sptprice", "strike", "rate", "volatility", "otime", "otype"):
            assert bs.read_device(k, 0, 1000).tobytes() == tab[k].tobytes(), k
            assert bs.read_device(k, n - 1000, 1000).tobytes() == tab[k].tobytes(), k
        assert bs.read_device("dgrefval", 5000, 1000).tobytes() == tab["dgrefval"].tobytes()
        assert_parity(p[:1000], _golden_prices("table1k", "f32" if fp_bytes == 4 else "f64"), fp_bytes, "table period")
        # ERR_CHK over the whole set: the reference reports 0 errors on this table
        assert bs.run(3, err_chk=True) == 0


def test_native_10m_put_call_parity_and_bounds():
    n = 10_000_000
    s, k, r, v, t, o = inputgen_like(n, seed=77)
    with host.BlackScholesGPU(n, with_dgrefval=False) as bs:
        bs.set_inputs(s, k, r, v, t, np.zeros(n, np.int32))
        bs.price(1)
        call = bs.prices.astype(np.float64)
        bs.host("otype")[:] = 1
        bs.mark_dirty()
        bs.price(1)
        put = bs.prices.astype(np.float64)
    s64, k64, r64, t64 = (a.astype(np.float64) for a in (s, k, r, t))
    fwd = s64 - k64 * np.exp(-r64 * t64)
    assert np.abs((call - put) - fwd).max() <= 1e-4        # C - P = S - K e^{-rT}
    assert (call >= np.maximum(fwd, 0) - 1e-4).all() and (call <= s64 + 1e-4).all()
    assert (put >= np.maximum(-fwd, 0) - 1e-4).all() and (put <= k64 + 1e-4).all()
    # and every one of the 10M calls and puts against the oracle (all host cores: well under a second)
    ref_c = oracle_lib.price_map(s, k, r, v, t, np.zeros(n, np.int32), 4).astype(np.float64)
    ref_p = oracle_lib.price_map(s, k, r, v, t, np.ones(n, np.int32), 4).astype(np.float64)
    worst = max(np.abs(call - ref_c).max(), np.abs(put - ref_p).max())
    print("fp32 fast, 2 x 10M random inputgen-range options: max|delta| vs oracle = %.3e" % worst)
    assert worst <= FP32_ABS_TOL


def test_maximum_option_count_int32():
    # the reference's `int numOptions` tops out at 2^31-1 options: 60 GB of SoA streams on one B200, device-resident
    n = 2**31 - 1
    with host.BlackScholesGPU(n, host_staging=False, with_dgrefval=False) as bs:
        bs.fill_synthetic(0)
        bs.run(2)
        assert bs.timing()["kernel_launches"] == 2
        ref = _golden_prices("table1k", "f32")
        for first in (0, 1_000_000_000, n - 1999, n - 1000):      # n - 1000 covers the ragged last 3 options
            got = bs.read_device("prices", first, 1000).astype(np.float64)
            want = np.roll(ref, -(first % 1000))
            assert np.abs(got - want).max() <= FP32_ABS_TOL, first
        with pytest.raises(host.BsGpuError):
            host.BlackScholesGPU(2**31)                            # one more does not fit an int: rejected


def test_subshard_pipeline_equals_run_order():
    # shards above 256 MiB are priced sub-shard by sub-shard inside bs_gpu_price (copies hidden behind the runs of
    # the neighbouring sub-shard); the prices, the ERR_CHK count and the offender list must not depend on it
    n = 10_000_003                                        # 2 sub-shards, ragged tail in the second
    s, k, r, v, t, o = inputgen_like(n, seed=5)
    ref = oracle_lib.price_map(s[:4096], k[:4096], r[:4096], v[:4096], t[:4096], o[:4096], 4)
    bad_rows = np.array([0, 1, 4_999_999, 5_000_000, 5_000_319, 7_777_777, n - 2, n - 1])
    results = []
    for sub in (True, False):
        with host.BlackScholesGPU(n, subshards=sub) as bs:
            bs.set_inputs(s, k, r, v, t, o)
            bs.price(1)
            dg = bs.prices.copy()
            dg[bad_rows] += 1.0
            bs.host("dgrefval")[:] = dg
            bs.mark_dirty()
            errs = bs.price(3, err_chk=True)
            tm = bs.timing()
            assert tm["h2d_bytes"] == n * 28 and tm["d2h_bytes"] == n * 4 and tm["kernel_launches"] == 3
            results.append((bs.prices.copy(), errs, bs.errors().tolist(), tm["pipeline_ms"]))
    assert results[0][0].tobytes() == results[1][0].tobytes()
    assert results[0][1] == results[1][1] == 3 * len(bad_rows)
    assert results[0][2] == results[1][2] == bad_rows.tolist()
    assert np.abs(results[0][0][:4096] - ref).max() <= FP32_ABS_TOL
    print("bs_gpu_price 10M options x 3 runs: %.2f ms with sub-shards, %.2f ms in run order" % (results[0][3], results[1][3]))


@pytest.mark.parametrize("fp_bytes,math,mname", [(4, host.MATH_IEEE, "ieee"), (8, host.MATH_FAST, "fast"), (8, host.MATH_IEEE, "ieee")])
def test_native_10m_every_option_against_oracle(fp_bytes, math, mname):
    n = 10_000_000
    dt = np.float32 if fp_bytes == 4 else np.float64
    inputs = inputgen_like(n, seed=78, dtype=dt)
    got, _, _ = gpu_prices(inputs, fp_bytes, num_runs=1, math=math, with_dgrefval=False)
    ref = oracle_prices(inputs, fp_bytes)
    worst = assert_parity(got, ref, fp_bytes, "10M/%s" % mname)
    print("fp%d %s, 10M random inputgen-range options: max|delta| vs oracle = %.3e, bit-identical %.1f%%" % (
        8 * fp_bytes, mname, worst, 100.0 * float(np.mean(got == ref))))


def test_scaling_homogeneity():
    # price(2s, 2k) == 2 price(s, k): doubling is exact in binary floating point
    inputs = list(inputgen_like(100000, seed=12))
    a, _, _ = gpu_prices(inputs, 4, math=host.MATH_IEEE)
    inputs[0] = inputs[0] * 2
    inputs[1] = inputs[1] * 2
    b, _, _ = gpu_prices(inputs, 4, math=host.MATH_IEEE)
    assert np.abs(b - 2 * a).max() <= 2e-5



# ---- BS_MATH_REFERENCE: the reference's fp32 build as compiled (operation order + double promotions) ------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp32_reference_mode_against_golden_flat_bound(name):
    # FLAT 1e-4 on every golden, edge2k (operands up to 2000) included: no magnitude scaling for this mode -- and in
    # fact the reference CPU output itself: the prices, written the way the reference writes them ("%.18f",
    # blackscholes.c:936), equal the golden file produced by the compiled reference token for token.
    inputs, d = _golden_inputs(name, 4)
    got, _, _ = gpu_prices(inputs, 4, num_runs=1, math=host.MATH_REFERENCE)
    ref = _golden_prices(name, "f32")
    worst = assert_parity(got, ref, 4, "%s/reference" % name)
    n, toks = oracle_lib.read_prices_text(golden_path(name, "ref_f32.txt"))
    mine = ["%.18f" % float(x) for x in got]
    same = sum(a == b for a, b in zip(mine, toks))
    print("fp32 %-9s reference-mode max|delta| = %.3e, %d of %d output lines identical to the reference's" % (name, worst, same, n))
    assert n == len(mine) and same == n


def test_fp32_reference_mode_native_10m():
    n = 10_000_000
    inputs = inputgen_like(n, seed=79)
    got, _, _ = gpu_prices(inputs, 4, num_runs=1, math=host.MATH_REFERENCE, with_dgrefval=False)
    ref = oracle_prices(inputs, 4)
    worst = assert_parity(got, ref, 4, "10M/reference")
    exact = float(np.mean(got == ref))
    print("fp32 reference mode, 10M random inputgen-range options: max|delta| vs oracle = %.3e, bit-identical %.3f%%" % (worst, 100 * exact))
    assert worst == 0.0 and exact == 1.0


def test_fp32_adversarial_sweep_reduced():
    # one slice (one rate, calls and puts: 4.6M options) of the structured sweep of tests/fp32_adversarial.py -- every (v, t) of
    # the inputgen grid x 8 spots x d1 in [-6, 6] -- in all three modes; the full 246M-option sweep is profiles/r02_fp32_adversarial.json
    import fp32_adversarial as adv
    gen = adv.sweep(True)
    for _ in range(2):
        inputs = next(gen)
        ref = oracle_prices(inputs, 4).astype(np.float64)
        for math, bound in ((host.MATH_FAST, 9e-5), (host.MATH_IEEE, 9e-5), (host.MATH_REFERENCE, 0.0)):
            got, _, _ = gpu_prices(inputs, 4, math=math, with_dgrefval=False)
            worst = float(np.abs(got.astype(np.float64) - ref).max())
            assert worst <= bound, (math, worst)


def test_fp32_reference_mode_err_chk_and_sizes():
    for n in (1, 3, 5, 257, 65537):
        inputs = inputgen_like(n, seed=n + 100)
        got, _, _ = gpu_prices(inputs, 4, num_runs=2, math=host.MATH_REFERENCE)
        assert_parity(got, oracle_prices(inputs, 4), 4, "n=%d/reference" % n)
    inputs, d = _golden_inputs("table1k", 4)
    _, errs, bad = gpu_prices(inputs, 4, num_runs=2, dgrefval=d["dgrefval"], err_chk=True, math=host.MATH_REFERENCE)
    assert errs == 0 and len(bad) == 0   # the reference's own ERR_CHK verdict on this table


# ---- BS_MATH_REFERENCE, fptype=double: IEEE operation order + glibc's own exp/log (csrc/bs_libm_f64.h) ----------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp64_reference_mode_is_the_reference_output(name):
    # the prices, written the way the reference writes them ("%.18f", blackscholes.c:936), equal the golden file produced by
    # the compiled fptype=double reference token for token -- every golden, edge2k included
    inputs, d = _golden_inputs(name, 8)
    got, _, _ = gpu_prices(inputs, 8, num_runs=1, math=host.MATH_REFERENCE)
    ref = _golden_prices(name, "f64")
    n, toks = oracle_lib.read_prices_text(golden_path(name, "ref_f64.txt"))
    mine = ["%.18f" % float(x) for x in got]
    same = sum(a == b for a, b in zip(mine, toks))
    print("fp64 %-9s reference-mode max|delta| = %.3e, %d of %d output lines identical to the reference's" % (name, np.abs(got - ref).max(), same, n))
    assert n == len(mine) and same == n   # (the text holds 18 decimals, not the doubles' bits: those are compared with the oracle below)
    assert got.tobytes() == oracle_prices(inputs, 8).tobytes()


@pytest.mark.parametrize("n", [1, 2, 3, 5, 257, 65537, 1000003])
def test_fp64_reference_mode_bit_identical_to_oracle_sizes(n):
    inputs = inputgen_like(n, seed=n + 300, dtype=np.float64)
    got, _, _ = gpu_prices(inputs, 8, num_runs=2, math=host.MATH_REFERENCE)
    ref = oracle_prices(inputs, 8)
    assert got.tobytes() == ref.tobytes(), "n=%d: %d of %d prices differ, max |delta| %.3e" % (n, int((got != ref).sum()), n, np.abs(got - ref).max())


def test_fp64_reference_mode_native_10m_and_wide_operands():
    # 10M random options of the inputgen range, then 2M with operands far outside it (spot/strike over twelve decades,
    # volatilities and maturities that drive |d| beyond 38, where exp(-d^2/2) is subnormal or zero, and rates that push
    # exp(-r t) through glibc's |x| >= 512 branch): every price equal to the oracle's, NaN and inf included
    n = 10_000_000
    inputs = inputgen_like(n, seed=81, dtype=np.float64)
    got, _, _ = gpu_prices(inputs, 8, num_runs=1, math=host.MATH_REFERENCE, with_dgrefval=False)
    ref = oracle_prices(inputs, 8)
    exact = float(np.mean(got == ref))
    print("fp64 reference mode, 10M random inputgen-range options: bit-identical %.4f%%, max|delta| = %.3e" % (100 * exact, np.abs(got - ref).max()))
    assert got.tobytes() == ref.tobytes()
    rng = np.random.default_rng(5)
    m = 2_000_000
    s = 10.0 ** rng.uniform(-6, 6, m)
    k = s * 10.0 ** rng.uniform(-3, 3, m)
    r = rng.choice([0.0, 0.05, -0.05, 3.0, 600.0, -600.0, 800.0], m, p=[0.1, 0.5, 0.1, 0.1, 0.1, 0.05, 0.05])
    v = 10.0 ** rng.uniform(-4, 1, m)
    t = 10.0 ** rng.uniform(-4, 1.5, m)
    o = rng.integers(0, 2, m).astype(np.int32)
    wide = (s, k, r, v, t, o)
    got, _, _ = gpu_prices(wide, 8, num_runs=1, math=host.MATH_REFERENCE, with_dgrefval=False)
    ref = oracle_prices(wide, 8)
    both_nan = np.isnan(got) & np.isnan(ref)
    same = both_nan | (got.view(np.uint64) == ref.view(np.uint64))
    print("fp64 reference mode, 2M wide-range options: %d differ; %d NaN, %d inf, %d zero or subnormal prices in the reference output"
          % (int((~same).sum()), int(np.isnan(ref).sum()), int(np.isinf(ref).sum()), int((np.abs(ref) < 2.3e-308).sum())))
    assert same.all()


def test_fp64_reference_mode_degenerate_rows_and_err_chk():
    s = np.array([100.0, 90.0, 100.0, 100.0, 100.0, 1e-310, 100.0, 100.0, 100.0, 100.0, 0.0, -5.0])
    k = np.array([90.0, 100.0, 100.0, 90.0, 110.0, 100.0, 1e308, 90.0, 90.0, 90.0, 100.0, 100.0])
    r = np.array([0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, -800.0, 800.0, 0.05, 0.05])
    v = np.array([0.2, 0.2, 0.2, 0.0, 0.0, 0.2, 0.2, 1e-160, 0.2, 0.2, 0.2, 0.2])
    t = np.array([0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    for o in (0, 1):
        inputs = (s, k, r, v, t, np.full(len(s), o, np.int32))
        ref = oracle_prices(inputs, 8)
        got, _, _ = gpu_prices(inputs, 8, math=host.MATH_REFERENCE)
        assert ((np.isnan(got) & np.isnan(ref)) | (got.view(np.uint64) == ref.view(np.uint64))).all(), (o, got, ref)
    inputs, d = _golden_inputs("table1k", 8)
    _, errs, bad = gpu_prices(inputs, 8, num_runs=2, dgrefval=d["dgrefval"], err_chk=True, math=host.MATH_REFERENCE)
    assert errs == 0 and len(bad) == 0   # the fptype=double reference's own ERR_CHK verdict on this table


# ---- the CAF Map's entry: AoS DataCont records (blackscholes.c:482-570) -----------------------------------------
def _datacont(inputs):
    s, k, r, v, t, o = inputs
    rec = np.zeros(len(s), dtype=host.DATACONT_DTYPE)
    rec["otype"], rec["sptprice"], rec["strike"], rec["rate"], rec["volatility"], rec["otime"] = o, s, k, r, v, t
    return rec


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 65537, 1000003])
def test_price_aos_equals_soa(n):
    inputs = inputgen_like(n, seed=n + 7)
    soa, _, _ = gpu_prices(inputs, 4, num_runs=2)
    with host.BlackScholesGPU(n) as bs:
        aos = bs.price_aos(_datacont(inputs), 2)
        tm = bs.timing()
    assert aos.tobytes() == soa.tobytes()
    assert tm["kernel_launches"] == 2 and tm["h2d_bytes"] == 24 * n and tm["d2h_bytes"] == 4 * n


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_price_aos_against_reference_golden(name):
    inputs, d = _golden_inputs(name, 4)
    with host.BlackScholesGPU(len(inputs[0]), math=host.MATH_REFERENCE) as bs:
        got = bs.price_aos(_datacont(inputs), host.NUM_RUNS)
    assert_parity(got, _golden_prices(name, "f32"), 4, "%s/aos" % name)


def test_price_aos_caf_message_quirk_and_errors():
    # the reference's CAF driver builds data_vec(numOptions) and then push_back()s numOptions more (blackscholes.c:772-777):
    # the message holds 2N records, the first N zero-initialised.  Zero records price to NaN, as on the CPU.
    inputs = inputgen_like(1000, seed=3)
    rec = np.concatenate([np.zeros(1000, dtype=host.DATACONT_DTYPE), _datacont(inputs)])
    z = np.zeros(1000, np.float32)
    ref_zero = oracle_lib.price_map(z, z, z, z, z, np.zeros(1000, np.int32), 4)
    for math in (host.MATH_REFERENCE, host.MATH_IEEE, host.MATH_FAST):
        with host.BlackScholesGPU(2000, math=math) as bs:
            got = bs.price_aos(rec, 1)
        assert np.isnan(ref_zero).all() and np.isnan(got[:1000]).all()
        assert_parity(got[1000:], oracle_prices(inputs, 4), 4, "caf message tail")
    with host.BlackScholesGPU(1000) as bs:
        with pytest.raises(host.BsGpuError):
            bs.price_aos(rec, 1)                      # 2000 records into a 1000-option context
    with host.BlackScholesGPU(1000, fp_bytes=8) as bs:
        with pytest.raises(host.BsGpuError):
            bs.price_aos(rec[:1000], 1)               # DataCont holds floats


@pytest.mark.parametrize("name", ["hull4", "table1k", "ragged37"])
def test_caf_message_path_harness(name, tmp_path):
    # integration/blackscholes.c.caf_cuda.patch cannot be compiled (CAF is un-vendored upstream); the harness reproduces the
    # reference's CAF_V3 driver around the patched handler -- DataCont records, the 2N-record message, NUM_RUNS requests
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "caf_harness")
    lib_dir = os.path.dirname(host.LIB_PATH)
    subprocess.run(["g++", "-O2", "-std=c++11", "-I", os.path.join(root, "include"), os.path.join(root, "integration", "caf_map_cuda_harness.cpp"),
                    "-o", exe, "-L", lib_dir, "-lbs_gpu", "-Wl,-rpath," + lib_dir], check=True)
    out = str(tmp_path / "prices.txt")
    cp = subprocess.run([exe, "1", golden_path(name, "in.txt"), out], capture_output=True, text=True)
    assert cp.returncode == 0, cp.stdout + cp.stderr
    ref = _golden_prices(name, "f32")
    n = len(ref)
    assert "message records: %d, prices returned: %d" % (2 * n, 2 * n) in cp.stdout
    assert "zero records priced NaN: %d of %d" % (n, n) in cp.stdout
    cnt, toks = oracle_lib.read_prices_text(out)
    assert cnt == n
    assert_parity(np.array([float(t) for t in toks]), ref, 4, "%s/caf harness" % name)


def test_price_aos_sharded():
    _need_two_gpus("test_price_aos_sharded")
    n = 1000003
    inputs = inputgen_like(n, seed=15)
    g = min(host.device_count(), 8)
    with host.BlackScholesGPU(n, devices=[0]) as bs:
        one = bs.price_aos(_datacont(inputs), 2)
    with host.BlackScholesGPU(n, devices=list(range(g))) as bs:
        many = bs.price_aos(_datacont(inputs), 2)
    assert one.tobytes() == many.tobytes()


# ---- a failed launch must come back as an error, never as prices (VERDICT r1 "Next" 9) --------------------------
@pytest.mark.parametrize("use_graph", [True, False])
@pytest.mark.parametrize("n", [1000, 10_000_003])
def test_launch_failure_is_reported(n, use_graph):
    inputs = inputgen_like(n, seed=2)
    with host.BlackScholesGPU(n, variant=8, use_graph=use_graph, with_dgrefval=False) as bs:   # variant bit 3: fault injection
        bs.set_inputs(*inputs)
        bs.prices[:] = -7.0
        for call in (lambda: bs.price(3), lambda: (bs.upload(), bs.run(3))):
            with pytest.raises(host.BsGpuError) as ei:
                call()
            assert ei.value.status == -3                                  # BS_GPU_ERR_CUDA
            assert "invalid" in str(ei.value).lower() or "shared" in str(ei.value).lower(), str(ei.value)
        assert (bs.prices == -7.0).all()                                  # nothing was written as a "price"
    # the device is still usable afterwards
    got, _, _ = gpu_prices(inputgen_like(1000, seed=2), 4)
    assert_parity(got, oracle_prices(inputgen_like(1000, seed=2), 4), 4, "after a failed launch")


# ---- multi-GPU: contiguous shards, no collective -------------------------------------------------------
def _need_two_gpus(what):
    if host.device_count() < 2:
        msg = "SKIPPED (LOUD): %s needs >= 2 GPUs in ONE process; this box shows %d. The in-process multi-device path " \
              "(bs_gpu_init(ctx, G, ...)) is then covered by bench.py's `inproc` record at --gpus N > 1." % (what, host.device_count())
        print("\n" + "!" * 100 + "\n" + msg + "\n" + "!" * 100)
        pytest.skip(msg)


def test_sharded_equals_single_device():
    _need_two_gpus("test_sharded_equals_single_device")
    n = 1000003
    inputs = inputgen_like(n, seed=21)
    one, _, _ = gpu_prices(inputs, 4, num_runs=2, devices=[0])
    g = min(host.device_count(), 8)
    many, _, _ = gpu_prices(inputs, 4, num_runs=2, devices=list(range(g)))
    assert many.tobytes() == one.tobytes()
    with host.BlackScholesGPU(n, devices=list(range(g))) as bs:
        sh = bs.shards()
        assert [d for d, _, _ in sh] == list(range(g))
        assert sh[0][1] == 0 and sum(c for _, _, c in sh) == n
        assert all(sh[i][1] + sh[i][2] == sh[i + 1][1] for i in range(g - 1))
        assert max(c for _, _, c in sh) - min(c for _, _, c in sh) <= 1


def test_sharded_bs_gpu_price_native_bit_equal():
    # north_star item 3 through the public call: the native set priced by ONE context over every GPU of the box
    # (one host thread + stream per device, gather = each device's D2H into its slice of the pinned prices array)
    _need_two_gpus("test_sharded_bs_gpu_price_native_bit_equal")
    n = 10_000_000
    g = min(host.device_count(), 8)
    inputs = inputgen_like(n, seed=31)
    one, _, _ = gpu_prices(inputs, 4, num_runs=3, devices=[0], with_dgrefval=False)
    many, _, _ = gpu_prices(inputs, 4, num_runs=3, devices=list(range(g)), with_dgrefval=False)
    assert many.tobytes() == one.tobytes()


def test_async_discovery_clamps_to_the_devices_present():
    # BS_GPU_FLAG_ASYNC_DISCOVERY: init never waits for CUDA; num_gpus is an upper bound
    inputs = inputgen_like(100003, seed=8)
    base, _, _ = gpu_prices(inputs, 4)
    with host.BlackScholesGPU(100003, num_gpus=64, async_discovery=True) as bs:
        bs.set_inputs(*inputs)                      # staging buffers are usable before any device is up
        bs.price(2)
        assert len(bs.shards()) == min(64, host.device_count())
        assert bs.prices.tobytes() == base.tobytes()
    with host.BlackScholesGPU(3, num_gpus=64, async_discovery=True) as bs:
        assert len(bs.shards()) <= 3                # never more shards than options


# ---- the drop-in driver binary ------------------------------------------------------------------------
BIN = os.path.join(os.path.dirname(host.LIB_PATH), "..", "bin")


@pytest.mark.parametrize("exe,sfx,fp_bytes", [("blackscholes_gpu", "f32", 4), ("blackscholes_gpu_fp64", "f64", 8)])
@pytest.mark.parametrize("name", ["hull4", "table1k", "ragged37"])
def test_driver_binary_end_to_end(name, exe, sfx, fp_bytes, tmp_path):
    out = str(tmp_path / "prices.txt")
    cp = subprocess.run([os.path.join(BIN, exe), "1", golden_path(name, "in.txt"), out], capture_output=True, text=True)
    assert cp.returncode == 0, cp.stdout + cp.stderr
    lines = cp.stdout.splitlines()
    n = len(_golden_prices(name, sfx))
    assert lines[0] == "PARSEC Benchmark Suite"
    assert "Num of Options: %d" % n in lines and "Num of Runs: 100" in lines
    assert "Size of data: %d" % (n * ((36 if fp_bytes == 4 else 72) + 4)) in lines
    assert any(l.startswith("roi.time|") for l in lines)
    cnt, toks = oracle_lib.read_prices_text(out)
    assert cnt == n and all(len(t.split(".")[1]) == 18 for t in toks)
    assert_parity(np.array([float(t) for t in toks]), _golden_prices(name, sfx), fp_bytes, name)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_driver_binary_reference_math_writes_the_reference_file(name, tmp_path):
    # BS_GPU_MATH=reference: the drop-in driver's prices file is BYTE FOR BYTE the file the reference's fp32 CPU build wrote
    # (tests/golden/*.ref_f32.txt, produced by the compiled reference), and its ERR_CHK lines are the reference's lines
    out = str(tmp_path / "prices.txt")
    env = dict(os.environ, BS_GPU_MATH="reference")
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu_errchk"), "1", golden_path(name, "in.txt"), out], capture_output=True, text=True, env=env)
    assert cp.returncode == 0, cp.stdout + cp.stderr
    assert open(out, "rb").read() == open(golden_path(name, "ref_f32.txt"), "rb").read()
    gold = json.load(open(golden_path(name, "errchk.json")))["f32"]
    lines = cp.stdout.splitlines()
    assert gold["num_errors_line"] in lines
    mine = [l for l in lines if l.startswith("Error on ")]
    assert sorted(set(mine)) == sorted(set(gold["errors_one_run"])) and len(mine) == 100 * len(gold["errors_one_run"])
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu"), "1", golden_path(name, "in.txt"), out], capture_output=True, text=True,
                        env=dict(os.environ, BS_GPU_MATH="bogus"))
    assert cp.returncode == 1 and "BS_GPU_MATH" in cp.stdout


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_driver_binary_fp64_reference_math_writes_the_reference_file(name, tmp_path):
    # the same for the fptype=double build: byte for byte the file the reference's fp64 CPU build wrote, and its ERR_CHK lines
    out = str(tmp_path / "prices.txt")
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu_fp64_errchk"), "1", golden_path(name, "in.txt"), out], capture_output=True, text=True,
                        env=dict(os.environ, BS_GPU_MATH="reference"))
    assert cp.returncode == 0, cp.stdout + cp.stderr
    assert open(out, "rb").read() == open(golden_path(name, "ref_f64.txt"), "rb").read()
    gold = json.load(open(golden_path(name, "errchk.json")))["f64"]
    lines = cp.stdout.splitlines()
    assert gold["num_errors_line"] in lines
    mine = [l for l in lines if l.startswith("Error on ")]
    assert sorted(set(mine)) == sorted(set(gold["errors_one_run"])) and len(mine) == 100 * len(gold["errors_one_run"])


def test_driver_binary_err_chk_and_usage(tmp_path):
    out = str(tmp_path / "p.txt")
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu_errchk"), "1", golden_path("table1k", "in.txt"), out],
                        capture_output=True, text=True)
    assert cp.returncode == 0 and "Num Errors: 0" in cp.stdout.splitlines()
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu_errchk"), "1", golden_path("edge2k", "in.txt"), out],
                        capture_output=True, text=True)
    lines = cp.stdout.splitlines()
    num = [l for l in lines if l.startswith("Num Errors:")]
    errs = [l for l in lines if l.startswith("Error on ")]
    assert len(num) == 1 and int(num[0].split()[-1]) == len(errs) and len(errs) % 100 == 0 and len(errs) > 0
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu")], capture_output=True, text=True)
    assert cp.returncode == 1 and "Usage:" in cp.stdout and "<nthreads> <inputFile> <outputFile>" in cp.stdout
    # more "threads" than options: the reference's warning (blackscholes.c:707-710), then a normal run
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu"), "64", golden_path("hull4", "in.txt"), out], capture_output=True, text=True)
    assert cp.returncode == 0 and "WARNING: Not enough work, reducing number of threads to match number of options." in cp.stdout
    assert_parity(np.loadtxt(out, skiprows=1), _golden_prices("hull4", "f32"), 4, "hull4 with 64 threads")
    # an empty set is legal: header only in, header only out
    empty = tmp_path / "empty.txt"
    empty.write_text("0\n")
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu"), "1", str(empty), out], capture_output=True, text=True)
    assert cp.returncode == 0 and open(out).read() == "0\n"
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu"), "1", str(tmp_path / "missing.txt"), out], capture_output=True, text=True)
    assert cp.returncode == 1 and "ERROR: Unable to open file" in cp.stdout


# ---- the reference's OWN driver with only its Map swapped (integration/blackscholes.c.enable_cuda.patch) ----
@pytest.mark.skipif(oracle_lib.ref_binary("bs_ref_cuda") is None, reason="oracle/_ref/bs_ref_cuda not built")
@pytest.mark.parametrize("exe,sfx,fp_bytes", [("bs_ref_cuda", "f32", 4), ("bs_ref_cuda_fp64", "f64", 8)])
@pytest.mark.parametrize("name", ["hull4", "table1k", "ragged37"])
def test_reference_driver_with_cuda_map(name, exe, sfx, fp_bytes, tmp_path):
    out = str(tmp_path / "prices.txt")
    stdout, roi = oracle_lib.run_ref(exe, 1, golden_path(name, "in.txt"), out)
    assert roi is not None and "Num of Runs: 100" in stdout
    cnt, toks = oracle_lib.read_prices_text(out)
    assert_parity(np.array([float(t) for t in toks]), _golden_prices(name, sfx), fp_bytes, name)


@pytest.mark.skipif(oracle_lib.ref_binary("bs_ref_cuda_errchk") is None, reason="oracle/_ref/bs_ref_cuda_errchk not built")
def test_reference_driver_with_cuda_map_err_chk(tmp_path):
    stdout, _ = oracle_lib.run_ref("bs_ref_cuda_errchk", 1, golden_path("table1k", "in.txt"), str(tmp_path / "p.txt"))
    assert "Num Errors: 0" in stdout.splitlines()
    stdout, _ = oracle_lib.run_ref("bs_ref_cuda_errchk", 1, golden_path("edge2k", "in.txt"), str(tmp_path / "p.txt"))
    lines = stdout.splitlines()
    errs = [l for l in lines if l.startswith("Error on ")]
    num = int([l for l in lines if l.startswith("Num Errors:")][0].split()[-1])
    assert num == len(errs) and num % 100 == 0 and num >= 4800


def test_driver_soa_cache(tmp_path):
    import shutil
    inp = str(tmp_path / "in.txt")
    shutil.copy(golden_path("table1k", "in.txt"), inp)
    env = dict(os.environ, BS_GPU_SOA_CACHE="1")
    outs = []
    for expect in ("text (side-car written)", "binary SoA side-car (cache hit)"):
        out = str(tmp_path / ("p%d.txt" % len(outs)))
        cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu"), "1", inp, out], capture_output=True, text=True, env=env)
        assert cp.returncode == 0 and ("[BS_GPU] input: " + expect) in cp.stdout, cp.stdout
        outs.append(open(out).read())
    assert outs[0] == outs[1] and os.path.exists(inp + ".bssoa")
    # the side-car itself is a valid input file
    out = str(tmp_path / "p2.txt")
    cp = subprocess.run([os.path.join(BIN, "blackscholes_gpu_errchk"), "1", inp + ".bssoa", out], capture_output=True, text=True)
    assert cp.returncode == 0 and "binary SoA file" in cp.stdout and "Num Errors: 0" in cp.stdout and open(out).read() == outs[0]


@pytest.mark.skipif(oracle_lib.ref_binary("blackscholes_gpu_hooks") is None, reason="oracle/_ref/blackscholes_gpu_hooks not built")
def test_driver_with_parsec_hooks(tmp_path):
    # -DENABLE_PARSEC_HOOKS build of the drop-in driver (against the hooks shim): the ROI lines come from the hooks
    out = str(tmp_path / "p.txt")
    stdout, roi = oracle_lib.run_ref("blackscholes_gpu_hooks", 1, golden_path("table1k", "in.txt"), out)
    lines = stdout.splitlines()
    assert roi is not None and "[HOOKS] shim (oracle/hooks_shim/hooks.h)" in lines
    assert lines.index("[HOOKS] Entering ROI") < lines.index("[HOOKS] Leaving ROI") < lines.index("[HOOKS] Terminating")
    assert_parity(np.loadtxt(out, skiprows=1), _golden_prices("table1k", "f32"), 4, "hooks build")


def test_rt0_degenerate_fp32_fast_follows_reference():
    # t = 0 / v = 0 in fp32: the fast path must give what the reference gives (intrinsic value or NaN), via the
    # -inf clamp of ex2_accurate and MUFU's IEEE specials
    s = np.array([100.0, 90.0, 100.0, 100.0, 100.0], np.float32)
    k = np.array([90.0, 100.0, 100.0, 90.0, 110.0], np.float32)
    r = np.full(5, 0.05, np.float32)
    v = np.array([0.2, 0.2, 0.2, 0.0, 0.0], np.float32)
    t = np.array([0.0, 0.0, 0.0, 1.0, 1.0], np.float32)
    for o in (0, 1):
        inputs = (s, k, r, v, t, np.full(5, o, np.int32))
        ref = oracle_prices(inputs, 4)
        for math in (host.MATH_IEEE, host.MATH_FAST):
            got, _, _ = gpu_prices(inputs, 4, math=math)
            assert (np.isnan(got) == np.isnan(ref)).all(), (o, math, got, ref)
            m = ~np.isnan(ref)
            assert np.allclose(got[m], ref[m], rtol=0, atol=1e-4), (o, math, got, ref)
