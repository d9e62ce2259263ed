"""The oracle (oracle/bs_oracle.c) against outputs of the reference itself.

Bit-exact: the "%.18f" text the restatement writes must equal, byte for byte, what the unmodified
reference binaries wrote for the same input (tests/golden/*.ref_f32.txt / *.ref_f64.txt, produced by
tests/golden/make_golden.py), and -- where oracle/_ref is present -- what they write right now.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN_CASES, golden_path

FP = [(4, "f32"), (8, "f64")]


def _oracle_prices(name, fp_bytes):
    d = oracle_lib.load(golden_path(name, "in.txt"), fp_bytes)
    p = oracle_lib.price_map(d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"], fp_bytes)
    return d, p


@pytest.mark.parametrize("fp_bytes,sfx", FP)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_output_bit_for_bit(name, fp_bytes, sfx, tmp_path):
    d, p = _oracle_prices(name, fp_bytes)
    out = str(tmp_path / "prices.txt")
    oracle_lib.write(out, p, fp_bytes)
    assert open(out).read() == open(golden_path(name, "ref_%s.txt" % sfx)).read()


@pytest.mark.parametrize("fp_bytes,sfx", FP)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_errchk_matches_reference_count(name, fp_bytes, sfx):
    d, p = _oracle_prices(name, fp_bytes)
    cnt, idx = oracle_lib.errchk(p, d["dgrefval"], fp_bytes, cap=4096)
    gold = json.load(open(golden_path(name, "errchk.json")))[sfx]
    # the reference counts once per run: "Num Errors" == NUM_RUNS x per-pass count
    assert gold["num_errors_line"] == "Num Errors: %d" % (cnt * gold["num_runs"])
    # and names the same offending rows, with the same printed values
    lines = ["Error on %d. Computed=%.5f, Ref=%.5f, Delta=%.5f" % (i, p[i], d["dgrefval"][i], d["dgrefval"][i] - p[i])
             for i in idx]
    assert lines == gold["errors_one_run"]


def test_known_answers_hull():
    # values the reference build printed for the Hull textbook rows
    d, p = _oracle_prices("hull4", 4)
    assert ["%.18f" % x for x in p] == ["4.759418487548828125", "0.808598518371582031",
                                        "3.714607238769531250", "8.591659545898437500"]
    d, p = _oracle_prices("hull4", 8)
    assert ["%.18f" % x for x in p] == ["4.759422997128201160", "0.808599977156765348",
                                        "3.714601868896551196", "8.591659418825138061"]


def test_otype_mapping_only_P_is_put():
    d = oracle_lib.load(golden_path("edge2k", "in.txt"), 4)
    chars = [l.split()[6] for l in open(golden_path("edge2k", "in.txt")).read().splitlines()[1:]]
    assert set(chars) - {"P", "C"}, "fixture must hold type chars beyond P/C"
    assert d["otype"].tolist() == [1 if c == "P" else 0 for c in chars]


def test_cndf_symmetry_and_limits():
    for fp in (4, 8):
        assert oracle_lib.cndf(0.0, fp) == pytest.approx(0.5, abs=1e-7)
        for x in (0.1, 0.7, 1.96, 3.3, 6.0):
            assert oracle_lib.cndf(x, fp) + oracle_lib.cndf(-x, fp) == pytest.approx(1.0, abs=2e-7)
        assert oracle_lib.cndf(40.0, fp) == 1.0
        assert oracle_lib.cndf(-40.0, fp) == 0.0


def test_put_call_parity_fp64():
    rng = np.random.RandomState(3)
    n = 2000
    s = rng.uniform(20, 120, n); k = s * rng.uniform(0.7, 1.3, n); r = rng.uniform(0.0, 0.1, n)
    v = rng.uniform(0.05, 0.65, n); t = rng.uniform(0.05, 1.0, n)
    c = oracle_lib.price_map(s, k, r, v, t, np.zeros(n, np.int32), 8)
    p = oracle_lib.price_map(s, k, r, v, t, np.ones(n, np.int32), 8)
    np.testing.assert_allclose(c - p, s - k * np.exp(-r * t), rtol=0, atol=1e-10)


def test_loader_rejects_short_row(tmp_path):
    bad = tmp_path / "bad.txt"
    bad.write_text("2\n42.00 40.00 0.1000 0.00 0.20 0.50 C 0.00 4.7\n42.00 40.00 0.1000\n")
    with pytest.raises(IOError):
        oracle_lib.load(str(bad), 4)


@pytest.mark.skipif(oracle_lib.ref_binary("bs_ref_ff") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("exe,fp_bytes", [("bs_ref_ff", 4), ("bs_ref_skepu", 4), ("bs_ref_ff_fp64", 8)])
def test_oracle_matches_live_reference_binary(exe, fp_bytes, tmp_path):
    # not a stored fixture: run the compiled reference here and now on a fresh seeded input
    rng = np.random.RandomState(11)
    n = 5003
    rows = []
    for i in range(n):
        s = rng.uniform(5, 300); k = s * rng.uniform(0.5, 1.7)
        rows.append("%.4f %.4f %.4f 0.00 %.4f %.4f %s 0.00 0.0" % (
            s, k, rng.uniform(0, 0.15), rng.uniform(0.03, 0.9), rng.uniform(0.01, 3.0), "PC"[i % 2]))
    inp = tmp_path / "in.txt"
    inp.write_text("%d\n%s\n" % (n, "\n".join(rows)))
    ref_out = str(tmp_path / "ref.txt")
    oracle_lib.run_ref(exe, 3, str(inp), ref_out)
    d = oracle_lib.load(str(inp), fp_bytes)
    p = oracle_lib.price_map(d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"], fp_bytes)
    mine = str(tmp_path / "mine.txt")
    oracle_lib.write(mine, p, fp_bytes)
    assert open(mine).read() == open(ref_out).read()
