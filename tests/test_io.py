"""Loader / writer of libbs_gpu.so (include/bs_io.h) against the oracle's fscanf/fprintf restatement.

CPU-only: these are host functions of the library; no GPU call is made.
"""
import ctypes
import os
import struct

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN_CASES, golden_path
from p3arsec_b200 import host

KEYS = ("sptprice", "strike", "rate", "volatility", "otime", "otype", "dgrefval", "divq", "divs")


@pytest.mark.parametrize("fp_bytes", [4, 8])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_loader_bit_identical_to_fscanf(name, fp_bytes):
    path = golden_path(name, "in.txt")
    mine = host.load_options(path, fp_bytes)
    ref = oracle_lib.load(path, fp_bytes)
    assert mine["numOptions"] == ref["numOptions"]
    for k in KEYS:
        assert mine[k].dtype == ref[k].dtype
        assert mine[k].tobytes() == ref[k].tobytes(), k


@pytest.mark.parametrize("nthreads", [1, 2, 3, 7])
def test_loader_any_thread_count(nthreads):
    path = golden_path("edge2k", "in.txt")
    a = host.load_options(path, 4, nthreads=nthreads)
    b = oracle_lib.load(path, 4)
    for k in KEYS:
        assert a[k].tobytes() == b[k].tobytes()


def test_loader_large_parallel(tmp_path):
    # big enough that several byte ranges are really parsed concurrently
    exe = os.path.join(os.path.dirname(host.LIB_PATH), "..", "bin", "bs_inputgen")
    path = str(tmp_path / "in_200k.txt")
    assert os.system("%s 200003 %s" % (exe, path)) == 0
    a = host.load_options(path, 4)
    b = oracle_lib.load(path, 4)
    assert a["numOptions"] == 200003
    for k in KEYS:
        assert a[k].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize("fp_bytes", [4, 8])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_loader_random_tokens_bit_identical_to_fscanf(tmp_path, fp_bytes, seed):
    """Differential test of the token parser against fscanf (the oracle's loader): numbers of 1-25 digits with and without
    a fractional part, signs, leading zeros, trailing dots, tokens longer than the fast path's window, exponent and hex
    spellings, random white space -- on every thread count, so that byte-range boundaries fall inside tokens."""
    rng = np.random.RandomState(seed)

    def number():
        kind = rng.randint(0, 12)
        nint, nfrac = int(rng.randint(0, 9)), int(rng.randint(0, 9))
        if kind == 0:
            nint, nfrac = int(rng.randint(1, 26)), int(rng.randint(0, 26))     # more digits than one exact division takes
        elif kind == 1:
            nint, nfrac = int(rng.randint(31, 45)), 0                            # longer than the fast path's window
        ip = "".join(str(d) for d in rng.randint(0, 10, nint))
        fp = "".join(str(d) for d in rng.randint(0, 10, nfrac))
        if kind == 2:
            ip = "000" + ip
        tok = ip + ("." + fp if (nfrac or rng.rand() < 0.2) else "")
        if not any(ch.isdigit() for ch in tok):
            tok = "0" + tok
        if kind == 3:
            tok += "e%d" % rng.randint(-12, 12)
        if kind == 4:
            tok = "0x1.%xp%d" % (rng.randint(0, 1 << 20), rng.randint(-8, 8))
        if rng.rand() < 0.15:
            tok = "-" + tok
        elif kind == 5:
            tok = "+" + tok
        return tok

    n = 3000
    seps = [" ", "  ", "\t", " \t ", "\n"]
    rows = []
    for _ in range(n):
        toks = [number() for _ in range(6)] + [rng.choice(list("PCpx"))] + [number(), number()]
        rows.append("".join(t + seps[rng.randint(0, len(seps))] for t in toks))
    path = _write(tmp_path, "%d\n" % n + "\n".join(rows) + "\n")
    b = oracle_lib.load(path, fp_bytes)
    for nthreads in (1, 3, 8):
        a = host.load_options(path, fp_bytes, nthreads=nthreads)
        assert a["numOptions"] == n
        for k in KEYS:
            assert a[k].tobytes() == b[k].tobytes(), (k, nthreads)


def _write(tmp_path, text):
    p = tmp_path / "in.txt"
    p.write_text(text)
    return str(p)


ROW = "42.00 40.00 0.1000 0.00 0.20 0.50 C 0.00 4.759423036851750055"


def test_loader_free_form_whitespace_and_extra_rows(tmp_path):
    # fscanf does not care about line structure; rows beyond numOptions are ignored
    text = " 2\n\n" + ROW.replace(" ", "\t") + "   " + ROW.replace(" ", "\n") + "\n" + ROW + "\n"
    path = _write(tmp_path, text)
    a, b = host.load_options(path, 4), oracle_lib.load(path, 4)
    assert a["numOptions"] == 2
    for k in KEYS:
        assert a[k].tobytes() == b[k].tobytes()


@pytest.mark.parametrize("header,n", [("0x3", 3), ("011", 9), ("+2", 2)])
def test_header_is_percent_i(tmp_path, header, n):
    path = _write(tmp_path, header + "\n" + "\n".join([ROW] * 9) + "\n")
    assert host.read_header(path) == n
    assert oracle_lib.load(path, 4)["numOptions"] == n
    assert host.load_options(path, 4)["numOptions"] == n


def test_loader_odd_tokens_fall_back_to_fscanf(tmp_path):
    # '+' signs, exponents, hex floats, inf: all legal for fscanf("%f")
    row = "+42.00 4e1 0x1.999999999999ap-4 0.00 .20 5e-1 P 0.00 inf"
    path = _write(tmp_path, "1\n" + row + "\n")
    a, b = host.load_options(path, 4), oracle_lib.load(path, 4)
    for k in KEYS:
        assert a[k].tobytes() == b[k].tobytes(), k
    # glued type char: "%c" takes one char, "%f" continues right after it
    path = _write(tmp_path, "1\n42.00 40.00 0.1000 0.00 0.20 0.50 P0.00 4.75 \n")
    a, b = host.load_options(path, 4), oracle_lib.load(path, 4)
    assert b["otype"].tolist() == [1] and float(b["dgrefval"][0]) == pytest.approx(4.75)
    for k in KEYS:
        assert a[k].tobytes() == b[k].tobytes(), k


def test_only_capital_p_is_a_put(tmp_path):
    """blackscholes.c:761: otype = (OptionType == 'P') ? 1 : 0 -- any other character, 'p' included, prices as a call
    (SURVEY.md appendix A.2).  Fast path of the loader and its fscanf fallback against the oracle's loader."""
    rows = ["42.00 40.00 0.1000 0.00 0.20 0.50 %s 0.00 4.75" % c for c in "PCpcXx?1"]
    rows.append("+42.00 4e1 0.1000 0.00 .20 5e-1 p 0.00 4.75")   # odd tokens: the fscanf path
    path = _write(tmp_path, "%d\n" % len(rows) + "\n".join(rows) + "\n")
    a, b = host.load_options(path, 4), oracle_lib.load(path, 4)
    assert b["otype"].tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 0]
    assert a["otype"].tolist() == b["otype"].tolist()


@pytest.mark.parametrize("text", ["", "abc\n", "3\n" + ROW + "\n", "1\n42.00 40.00 0.1000 0.00 0.20 0.50 C 0.00\n",
                                  "1\n42.00 40.00 x 0.00 0.20 0.50 C 0.00 4.7\n"])
def test_loader_errors_like_reference(tmp_path, text):
    path = _write(tmp_path, text)
    with pytest.raises(IOError):
        oracle_lib.load(path, 4)
    with pytest.raises(host.BsIoError) as ei:
        host.load_options(path, 4)
    assert ei.value.status == host.IO_ERR_READ


def test_loader_missing_file(tmp_path):
    with pytest.raises(host.BsIoError) as ei:
        host.load_options(str(tmp_path / "nope.txt"), 4)
    assert ei.value.status == host.IO_ERR_OPEN


@pytest.mark.parametrize("fp_bytes,sfx", [(4, "f32"), (8, "f64")])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_writer_reproduces_reference_files(name, fp_bytes, sfx, tmp_path):
    # parse the reference's own output, write it back, expect the same bytes
    gold = golden_path(name, "ref_%s.txt" % sfx)
    n, toks = oracle_lib.read_prices_text(gold)
    vals = np.array([float(t) for t in toks], dtype=np.float32 if fp_bytes == 4 else np.float64)
    out = str(tmp_path / "out.txt")
    host.write_prices(out, vals)
    assert open(out).read() == open(gold).read()


def test_writer_exact_formatting_fuzz(tmp_path):
    rng = np.random.RandomState(5)
    bits = rng.randint(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)
    f32 = bits.view(np.float32)
    f32 = f32[np.isfinite(f32)]
    specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e-45, -1e-45, 1e-38, 3.4e38, 0.1, 0.3, 123456.789, 2.5e-19, 5e-19,
                         7.5e-19, 1.5e-18, 0.9999999, 9.999999, 16777216.0], dtype=np.float32)
    vals = np.concatenate([specials, f32, rng.uniform(0, 200, 100000).astype(np.float32)])
    out = str(tmp_path / "o.txt")
    host.write_prices(out, vals, nthreads=3)
    got = open(out).read().split("\n")
    assert got[0] == str(len(vals))
    for i in rng.choice(len(vals), 20000, replace=False).tolist() + list(range(len(specials))):
        assert got[1 + i] == "%.18f" % float(vals[i]), (i, vals[i])
    # doubles, including ties at the 18th decimal and values near powers of ten
    d = np.concatenate([rng.uniform(0, 100, 50000), rng.uniform(0, 1e-15, 1000), 10.0 ** rng.uniform(-30, 18, 5000),
                        np.array([0.5e-18, 1.5e-18, 2.5e-18, 0.999999999999999999, 1e18, 9.007199254740993e15, 1e300, 5e-324])])
    for x in d.tolist():
        assert host.format_price(x) == "%.18f\n" % x, x
        assert host.format_price(-x) == "%.18f\n" % -x, x
    for x in (float("inf"), float("-inf"), float("nan")):
        assert host.format_price(x) == "%.18f\n" % x


def test_writer_matches_oracle_writer_on_real_prices(tmp_path):
    d = oracle_lib.load(golden_path("edge2k", "in.txt"), 4)
    p = oracle_lib.price_map(d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"], 4)
    a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    host.write_prices(a, p)
    oracle_lib.write(b, p, 4)
    assert open(a).read() == open(b).read()


def test_writer_unwritable_path(tmp_path):
    with pytest.raises(host.BsIoError) as ei:
        host.write_prices(str(tmp_path / "no_such_dir" / "x.txt"), np.zeros(3, np.float32))
    assert ei.value.status == host.IO_ERR_OPEN


def test_empty_set(tmp_path):
    path = _write(tmp_path, "0\n")
    d = host.load_options(path, 4)
    assert d["numOptions"] == 0
    out = str(tmp_path / "o.txt")
    host.write_prices(out, np.zeros(0, np.float32))
    assert open(out).read() == "0\n"


# ---- binary SoA side-car (include/bs_io.h) ------------------------------------------
@pytest.mark.parametrize("fp_bytes", [4, 8])
def test_soa_sidecar_round_trip(fp_bytes, tmp_path):
    src = golden_path("edge2k", "in.txt")
    d = host.load_options(src, fp_bytes)
    soa = str(tmp_path / "edge2k.bssoa")
    host.write_soa(soa, d, src)
    assert os.path.getsize(soa) % 4096 == 0
    e = host.load_options(soa, fp_bytes)                       # bs_io_open recognises the magic
    assert e["numOptions"] == d["numOptions"]
    for k in ("sptprice", "strike", "rate", "volatility", "otime", "otype", "dgrefval"):
        assert e[k].tobytes() == d[k].tobytes(), k
    assert host.soa_matches(soa, src, fp_bytes) and not host.soa_matches(soa, src, 12 - fp_bytes)
    assert not host.soa_matches(soa, golden_path("hull4", "in.txt"), fp_bytes)
    with pytest.raises(host.BsIoError):                        # wrong fptype is refused, not converted
        host.load_options(soa, 12 - fp_bytes)


def test_soa_sidecar_goes_stale_with_its_source(tmp_path):
    src = tmp_path / "in.txt"
    src.write_text("1\n" + ROW + "\n")
    d = host.load_options(str(src), 4)
    soa = str(tmp_path / "in.txt.bssoa")
    host.write_soa(soa, d, str(src))
    assert host.soa_matches(soa, str(src), 4)
    src.write_text("1\n" + ROW.replace("42.00", "43.00") + "\n")
    os.utime(str(src), ns=(1, 1))
    assert not host.soa_matches(soa, str(src), 4)
    assert not host.soa_matches(str(tmp_path / "missing.bssoa"), str(src), 4)


def test_soa_rejects_truncated_file(tmp_path):
    d = host.load_options(golden_path("table1k", "in.txt"), 4)
    soa = str(tmp_path / "t.bssoa")
    host.write_soa(soa, d)
    blob = open(soa, "rb").read()
    open(soa, "wb").write(blob[: len(blob) // 2])
    with pytest.raises(host.BsIoError):
        host.load_options(soa, 4)


@pytest.mark.parametrize("stride", [2**64 - 4096, (2**64) // 7 + 4096, 8, 0])
def test_soa_rejects_a_crafted_stream_stride(stride, tmp_path):
    # a huge stream_stride must not be able to wrap the size computation of the header check (the file would
    # then be "big enough" and the loader would copy from far outside the mapping); a stride too small for the streams
    # is refused as well
    import struct
    d = host.load_options(golden_path("table1k", "in.txt"), 4)
    soa = str(tmp_path / "t.bssoa")
    host.write_soa(soa, d)
    blob = bytearray(open(soa, "rb").read())
    # SoaHeader: magic[8], u32 version, u32 fp_bytes, u64 num_options, u64 stream_stride, ...
    assert struct.unpack_from("<Q", blob, 16)[0] == 1000
    struct.pack_into("<Q", blob, 24, stride)
    open(soa, "wb").write(bytes(blob))
    with pytest.raises(host.BsIoError):
        host.load_options(soa, 4)
