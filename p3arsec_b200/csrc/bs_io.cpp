/*
 * bs_io.cpp -- parallel loader / writer behind include/bs_io.h.
 *
 * Loader: the file is mmap'ed; all host cores first count whitespace-separated tokens in their byte
 * range, a prefix sum gives every range its first token number, and then each core converts its tokens
 * (std::from_chars: correctly rounded, so bit-identical to the strtof/strtod inside fscanf) directly
 * into the SoA destination arrays: token t belongs to row t/9, field t%9 of the reference's
 * "%f %f %f %f %f %f %c %f %f" grammar (blackscholes.c:728).  Anything that is not a plain token of the
 * expected kind (glued fields, '+' signs, hex floats, ...) makes the loader start over with the C
 * library's fscanf, i.e. the reference's own code path, so odd files behave identically.
 *
 * Writer: "%.18f" is produced exactly (the decimal expansion of a binary float is finite, so 18
 * fractional digits with round-half-even need only a 128-bit integer), validated against snprintf in
 * tests/test_io.py; non-finite or huge values fall back to snprintf itself.
 */
#include "../../include/bs_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

struct bs_io_file {
    std::string path;
    const char *data = nullptr;  // mmap of the whole file
    size_t size = 0;
    size_t body = 0;  // offset just past the header number
    int fd = -1;
    bool soa = false;  // binary side-car instead of text
};

// ---- binary SoA side-car --------------------------------------------------------------------------------
struct SoaHeader {
    char magic[8];  // "BSSOA\1\0\0"
    uint32_t version;
    uint32_t fp_bytes;
    uint64_t num_options;
    uint64_t stream_stride;  // bytes between consecutive streams (each padded to SOA_ALIGN)
    uint64_t src_size;       // size / mtime of the text file this was parsed from (0 = generated)
    uint64_t src_mtime_ns;
    uint32_t has_dgrefval;
    uint32_t reserved;
};
static const char SOA_MAGIC[8] = {'B', 'S', 'S', 'O', 'A', 1, 0, 0};
static const size_t SOA_ALIGN = 4096;
static const int SOA_STREAMS = 7;  // sptprice strike rate volatility otime otype dgrefval

namespace {

// isspace() of the C locale: ' ' and \t \n \v \f \r (9..13)
inline bool is_space(char ch) { return ch == ' ' || (unsigned char)(ch - 9) <= 4; }

int host_threads(int want, size_t work_items, size_t min_per_thread)
{
    long n = want > 0 ? want : sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (want <= 0 && sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n < 1) n = 1;
    const size_t cap = std::max<size_t>(1, work_items / std::max<size_t>(min_per_thread, 1));
    return (int)std::min<size_t>((size_t)n, cap);
}

template <typename F> void parallel_for_chunks(int nthreads, F fn)
{
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; t++) th.emplace_back(fn, t);
    fn(0);
    for (auto &x : th) x.join();
}

template <typename FP> struct Dest {
    FP *f[9];  // per field; [6] unused
    int *otype;
};

const double POW10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18,
                          1e19, 1e20, 1e21, 1e22};

// Convert one token.  Returns false when the token is not a plain decimal float (caller falls back).
// Fast path (Clinger): a plain decimal [-]ddd[.ddd] whose digit string is an integer m exactly representable
// in FP (m < 2^24 for float, < 2^53 for double) divided by an exactly representable power of ten is ONE
// correctly rounded division, hence equal to strtof/strtod.  Plain decimals that are not wanted (`out` NULL,
// e.g. the 20-digit DGrefval column of a build without ERR_CHK) are validated syntactically and skipped: such
// a string cannot fail to convert.  Everything else goes through std::from_chars.
template <typename FP> inline bool parse_fp(const char *b, const char *e, FP *out)
{
    {
        const char *p = b;
        const bool neg = (p < e && *p == '-');
        if (neg) p++;
        uint64_t m = 0;
        int digits = 0, frac = 0, int_digits = 0;
        bool seen_dot = false, plain = (p < e);
        for (; p < e; p++) {
            const unsigned d = (unsigned)(*p - '0');
            if (d <= 9) {
                if (digits < 19) m = m * 10 + d;
                digits++;
                if (seen_dot) frac++; else int_digits++;
            } else if (*p == '.' && !seen_dot) {
                seen_dot = true;
            } else {
                plain = false;
                break;
            }
        }
        if (plain && digits > 0 && int_digits <= 30) {
            if (!out) return true;
            const uint64_t exact_limit = sizeof(FP) == 4 ? (1ull << 24) : (1ull << 53);
            const int pow_limit = sizeof(FP) == 4 ? 10 : 22;
            if (digits <= 19 && m < exact_limit && frac <= pow_limit) {
                const FP v = (FP)m / (FP)POW10[frac];
                *out = neg ? -v : v;
                return true;
            }
        }
    }
    FP v;
    if (sizeof(FP) == 4) {
        // libstdc++'s from_chars<float> serialises under load (measured 7x slower per token with 8 threads);
        // from_chars<double> does not.  Rounding the correctly rounded double to float is itself correctly rounded
        // unless that double IS a float midpoint (low 29 mantissa bits == 1000...0): rounding to double is
        // monotone and midpoints are doubles, so otherwise text and double lie on the same side of every
        // midpoint.  Midpoints and values outside the normal float range take the float parser.
        double d;
        auto rd = std::from_chars(b, e, d, std::chars_format::general);
        if (rd.ec == std::errc() && rd.ptr == e) {
            uint64_t bits;
            memcpy(&bits, &d, 8);
            const double ad = d < 0 ? -d : d;
            if ((bits & 0x1fffffffull) != 0x10000000ull && (ad == 0.0 || (ad > 1e-30 && ad < 1e30))) {
                const char c0 = (*b == '-') ? (e - b > 1 ? b[1] : 0) : *b;
                if (!((c0 >= '0' && c0 <= '9') || c0 == '.')) return false;
                if (out) *out = (FP)d;
                return true;
            }
        }
    }
    auto r = std::from_chars(b, e, v, std::chars_format::general);
    if (r.ec != std::errc() || r.ptr != e) return false;
    // "inf"/"nan" spellings are legal for both parsers but let fscanf decide on anything exotic
    const char c0 = (*b == '-') ? (e - b > 1 ? b[1] : 0) : *b;
    if (!((c0 >= '0' && c0 <= '9') || c0 == '.')) return false;
    if (out) *out = v;
    return true;
}

// The common case in one pass over the token: a plain decimal [-]ddd[.ddd] of at most 19 digits, terminated by white space
// (or the end of the file), converted by Clinger's exact division as in parse_fp.  Returns 1 with *end behind the token, or
// 0 for anything else -- exponents, hex floats, inf, more digits than one division can take, over-long tokens -- which the
// caller hands to parse_fp (same result for the tokens both accept: this is its fast path without the second scan).
template <typename FP> inline int parse_plain(const char *buf, size_t p, size_t hi, FP *out, size_t *end)
{
    const size_t lim = std::min(hi, p + 40);
    size_t q = p;
    const bool neg = buf[q] == '-';
    q += neg;
    if (!out) {  // an unwanted column (divq, divs, the 20-digit DGrefval of a build without ERR_CHK): syntax only, no value
        const size_t d0 = q;
        while (q < lim && (unsigned)(buf[q] - '0') <= 9) q++;
        const size_t int_digits = q - d0;
        size_t frac = 0;
        if (q < lim && buf[q] == '.') {
            const size_t f0 = ++q;
            while (q < lim && (unsigned)(buf[q] - '0') <= 9) q++;
            frac = q - f0;
        }
        if (q < hi && !is_space(buf[q])) return 0;
        if (int_digits + frac == 0 || int_digits > 30) return 0;
        *end = q;
        return 1;
    }
    uint64_t m = 0;
    const size_t d0 = q;
    for (; q < lim; q++) {
        const unsigned d = (unsigned)(buf[q] - '0');
        if (d > 9) break;
        m = m * 10 + d;
    }
    const size_t int_digits = q - d0;
    size_t frac = 0;
    if (q < lim && buf[q] == '.') {
        const size_t f0 = ++q;
        for (; q < lim; q++) {
            const unsigned d = (unsigned)(buf[q] - '0');
            if (d > 9) break;
            m = m * 10 + d;
        }
        frac = q - f0;
    }
    const size_t digits = int_digits + frac;
    if (q < hi && !is_space(buf[q])) return 0;  // not plain, or cut off by the 40-byte window
    if (digits == 0 || int_digits > 30) return 0;
    *end = q;
    const uint64_t exact_limit = sizeof(FP) == 4 ? (1ull << 24) : (1ull << 53);
    const size_t pow_limit = sizeof(FP) == 4 ? 10 : 22;
    if (digits > 19 || m >= exact_limit || frac > pow_limit) return 0;
    const FP v = (FP)m / (FP)POW10[frac];
    *out = neg ? -v : v;
    return 1;
}

template <typename FP>
int load_fast(const bs_io_file *f, size_t count, const Dest<FP> &dst, int nthreads)
{
    const char *buf = f->data;
    const size_t lo = f->body, hi = f->size;
    const size_t want_tokens = count * 9;
    if (count == 0) return BS_IO_OK;
    const int T = host_threads(nthreads, hi - lo, 1 << 16);
    std::vector<size_t> start(T + 1), ntok(T + 1, 0);
    for (int t = 0; t <= T; t++) start[t] = lo + (hi - lo) / T * t;
    start[T] = hi;

    // pass 1: token starts per byte range
    parallel_for_chunks(T, [&](int t) {
        // a token starts where a non-space follows a space; written without a loop-carried flag so that the
        // compiler vectorises it
        size_t n = 0, p = start[t];
        if (p == lo && p < start[t + 1]) { n += !is_space(buf[p]); p++; }
        for (; p < start[t + 1]; p++) n += (size_t)(!is_space(buf[p]) & is_space(buf[p - 1]));
        ntok[t + 1] = n;
    });
    for (int t = 0; t < T; t++) ntok[t + 1] += ntok[t];
    if (ntok[T] < want_tokens) return BS_IO_ERR_READ;  // short file: let fscanf produce the verdict

    // pass 2: convert
    std::vector<int> bad(T, 0);
    parallel_for_chunks(T, [&](int t) {
        size_t tok = ntok[t];
        size_t p = start[t];
        const size_t end = start[t + 1];
        // skip the tail of a token that started in the previous range
        if (p > lo && !is_space(buf[p - 1]))
            while (p < end && !is_space(buf[p])) p++;
        size_t row = tok / 9;
        int field = (int)(tok % 9);
        while (p < end && tok < want_tokens) {
            while (p < end && is_space(buf[p])) p++;
            if (p >= end) break;
            // (a token may run into the next range: it is ours)
            size_t q;
            if (field == 6) {
                q = p + 1;
                if (q < hi && !is_space(buf[q])) { bad[t] = 1; return; }
                dst.otype[row] = (buf[p] == 'P') ? 1 : 0;  // blackscholes.c:761
            } else {
                FP *slot = dst.f[field] ? dst.f[field] + row : nullptr;
                if (!parse_plain<FP>(buf, p, hi, slot, &q)) {
                    q = p;
                    while (q < hi && !is_space(buf[q])) q++;
                    if (!parse_fp<FP>(buf + p, buf + q, slot)) { bad[t] = 1; return; }
                }
            }
            tok++;
            if (++field == 9) { field = 0; row++; }
            p = q;
        }
    });
    for (int t = 0; t < T; t++)
        if (bad[t]) return BS_IO_ERR_INVALID;
    return BS_IO_OK;
}

// The reference's own loop, verbatim in behaviour: used when the fast path met something unusual.
template <typename FP>
int load_with_fscanf(const bs_io_file *f, size_t count, const Dest<FP> &dst)
{
    FILE *file = fopen(f->path.c_str(), "r");
    if (!file) return BS_IO_ERR_OPEN;
    int n = 0;
    if (fscanf(file, "%i", &n) != 1) { fclose(file); return BS_IO_ERR_READ; }
    const char *fmt = sizeof(FP) == 4 ? "%f %f %f %f %f %f %c %f %f" : "%lf %lf %lf %lf %lf %lf %c %lf %lf";
    for (size_t i = 0; i < count; i++) {
        FP v[8];
        char ty;
        const int rv = fscanf(file, fmt, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &ty, &v[6], &v[7]);
        if (rv != 9) { fclose(file); return BS_IO_ERR_READ; }
        for (int k = 0; k < 6; k++)
            if (dst.f[k]) dst.f[k][i] = v[k];
        if (dst.f[7]) dst.f[7][i] = v[6];
        if (dst.f[8]) dst.f[8][i] = v[7];
        dst.otype[i] = (ty == 'P') ? 1 : 0;
    }
    return fclose(file) == 0 ? BS_IO_OK : BS_IO_ERR_CLOSE;
}

template <typename FP>
int load_typed(bs_io_file *f, size_t count, void *spt, void *strike, void *rate, void *vol, void *otime, int *otype,
               void *ref, void *divq, void *divs, int nthreads)
{
    Dest<FP> d;
    d.f[0] = (FP *)spt; d.f[1] = (FP *)strike; d.f[2] = (FP *)rate; d.f[3] = (FP *)divq; d.f[4] = (FP *)vol;
    d.f[5] = (FP *)otime; d.f[6] = nullptr; d.f[7] = (FP *)divs; d.f[8] = (FP *)ref;
    d.otype = otype;
    const int st = load_fast<FP>(f, count, d, nthreads);
    if (st == BS_IO_OK) return st;
    return load_with_fscanf<FP>(f, count, d);
}

// ---- exact "%.18f\n" --------------------------------------------------------------------------
const uint64_t TEN18 = 1000000000000000000ull;

inline char *put_u64(char *p, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// Appends the text printf("%.18f\n", x) would produce; returns the new end.  Needs >= 352 bytes room.
char *format_price(char *p, double x)
{
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int bexp = (int)((bits >> 52) & 0x7ff);
    uint64_t mant = bits & 0xfffffffffffffull;
    if (bexp == 0x7ff || bexp >= 1075 + 10) {  // inf/nan, or |x| >= 2^63: leave it to the C library
        return p + snprintf(p, 352, "%.18f\n", x);
    }
    if (bits >> 63) *p++ = '-';
    int s;  // x = mant * 2^-s
    if (bexp == 0) s = 1074; else { mant |= 1ull << 52; s = 1075 - bexp; }
    uint64_t ipart, frac_num;
    if (s <= 0) { ipart = mant << (-s); frac_num = 0; s = 1; }
    else if (s >= 64) { ipart = 0; frac_num = mant; }
    else { ipart = mant >> s; frac_num = mant & ((1ull << s) - 1); }
    // q = round_half_even(frac_num * 10^18 / 2^s)
    uint64_t q = 0;
    if (frac_num) {
        const unsigned __int128 P = (unsigned __int128)frac_num * TEN18;  // < 2^113
        if (s <= 120) {
            q = (uint64_t)(P >> s);
            const unsigned __int128 rem = P & ((((unsigned __int128)1) << s) - 1);
            const unsigned __int128 half = ((unsigned __int128)1) << (s - 1);
            if (rem > half || (rem == half && (q & 1))) q++;
        }  // else P < 2^113 <= 2^(s-1)/..: strictly below one half -> 0
        if (q == TEN18) { q = 0; ipart++; }
    }
    p = put_u64(p, ipart);
    *p++ = '.';
    char digs[18];
    for (int i = 17; i >= 0; i--) { digs[i] = (char)('0' + q % 10); q /= 10; }
    memcpy(p, digs, 18);
    p += 18;
    *p++ = '\n';
    return p;
}

}  // namespace

extern "C" {

int bs_io_open(const char *path, bs_io_file **file, long long *num_options)
{
    if (!path || !file) return BS_IO_ERR_INVALID;
    *file = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return BS_IO_ERR_OPEN;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); return BS_IO_ERR_OPEN; }
    bs_io_file *f = new (std::nothrow) bs_io_file();
    if (!f) { close(fd); return BS_IO_ERR_NOMEM; }
    f->path = path;
    f->fd = fd;
    f->size = (size_t)st.st_size;
    if (f->size) {
        void *m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); delete f; return BS_IO_ERR_OPEN; }
        f->data = (const char *)m;
        madvise(m, f->size, MADV_WILLNEED);    // advice values are not flags: one call each
        madvise(m, f->size, MADV_SEQUENTIAL);
    }
    if (f->size >= sizeof(SoaHeader) && memcmp(f->data, SOA_MAGIC, 8) == 0) {
        SoaHeader h;
        memcpy(&h, f->data, sizeof(h));
        // Validated by division: a crafted stream_stride must not be able to wrap the size computation.  The
        // stride has to cover the widest stream (fptype, and the int32 otype stream).
        const uint64_t widest = h.fp_bytes == 8 ? 8 : 4;
        const bool stride_ok = f->size >= SOA_ALIGN && h.stream_stride <= (f->size - SOA_ALIGN) / SOA_STREAMS;
        if (h.version != 1 || (h.fp_bytes != 4 && h.fp_bytes != 8) || h.num_options > 2147483647ull || !stride_ok ||
            h.stream_stride < h.num_options * widest) {
            bs_io_close(f);
            return BS_IO_ERR_READ;
        }
        f->soa = true;
        f->body = SOA_ALIGN;
        if (num_options) *num_options = (long long)h.num_options;
        *file = f;
        return BS_IO_OK;
    }
    // header: fscanf("%i") == optional whitespace, then strtol(base 0)   (blackscholes.c:701)
    char head[64];
    size_t p = 0;
    while (p < f->size && is_space(f->data[p])) p++;
    const size_t take = std::min<size_t>(sizeof(head) - 1, f->size - p);
    memcpy(head, f->data ? f->data + p : "", take);
    head[take] = 0;
    char *endp = nullptr;
    const long v = strtol(head, &endp, 0);
    if (endp == head) { bs_io_close(f); return BS_IO_ERR_READ; }
    f->body = p + (size_t)(endp - head);
    if (num_options) *num_options = (long long)(int)v;  // stored into an `int` by the reference
    *file = f;
    return BS_IO_OK;
}

static int load_soa(bs_io_file *f, int fp_bytes, size_t count, void *spt, void *strike, void *rate, void *vol, void *otime,
                    int *otype, void *dgrefval, void *divq, void *divs, int nthreads)
{
    SoaHeader h;
    memcpy(&h, f->data, sizeof(h));
    if ((int)h.fp_bytes != fp_bytes || count > h.num_options) return BS_IO_ERR_READ;
    if (dgrefval && !h.has_dgrefval) return BS_IO_ERR_READ;
    void *dst[SOA_STREAMS] = {spt, strike, rate, vol, otime, otype, dgrefval};
    const int T = host_threads(nthreads, count, 1 << 16);
    parallel_for_chunks(T, [&](int t) {
        const size_t lo = count * t / T, hi = count * (t + 1) / T;
        for (int k = 0; k < SOA_STREAMS; k++) {
            if (!dst[k]) continue;
            const size_t eb = (k == 5) ? sizeof(int) : (size_t)fp_bytes;
            memcpy((char *)dst[k] + lo * eb, f->data + SOA_ALIGN + h.stream_stride * k + lo * eb, (hi - lo) * eb);
        }
    });
    // divq / divs are not kept (the Map never reads them; inputgen writes 0.00): report zeros
    if (divq) memset(divq, 0, count * fp_bytes);
    if (divs) memset(divs, 0, count * fp_bytes);
    return BS_IO_OK;
}

int bs_io_load(bs_io_file *f, int fp_bytes, size_t count, void *spt, void *strike, void *rate, void *vol, void *otime,
               int *otype, void *dgrefval, void *divq, void *divs, int nthreads)
{
    if (!f || (fp_bytes != 4 && fp_bytes != 8)) return BS_IO_ERR_INVALID;
    if (count && (!spt || !strike || !rate || !vol || !otime || !otype)) return BS_IO_ERR_INVALID;
    if (f->soa) return load_soa(f, fp_bytes, count, spt, strike, rate, vol, otime, otype, dgrefval, divq, divs, nthreads);
    if (fp_bytes == 4) return load_typed<float>(f, count, spt, strike, rate, vol, otime, otype, dgrefval, divq, divs, nthreads);
    return load_typed<double>(f, count, spt, strike, rate, vol, otime, otype, dgrefval, divq, divs, nthreads);
}

int bs_io_close(bs_io_file *f)
{
    if (!f) return BS_IO_OK;
    int rc = BS_IO_OK;
    if (f->data) munmap((void *)f->data, f->size);
    if (f->fd >= 0 && close(f->fd) != 0) rc = BS_IO_ERR_CLOSE;
    delete f;
    return rc;
}

int bs_io_write_prices(const char *path, int fp_bytes, size_t count, const void *prices, int nthreads)
{
    if (!path || (fp_bytes != 4 && fp_bytes != 8) || (count && !prices)) return BS_IO_ERR_INVALID;
    FILE *file = fopen(path, "w");
    if (!file) return BS_IO_ERR_OPEN;
    if (fprintf(file, "%i\n", (int)count) < 0) { fclose(file); return BS_IO_ERR_WRITE; }
    const size_t BLOCK = 1u << 20;  // rows formatted per round (bounds memory: ~24 MB of text per round)
    const int T = host_threads(nthreads, std::min(count, BLOCK), 4096);
    std::vector<std::vector<char>> text(T);
    std::vector<size_t> used(T);
    for (size_t base = 0; base < count; base += BLOCK) {
        const size_t rows = std::min(BLOCK, count - base);
        parallel_for_chunks(T, [&](int t) {
            const size_t lo = base + rows * t / T, hi = base + rows * (t + 1) / T;
            std::vector<char> &out = text[t];
            if (out.size() < (hi - lo) * 32 + 512) out.resize((hi - lo) * 32 + 512);
            char *p = out.data();
            for (size_t i = lo; i < hi; i++) {
                const double x = fp_bytes == 4 ? (double)((const float *)prices)[i] : ((const double *)prices)[i];
                if ((size_t)(p - out.data()) + 400 > out.size()) {  // only after a giant snprintf fallback
                    const size_t off = p - out.data();
                    out.resize(out.size() * 2);
                    p = out.data() + off;
                }
                p = format_price(p, x);
            }
            used[t] = p - out.data();
        });
        for (int t = 0; t < T; t++)
            if (used[t] && fwrite(text[t].data(), 1, used[t], file) != used[t]) { fclose(file); return BS_IO_ERR_WRITE; }
    }
    return fclose(file) == 0 ? BS_IO_OK : BS_IO_ERR_CLOSE;
}

int bs_io_is_soa(const bs_io_file *f) { return f && f->soa ? 1 : 0; }

int bs_io_soa_write(const char *path, int fp_bytes, size_t count, const void *spt, const void *strike, const void *rate,
                    const void *vol, const void *otime, const int *otype, const void *dgrefval, const char *source_path)
{
    if (!path || (fp_bytes != 4 && fp_bytes != 8)) return BS_IO_ERR_INVALID;
    if (count && (!spt || !strike || !rate || !vol || !otime || !otype)) return BS_IO_ERR_INVALID;
    SoaHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, SOA_MAGIC, 8);
    h.version = 1;
    h.fp_bytes = (uint32_t)fp_bytes;
    h.num_options = count;
    h.stream_stride = (count * (size_t)fp_bytes + SOA_ALIGN - 1) / SOA_ALIGN * SOA_ALIGN;
    h.has_dgrefval = dgrefval ? 1 : 0;
    if (source_path) {
        struct stat st;
        if (stat(source_path, &st) != 0) return BS_IO_ERR_OPEN;
        h.src_size = (uint64_t)st.st_size;
        h.src_mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    }
    // written under a temporary name and renamed, so a reader never sees a half-written side-car
    const std::string tmp = std::string(path) + ".tmp";
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return BS_IO_ERR_OPEN;
    std::vector<char> head(SOA_ALIGN, 0);
    memcpy(head.data(), &h, sizeof(h));
    bool ok = pwrite(fd, head.data(), SOA_ALIGN, 0) == (ssize_t)SOA_ALIGN;
    const void *src[SOA_STREAMS] = {spt, strike, rate, vol, otime, otype, dgrefval};
    for (int k = 0; k < SOA_STREAMS && ok; k++) {
        const size_t eb = (k == 5) ? sizeof(int) : (size_t)fp_bytes;
        size_t left = src[k] ? count * eb : 0, done = 0;
        while (left && ok) {
            const ssize_t w = pwrite(fd, (const char *)src[k] + done, std::min<size_t>(left, (size_t)1 << 30),
                                     (off_t)(SOA_ALIGN + h.stream_stride * k + done));
            if (w <= 0) ok = false; else { done += (size_t)w; left -= (size_t)w; }
        }
    }
    ok = ok && ftruncate(fd, (off_t)(SOA_ALIGN + h.stream_stride * SOA_STREAMS)) == 0;
    if (close(fd) != 0) ok = false;
    if (!ok) { unlink(tmp.c_str()); return BS_IO_ERR_WRITE; }
    if (rename(tmp.c_str(), path) != 0) { unlink(tmp.c_str()); return BS_IO_ERR_WRITE; }
    return BS_IO_OK;
}

int bs_io_soa_matches(const char *soa_path, const char *source_path, int fp_bytes)
{
    if (!soa_path || !source_path) return 0;
    struct stat st;
    if (stat(source_path, &st) != 0) return 0;
    const int fd = open(soa_path, O_RDONLY);
    if (fd < 0) return 0;
    SoaHeader h;
    const bool got = read(fd, &h, sizeof(h)) == (ssize_t)sizeof(h);
    close(fd);
    if (!got || memcmp(h.magic, SOA_MAGIC, 8) != 0 || h.version != 1 || (int)h.fp_bytes != fp_bytes) return 0;
    const uint64_t mtime = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    return (h.src_size == (uint64_t)st.st_size && h.src_mtime_ns == mtime) ? 1 : 0;
}

/* Exposed for tests: format one value the way the writer does.  Returns the length (without NUL). */
int bs_io_format_price(double x, char *out, size_t cap)
{
    char tmp[400];
    char *e = format_price(tmp, x);
    const size_t n = (size_t)(e - tmp);
    if (n + 1 > cap) return BS_IO_ERR_INVALID;
    memcpy(out, tmp, n);
    out[n] = 0;
    return (int)n;
}

}  // extern "C"
