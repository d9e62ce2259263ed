/*
 * bs_libm_f64.h -- exp / log (double) that return what the reference's libm returns, bit for bit.
 *
 * The reference's fptype=double build (BASELINE.json configs[2]) calls glibc's exp three times per option and log once --
 * blackscholes.c:152,226,248; with sqrt (correctly rounded) they are the only arithmetic of the path that lives in a
 * third-party dependency.  Every other operation of that build is an individually rounded IEEE double
 * operation, which the IEEE-order kernel reproduces exactly (price_f64 in bs_kernels.cuh); with libdevice's exp/log it is
 * bit-identical to the CPU on 89-100 % of rows and one ulp of exp or log away on the rest.  BS_MATH_REFERENCE with
 * fp_bytes = 8 uses the two functions below instead and returns the reference CPU prices bit for bit.
 *
 * Dependency restated: GNU libc 2.39 (the image's libm), sysdeps/ieee754/dbl-64/e_exp.c and e_log.c -- the algorithms of
 * ARM's optimized-routines adopted in glibc 2.28:
 *   exp: x = k ln2/128 + r; 2^(k/128) = scale (1 + tail) from a 128-entry table; exp(r) - 1 by a degree-5 polynomial;
 *        result scale + scale * tmp; |x| >= 512 through specialcase() (scale split so that it cannot over/underflow),
 *        |x| < 2^-54 as 1 + x;
 *   log: x = 2^k z, z in [0x1.6p-1, 0x1.6p0), 128-entry table {1/c, log c}, r = z/c - 1 by one fma, hi + lo splitting of
 *        k ln2 + log c + r, degree-6 polynomial; arguments within [1 - 2^-4, 1 + 0x1.09p-4) by a degree-12 polynomial in
 *        x - 1 with a Dekker split of the r^2/2 term.
 * The operation ORDER AND FUSION below is that of the code x86-64 glibc actually runs on a CPU with FMA3 + AVX2 (the
 * ifunc variants __exp_fma / __log_fma, read from the image's libm with objdump): which additions are contracted into
 * fma is visible in the last bit.  The tables are read out of the same libm (tools/gen_libm_f64_tables.py).
 * PINNED: tools/libm_f64_host_check.cpp compares this header, compiled for the host, with the running libm on 2 x 10^9
 * arguments per function (all exponents, the |x| >= 512 and subnormal-result ranges of exp, the near-1 range of log,
 * zeros / infinities / NaNs / subnormals): 0 mismatches on glibc 2.39; tests/test_libm_f64.py runs a subset.  On the
 * device the same double operations are IEEE-exact (DFMA / DMUL / DADD with explicit rounding), so the device results
 * are the host results.
 *
 * Host-or-device code, like bs_libm_f32.h.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "bs_libm_f64_tables.h"

#if defined(__CUDACC__)
#define BSL64_HD __device__ __forceinline__
#else
#define BSL64_HD static inline
#endif

namespace bsl64 {

BSL64_HD uint64_t d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
BSL64_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
// individually rounded operations; the only contractions are the fma() calls written out below
BSL64_HD double mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host builds of this header use -ffp-contract=off
#endif
}
BSL64_HD double add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
BSL64_HD double sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
BSL64_HD double fmad(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// e_exp.c: specialcase() -- the exponent of `scale` may have left the normal range (512 <= |x| < 1024)
BSL64_HD double exp_specialcase(double tmp, uint64_t sbits, uint64_t ki)
{
    if ((ki & 0x80000000ull) == 0) {
        // k > 0: the exponent of scale might have overflowed by <= 460
        sbits -= 1009ull << 52;
        const double scale = u2d(sbits);
        return mul(u2d(0x7f00000000000000ull), fmad(scale, tmp, scale));  // 0x1p1009 * (scale + scale * tmp); inf on overflow
    }
    // k < 0: special care in the subnormal range
    sbits += 1022ull << 52;
    const double scale = u2d(sbits);
    const double p = mul(scale, tmp);  // the product is shared by y and lo, so it is rounded on its own (no fma here)
    double y = add(scale, p);
    if (y < 1.0) {
        // round y to the right precision before scaling it into the subnormal range (avoids double rounding)
        double lo = add(sub(scale, y), p);
        const double hi = add(1.0, y);
        lo = add(add(sub(1.0, hi), y), lo);
        y = sub(add(hi, lo), 1.0);
        if (y == 0.0) y = 0.0;  // no -0.0
    }
    return mul(u2d(0x0010000000000000ull), y);  // 0x1p-1022 * y
}

// glibc 2.39 exp (__exp_fma)
BSL64_HD double exp_glibc(double x)
{
    const uint64_t ix = d2u(x);
    uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {  // top12(0x1p-54) = 0x3c9, top12(512.0) = 0x408
        if (abstop - 0x3c9u >= 0x80000000u) return add(1.0, x);  // |x| < 2^-54 (also 0 and subnormals): 1 + x
        if (abstop >= 0x409u) {                                  // |x| >= 1024, inf, NaN
            if (ix == 0xfff0000000000000ull) return 0.0;         // exp(-inf)
            if (abstop >= 0x7ffu) return add(1.0, x);            // +inf, NaN
            return (ix >> 63) ? 0.0 : u2d(0x7ff0000000000000ull);  // underflow / overflow
        }
        abstop = 0;  // 512 <= |x| < 1024: the large case is handled below
    }
    const double INVLN2N = u2d(EXP_CONST[0]), SHIFT = u2d(EXP_CONST[1]), NEGLN2HIN = u2d(EXP_CONST[2]), NEGLN2LON = u2d(EXP_CONST[3]);
    const double C2 = u2d(EXP_CONST[4]), C3 = u2d(EXP_CONST[5]), C4 = u2d(EXP_CONST[6]), C5 = u2d(EXP_CONST[7]);
    // exp(x) = 2^(k/N) exp(r), x = ln2/N k + r, r in [-ln2/2N, ln2/2N]
    double kd = fmad(x, INVLN2N, SHIFT);   // z + Shift, contracted
    const uint64_t ki = d2u(kd);
    kd = sub(kd, SHIFT);
    double r = fmad(kd, NEGLN2HIN, x);
    r = fmad(kd, NEGLN2LON, r);
    // 2^(k/N) ~= scale (1 + tail)
    const uint32_t idx = 2 * (uint32_t)(ki & 127);
    const uint64_t top = ki << 45;
    const double tail = u2d(EXP_TAB[idx]);
    const uint64_t sbits = EXP_TAB[idx + 1] + top;
    const double r2 = mul(r, r);
    // tmp = tail + r + r2 (C2 + r C3) + r2 r2 (C4 + r C5)
    double tmp = fmad(fmad(r, C3, C2), r2, add(r, tail));
    tmp = fmad(mul(r2, r2), fmad(r, C5, C4), tmp);
    if (abstop == 0) return exp_specialcase(tmp, sbits, ki);
    const double scale = u2d(sbits);
    return fmad(scale, tmp, scale);
}

// glibc 2.39 log (__log_fma)
BSL64_HD double log_glibc(double x)
{
    uint64_t ix = d2u(x);
    // LO = asuint64(1.0 - 0x1p-4), HI = asuint64(1.0 + 0x1.09p-4)
    if (ix - 0x3fee000000000000ull < 0x3ff1090000000000ull - 0x3fee000000000000ull) {
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double B0 = u2d(LOG_CONST[7]), B1 = u2d(LOG_CONST[8]), B2 = u2d(LOG_CONST[9]), B3 = u2d(LOG_CONST[10]), B4 = u2d(LOG_CONST[11]),
                     B5 = u2d(LOG_CONST[12]), B6 = u2d(LOG_CONST[13]), B7 = u2d(LOG_CONST[14]), B8 = u2d(LOG_CONST[15]), B9 = u2d(LOG_CONST[16]),
                     B10 = u2d(LOG_CONST[17]);
        const double r = sub(x, 1.0);
        const double r2 = mul(r, r);
        const double r3 = mul(r, r2);
        // y = r3 (B1 + r B2 + r2 B3 + r3 (B4 + r B5 + r2 B6 + r3 (B7 + r B8 + r2 B9 + r3 B10)))   -- r3 * applied last, fused
        const double p1 = fmad(r2, B3, fmad(r, B2, B1));
        const double p2 = fmad(r2, B6, fmad(r, B5, B4));
        double p3 = fmad(r2, B9, fmad(r, B8, B7));
        p3 = fmad(r3, B10, p3);
        double p = fmad(p3, r3, p2);
        p = fmad(p, r3, p1);
        // Dekker split of r: rhi = r + w - w with w = r 2^27 (both sums contracted)
        const double TWO27 = 134217728.0;
        const double rw = fmad(r, TWO27, r);
        const double rhi = fmad(-TWO27, r, rw);
        const double rlo = sub(r, rhi);
        const double rhi2 = mul(rhi, rhi);
        const double hi = fmad(rhi2, B0, r);           // r + rhi^2 B0
        double lo = fmad(rhi2, B0, sub(r, hi));        // r - hi + rhi^2 B0
        lo = fmad(mul(B0, rlo), add(r, rhi), lo);      // lo += B0 rlo (rhi + r)
        const double y = fmad(p, r3, lo);              // r3 p + lo
        return add(hi, y);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return u2d(0xfff0000000000000ull);                // log(+-0) = -inf
        if (ix == 0x7ff0000000000000ull) return x;                         // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return u2d(0x7ff8000000000000ull);  // negative or NaN: NaN
        ix = d2u(mul(x, 4503599627370496.0));                              // subnormal: normalise by 2^52
        ix -= 52ull << 52;
    }
    const double LN2HI = u2d(LOG_CONST[0]), LN2LO = u2d(LOG_CONST[1]);
    const double A0 = u2d(LOG_CONST[2]), A1 = u2d(LOG_CONST[3]), A2 = u2d(LOG_CONST[4]), A3 = u2d(LOG_CONST[5]), A4 = u2d(LOG_CONST[6]);
    // x = 2^k z, z in [OFF, 2 OFF), OFF = 0x3fe6000000000000
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = u2d(LOG_TAB[2 * i]), logc = u2d(LOG_TAB[2 * i + 1]);
    const double z = u2d(iz);
    const double r = fmad(z, invc, -1.0);
    const double kd = (double)k;
    // hi + lo = r + log(c) + k ln2
    const double w = fmad(kd, LN2HI, logc);
    const double hi = add(w, r);
    const double lo = fmad(kd, LN2LO, add(sub(w, hi), r));
    const double r2 = mul(r, r);
    // y = lo + r2 A0 + r r2 (A1 + r A2 + r2 (A3 + r A4)) + hi
    const double q = fmad(fmad(r, A4, A3), r2, fmad(r, A2, A1));
    const double y = fmad(mul(r, r2), q, fmad(r2, A0, lo));
    return add(y, hi);
}

}  // namespace bsl64
