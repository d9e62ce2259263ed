/*
 * bs_math_f64.h -- branch-free fp64 building blocks of the BS_MATH_FAST fp64 kernel.
 *
 * The fp64 Map is instruction-bound, not HBM-bound, when built from libdevice's exp()/log() and IEEE-rounded
 * divide/sqrt (412 SASS instructions per option, profiles/r01_ncu_f64_u2_444x256.txt).  These replacements keep
 * every result within ~2 ulp but drop the special-case handling the pricing formula cannot reach:
 *
 *   rcp_f64      MUFU.RCP64H seed (>= 19 good bits) + ONE cubic Newton step  x(1 + e + e^2)   -> 4 ops
 *   rsqrt_f64    MUFU.RSQ64H seed + ONE cubic step  y(1 + e/2 + 3e^2/8)                        -> 6 ops
 *   exp_f64      x = (64k + j) ln2/64 + r, |r| <= ln2/128: 2^k * T[j] * (Taylor degree 5); 0 below -708  -> 11 ops
 *   log_f64      x = 2^e m; r = m * RC[i] - 1 with i = top 6 mantissa bits, |r| < 2^-7:
 *                (e ln2 + LC[i]) + log1p(r) (degree 7)                                           -> 12 ops
 * The two 64-entry tables (bs_tables_f64.h, generated with 60-digit arithmetic by tools/gen_tables_f64.py) sit in
 * shared memory on the device (Fp64Tables, 1.5 KB, filled once per CTA).
 *
 * Written as host-or-device code (device under nvcc, host under g++): on the host the hardware seeds are emulated (a reciprocal truncated to 20
 * mantissa bits), so tests/test_math_f64.py can measure the ulp error of every block against libm and of the
 * whole price against the oracle WITHOUT a GPU (tools/math_f64_host_check.cpp).
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "bs_tables_f64.h"

#if defined(__CUDACC__)
#define BS_HD __device__ __forceinline__  /* under nvcc: device code only (the tables are __constant__) */
#else
#define BS_HD inline                      /* plain C++: the host emulation used by the CPU tests */
#endif

namespace bsm {

BS_HD double from_bits(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
BS_HD uint64_t to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}

// Scalar constant i of bs_tables_f64.h (a __constant__ bank operand / LDCU on the device).
BS_HD double kd(int i) { return from_bits(KD_BITS[i]); }

// ---- hardware seeds (emulated on the host with the same ~2^-20 accuracy) ---------------------------
BS_HD double seed_rcp(double b)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    return x;
#else
    return from_bits(to_bits(1.0 / from_bits(to_bits(b) & 0xffffffff00000000ull)) & 0xffffffff00000000ull);
#endif
}
BS_HD double seed_rsqrt(double b)
{
#if defined(__CUDA_ARCH__)
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    return x;
#else
    return from_bits(to_bits(1.0 / sqrt(from_bits(to_bits(b) & 0xffffffff00000000ull))) & 0xffffffff00000000ull);
#endif
}

// 1/b for normal b: seed error e0 <= ~2^-19, after x(1+e+e^2) the error is e0^3 ~ 2^-57 plus rounding.
BS_HD double rcp_f64(double b)
{
    double x = seed_rcp(b);
    double e = fma(-b, x, 1.0);
    double c = fma(e, e, e);
    return fma(x, c, x);
}

// 1/sqrt(t) for normal t > 0: e = 1 - t y^2, y (1 + e/2 + 3 e^2 / 8), error (5/16) e^3 ~ 2^-59 plus rounding.
BS_HD double rsqrt_f64(double t)
{
    double y = seed_rsqrt(t);
    double a = t * y;
    double e = fma(-a, y, 1.0);
    double c = fma(0.375, e, 0.5);
    double ye = y * e;
    return fma(ye, c, y);
}

// Layout of the table block handed to exp_f64 / log_f64: [0,64) = 2^(j/64); then 64 pairs {1/c_i, log c_i (- ln2 for
// i >= LOG_SPLIT)}.  On the device it lives in shared memory; on the host in a static array.
enum { TAB_EXP = 0, TAB_LOG = 64, TAB_DOUBLES = 64 + 128 };

BS_HD void fill_tables(double *tab, int first, int step)
{
    for (int i = first; i < TAB_DOUBLES; i += step)
        tab[i] = from_bits(i < TAB_LOG ? EXP2_64_BITS[i] : LOG_RC_LC_BITS[(i - TAB_LOG) >> 1][(i - TAB_LOG) & 1]);
}

// exp(x) for x <= ~700 (the Map only needs x <= 0): exact zero below -708 (no subnormal results).
BS_HD double exp_f64(double x, const double *tab)
{
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51: rounds to nearest integer in the low word
    double nd = fma(x, kd(K_EXP_INV), MAGIC);
    const int n = (int)(uint32_t)to_bits(nd);  // round(x * 64/ln2) = 64 k + j
    nd -= MAGIC;
    double r = fma(nd, -kd(K_EXP_HI), x);
    r = fma(nd, -kd(K_EXP_LO), r);                // |r| <= ln2/128: r^6/720 < 4e-17
    // expm1(r) = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120); the result is T + T*expm1(r) in one fma, so only the
    // table entry's and the final rounding (<= 1 ulp together) reach the result
    double q = fma(kd(K_EXP_C5), r, kd(K_EXP_C4));
    q = fma(q, r, kd(K_EXP_C3));
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);
    const double T = tab[TAB_EXP + (n & 63)];
    const double m = fma(T, em1, T);              // in [0.99, 2.0)
    const uint64_t scaled = to_bits(m) + ((uint64_t)(int64_t)(n >> 6) << 52);
    return x < -708.0 ? 0.0 : from_bits(scaled);
}

// log(x) for normal x > 0.  Absolute error ~1e-16 (relative to max(1, |log x|)); arguments close to 1 keep a small
// relative error because the table entry of their interval is already reduced by ln 2 (LOG_SPLIT).
BS_HD double log_f64(double x, const double *tab)
{
    const uint64_t b = to_bits(x);
    const int i = (int)(b >> 46) & 63;            // top 6 mantissa bits: m in [1 + i/64, 1 + (i+1)/64)
    const int e = (int)(b >> 52) - 1023 + (i >= LOG_SPLIT ? 1 : 0);
    const double m = from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    const double rc = tab[TAB_LOG + 2 * i], lc = tab[TAB_LOG + 2 * i + 1];
    const double r = fma(m, rc, -1.0);            // |r| < 2^-7: r^8/8 < 2e-18
    double q = fma(kd(K_LOG_C7), r, kd(K_LOG_C6));  // r - r^2/2 + r^3/3 - ... + r^7/7
    q = fma(q, r, kd(K_LOG_C5));
    q = fma(q, r, -0.25);
    q = fma(q, r, kd(K_LOG_C3));
    q = fma(q, r, -0.5);
    const double lp = fma(r * r, q, r);           // log1p(r)
    const double ed = (double)e;
    return fma(ed, kd(K_LN2_HI), lc + fma(ed, kd(K_LN2_LO), lp));
}

// 1 - N(|d|) given k = 1/(1 + 0.2316419|d|): n(d) poly(k), constants of CNDF (blackscholes.c:126,:156,:164-170)
// pre-multiplied by 1/sqrt(2 pi).
BS_HD double cndf_tail_f64(double d, double k, const double *tab)
{
    double e = exp_f64((-0.5 * d) * d, tab);
    double p = fma(k, kd(K_CNDF_A5), kd(K_CNDF_A4));
    p = fma(k, p, kd(K_CNDF_A3));
    p = fma(k, p, kd(K_CNDF_A2));
    p = fma(k, p, kd(K_CNDF_A1));
    return (p * k) * e;
}

// The whole option (BlkSchlsEqEuroNoDiv, blackscholes.c:190-258) for s, k, v, t > 0 finite.  `ok` is false for
// degenerate inputs (den = v sqrt(t) not a positive normal number): the caller then uses the IEEE-order path.
BS_HD double price_f64_fast(double s, double k, double r, double v, double t, int otype, bool *ok, const double *tab)
{
    const double y = rsqrt_f64(t);           // 1/sqrt(t)
    const double sq = t * y;                 // sqrt(t)                       :224
    const double den = v * sq;               // xDen                          :238
    const double rkv = rcp_f64(k * v);       // one reciprocal serves 1/k and 1/v
    const double inv_k = rkv * v, inv_v = rkv * k;
    const double rden = y * inv_v;           // 1/(v sqrt t)
    const double lg = log_f64(s * inv_k, tab);  // log(s/k)                   :226
    const double drift = fma(0.5 * v, v, r); // r + v^2/2                     :231-234
    const double d1 = fma(drift, t, lg) * rden;  //                           :235-239
    const double d2 = d1 - den;              //                               :240
    const double fv = k * exp_f64(-r * t, tab);  // strike exp(-r t)          :248
    const double a1 = fma(fabs(d1), kd(K_CNDF_C), 1.0), a2 = fma(fabs(d2), kd(K_CNDF_C), 1.0);
    const double rab = rcp_f64(a1 * a2);     // one reciprocal serves both CNDF arguments   :156-158
    const double w1 = cndf_tail_f64(d1, rab * a2, tab);
    const double w2 = cndf_tail_f64(d2, rab * a1, tab);
    const bool put = otype != 0;
    // N(x) = w for x < 0 and 1-w otherwise; a put needs N(-x)                :249-255
    const double x1 = ((d1 < 0.0) != put) ? w1 : 1.0 - w1;
    const double x2 = ((d2 < 0.0) != put) ? w2 : 1.0 - w2;
    const double c = fma(s, x1, -(fv * x2));
    // Domain of the blocks above: t, k*v and s/k positive normal numbers well inside the exponent range
    // (biased exponent in [0x100, 0x6ff], i.e. 2^-767 .. 2^768); the product a1*a2 = (1 + c|d1|)(1 + c|d2|) fed to the
    // shared reciprocal finite (it overflows for v below ~1e-154, where |d| ~ 1/v: rcp(inf) = 0 and 0 * inf = NaN
    // where the reference still returns the intrinsic value); and |r t| < 2^9, so exp_f64's exponent arithmetic cannot
    // wrap (the reference returns inf / NaN there, exp_f64 a finite number).  Integer range checks on the high
    // words; everything else (t = 0, v = 0, s <= 0, NaN, inf, denormals, overflow) is left to the IEEE-order path.
    const uint32_t LO = 0x10000000u, SPAN = 0x60000000u;
    const uint32_t hi_a = (uint32_t)(to_bits(a1 * a2) >> 32), hi_rt = (uint32_t)(to_bits(r * t) >> 32) & 0x7fffffffu;
    *ok = ((uint32_t)(to_bits(t) >> 32) - LO < SPAN) && ((uint32_t)(to_bits(k * v) >> 32) - LO < SPAN) &&
          ((uint32_t)(to_bits(s * inv_k) >> 32) - LO < SPAN) && (hi_a < 0x6ff00000u) && (hi_rt < 0x40800000u);
    return put ? -c : c;
}

}  // namespace bsm
