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

// m * 2^k for a normal m and a normal result, as an integer add on the high word only (one IADD on the device; the
// 64-bit form costs an add with carry on a register pair).
BS_HD double scale_by_pow2(double m, int k)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(m) + (k << 20), __double2loint(m));
#else
    return from_bits(to_bits(m) + ((uint64_t)(int64_t)k << 52));
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
    double c = fma(kd(K_RSQ_C2), e, 0.5);     // 0.375 from the constant bank: DFMA takes one immediate, and 0.5 is it
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

// The blackscholes fp64 kernel's own, four times finer tables (6 KB): [0,256) = 2^(j/256); then 256 pairs {1/c_i, log c_i
// (- ln2 for i >= LOG256_SPLIT)}.  Finer intervals buy shorter polynomials -- degree 4 instead of 5 for exp (|r| <= ln2/512:
// r^5/120 < 4e-17), degree 5 instead of 7 for log (|r| < 2^-9: r^6/6 < 1e-17) -- i.e. four FP64 instructions less per option
// at the same accuracy; the kernel is FP64-issue- and, sustained, power-bound.
// Layout: [0,512) the 256 log pairs (4 KB), [512,768) 2^(j/256) (2 KB).  On the device the block must start on a 4 KB
// boundary of shared memory (TAB256_ALIGN; both kernels place it so with place_tables256): each sub-table is then aligned to
// its own size and an entry's address is `base | index bits` -- one LOP3 on the bits as they come out of the argument,
// instead of mask, shift and add (BS_F64_TAB_OR; 7 integer instructions less per option of ~95).
enum { TAB256_LOG = 0, TAB256_EXP = 512, TAB256_DOUBLES = 512 + 256, TAB256_ALIGN = 4096 };
#ifndef BS_F64_TAB_OR
#define BS_F64_TAB_OR 1
#endif

// With BS_F64_TAB_OR the device's copy of entry j of the exp table is 2^(j/256) with j * 2^12 subtracted from its high
// word: exp256_core_f64 then adds n * 2^12 = (256 k + j) * 2^12 to that word, which puts k into the exponent field and takes
// the bias out again -- the index bits need not be masked off n first.
BS_HD void fill_tables256(double *tab, int first, int step)
{
    for (int i = first; i < TAB256_DOUBLES; i += step) {
        uint64_t b = i < TAB256_EXP ? LOG256_RC_LC_BITS[i >> 1][i & 1] : EXP2_256_BITS[i - TAB256_EXP];
#if defined(__CUDA_ARCH__) && BS_F64_TAB_OR
        if (i >= TAB256_EXP) b -= (uint64_t)(i - TAB256_EXP) << 44;
#endif
        tab[i] = from_bits(b);
    }
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

// Where log_f64 reads its table pair {1/c_i, log c_i} from.  PlainLogTab: the block filled by fill_tables.  Kernels whose
// lanes look up unrelated intervals at the same time may pass a bank-replicated copy instead (sw_kernels.cuh).
struct PlainLogTab {
    const double *tab;
    BS_HD void pair(int i, double &rc, double &lc) const
    {
        rc = tab[TAB_LOG + 2 * i];
        lc = tab[TAB_LOG + 2 * i + 1];
    }
};

// log(x) for normal x > 0.  Absolute error ~1e-16 (relative to max(1, |log x|)); arguments close to 1 keep a small
// relative error because the table entry of their interval is already reduced by ln 2 (LOG_SPLIT).
template <class LT>
BS_HD double log_f64_t(double x, const LT &lt)
{
    const uint64_t b = to_bits(x);
    const int i = (int)(b >> 46) & 63;            // top 6 mantissa bits: m in [1 + i/64, 1 + (i+1)/64)
    // exponent, + 1 from LOG_SPLIT on: adding 64 - LOG_SPLIT to the index bits carries into the exponent field exactly then
    const int e = (int)(((uint32_t)(b >> 32) + ((64u - LOG_SPLIT) << 14)) >> 20) - 1023;
    const double m = from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    double rc, lc;
    lt.pair(i, rc, lc);
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
BS_HD double log_f64(double x, const double *tab)
{
    const PlainLogTab lt = {tab};
    return log_f64_t(x, lt);
}

// exp(x) without exp_f64's underflow guard, for callers that have range-checked x themselves (|x| < 700); the
// exponent is added to the high word only (one integer add instead of a 64-bit add with carry).
BS_HD double exp_core_f64(double x, const double *tab)
{
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51
    double nd = fma(x, kd(K_EXP_INV), MAGIC);
    const int n = (int)(uint32_t)to_bits(nd);  // round(x * 64/ln2) = 64 k + j
    nd -= MAGIC;
    double r = fma(nd, -kd(K_EXP_HI), x);
    r = fma(nd, -kd(K_EXP_LO), r);
    double q = fma(kd(K_EXP_C5), r, kd(K_EXP_C4));
    q = fma(q, r, kd(K_EXP_C3));
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);
    const double T = tab[TAB_EXP + (n & 63)];
    return scale_by_pow2(fma(T, em1, T), n >> 6);  // the mantissa product is in [0.99, 2.0)
}

// exp(x), |x| < 700, from the 256-entry table: 9 FP64 instructions.  `tab` is the block filled by fill_tables256.
BS_HD double exp256_core_f64(double x, const double *tab)
{
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51
    double nd = fma(x, kd(K_EXP256_INV), MAGIC);
    const int n = (int)(uint32_t)to_bits(nd);  // round(x * 256/ln2) = 256 k + j
    nd -= MAGIC;
    double r = fma(nd, -kd(K_EXP256_HI), x);
    r = fma(nd, -kd(K_EXP256_LO), r);            // |r| <= ln2/512
    double q = fma(kd(K_EXP_C4), r, kd(K_EXP_C3));
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);         // r + r^2/2 + r^3/6 + r^4/24
#if defined(__CUDA_ARCH__) && BS_F64_TAB_OR
    // biased entry j at (base + 4 KB) | (j << 3); its high word + n * 2^12 is the high word of 2^k 2^(j/256) (fill_tables256):
    // one LOP3 for the address, one IMAD for the scaling, and the scaled T goes through the final FMA
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + TAB256_EXP * 8;
    int lo, hi;
    asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(base | (((uint32_t)n << 3) & 0x7f8u)));
    asm("mad.lo.s32 %0, %1, 4096, %0;" : "+r"(hi) : "r"(n));    // in place: the loaded pair becomes T
    const double T = __hiloint2double(hi, lo);                   // 2^k 2^(j/256), normal for |x| < 700
    return fma(T, em1, T);
#else
    const double T = tab[TAB256_EXP + (n & 255)];
    return scale_by_pow2(fma(T, em1, T), n >> 8);
#endif
}

// log(x) for normal x > 0 from the 256-entry table: 9 FP64 instructions.
BS_HD double log256_f64(double x, const double *tab)
{
    const uint64_t b = to_bits(x);
    const double m = from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
#if defined(__CUDA_ARCH__) && BS_F64_TAB_OR
    // pair i at base | (i << 4), i = the top 8 mantissa bits = bits 12..19 of the high word: (hi >> 8) & 0xff0
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + TAB256_LOG * 8;
    const uint32_t hw = (uint32_t)(b >> 32);
    const uint32_t addr = base | ((hw >> 8) & 0xff0u);
    double rc, lc;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rc), "=d"(lc) : "r"(addr));
    // e + 1 from LOG256_SPLIT on: adding (256 - SPLIT) to the index bits carries into the exponent field exactly then
    const int e = (int)((hw + ((256u - LOG256_SPLIT) << 12)) >> 20) - 1023;
#else
    const int i = (int)(b >> 44) & 255;          // top 8 mantissa bits: m in [1 + i/256, 1 + (i+1)/256)
    const int e = (int)(b >> 52) - 1023 + (i >= LOG256_SPLIT ? 1 : 0);
    const double rc = tab[TAB256_LOG + 2 * i], lc = tab[TAB256_LOG + 2 * i + 1];
#endif
    const double r = fma(m, rc, -1.0);           // |r| < 2^-9
    double q = fma(kd(K_LOG_C5), r, -0.25);      // r - r^2/2 + r^3/3 - r^4/4 + r^5/5
    q = fma(q, r, kd(K_LOG_C3));
    q = fma(q, r, -0.5);
    const double lp = fma(r * r, q, r);          // log1p(r)
    const double ed = (double)e;
    return fma(ed, kd(K_LN2_HI), lc + fma(ed, kd(K_LN2_LO), lp));
}

// Which tables the blackscholes fp64 kernel uses: the 256-entry ones (default) or, for A/B measurements on one board
// (-DBS_F64_TAB64=1, lib/libbs_gpu_tab64.so), the 64-entry ones it shares with the swaptions kernels.
#ifndef BS_F64_TAB64
#define BS_F64_TAB64 0
#endif
#if BS_F64_TAB64
#define BS_F64_LOG log_f64
#define BS_F64_EXP exp_core_f64
#define BS_F64_FILL_TABLES fill_tables
enum { BS_F64_TAB_DOUBLES = TAB_DOUBLES, BS_F64_TAB_PAD = 0 };
#else
#define BS_F64_LOG log256_f64
#define BS_F64_EXP exp256_core_f64
#define BS_F64_FILL_TABLES fill_tables256
enum { BS_F64_TAB_DOUBLES = TAB256_DOUBLES, BS_F64_TAB_PAD = BS_F64_TAB_OR ? TAB256_ALIGN : 0 };
#endif
// Where the kernel's table block goes inside `raw` (BS_F64_TAB_DOUBLES * 8 + BS_F64_TAB_PAD bytes of shared memory): the
// first address whose offset in the shared window is a multiple of the alignment the `base | index` addressing needs.
#if defined(__CUDACC__)
__device__ __forceinline__ double *place_tables(void *raw)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(raw);
    return reinterpret_cast<double *>(static_cast<unsigned char *>(raw) + (BS_F64_TAB_PAD ? ((0u - a) & (uint32_t)(BS_F64_TAB_PAD - 1)) : 0u));
}
#endif

// p with its sign flipped unless the sign bit of x is set: one LOP3 on the high word (p_hi ^ (~x & SIGN)).
BS_HD double flip_sign_unless(double p, uint32_t x)
{
    return from_bits(to_bits(p) ^ ((uint64_t)(~x & 0x80000000u) << 32));
}
// a for x >= 0, (+)0 for x < 0.  On the device only the high word is selected: what remains for x < 0 is a denormal below
// 2^-1042 instead of an exact zero -- it enters the price through one addition, 300 orders of magnitude below the
// tolerance, and saves the second FSEL of a 64-bit select (the kernel is bound by instruction issue).
BS_HD double zero_if_negative(double a, int x)
{
#if defined(__CUDA_ARCH__) && BS_F64_TAB_OR
    return __hiloint2double(x < 0 ? 0 : __double2hiint(a), __double2loint(a));
#else
    return x < 0 ? 0.0 : a;
#endif
}

// poly(k) with k = 1/(1 + 0.2316419|d|): the CNDF tail 1 - N(|d|) is exp(-d^2/2) k poly(k) / sqrt(2 pi); the constants
// of CNDF (blackscholes.c:126,:156,:164-170) are pre-multiplied by 1/sqrt(2 pi).
BS_HD double cndf_poly_f64(double k)
{
    double p = fma(k, kd(K_CNDF_A5), kd(K_CNDF_A4));
    p = fma(k, p, kd(K_CNDF_A3));
    p = fma(k, p, kd(K_CNDF_A2));
    p = fma(k, p, kd(K_CNDF_A1));
    return p * k;
}

// 1 - N(|d|) given k = 1/(1 + 0.2316419|d|)
BS_HD double cndf_tail_f64(double d, double k, const double *tab)
{
    return cndf_poly_f64(k) * exp_f64((-0.5 * d) * d, tab);
}

// The whole option (BlkSchlsEqEuroNoDiv, blackscholes.c:190-258) for s, k, v, t > 0 finite.  `ok` is false for
// inputs outside the domain of the blocks below: the caller then uses the IEEE-order path.
//
// Restructured for the fp64 pipe -- the kernel is bound by FP64 issue, not by HBM (profiles/r02_ncu_f64_*.txt):
//   * xDen and 1/xDen from ONE rsqrt of xDen^2 = v^2 t (sqrt t itself is never needed), and the numerator of d1 as
//     r t + xDen^2 / 2 + log(s/k) with the r t that exp(-r t) needs anyway; s/k through a 3-instruction reciprocal;
//   * the second CNDF exponential is never evaluated.  With P_j = k_j poly(k_j) the two tails are
//     w1 = e1 P1 and w2 = e2 P2, e_j = exp(-d_j^2/2), and the Black-Scholes identity  s n(d1) = fv n(d2)
//     (d1 den - den^2/2 = log(s/k) + r t exactly) gives  fv w2 = s e1 P2.  Hence, with g = s e1,
//         s N1 - fv N2 = g (+-P1 -+ P2) + ([N1 = 1 - w1] s - [N2 = 1 - w2] fv):
//     one exponential, signs and the two optional terms selected by integer masks on the sign bits of d1, d2;
//   * one shared reciprocal for the two CNDF arguments (as before).
// With the 256-entry tables (exp256_core_f64, log256_f64): 69 fp64 operations per option (94 in round 1).  Rounding errors of d1 enter the price only multiplied by den (the
// common shift of d1 and d2 cancels in the identity), i.e. at the 1e-15 level: measured 8e-14 worst absolute
// distance to the fp64 CPU build on the inputgen table (tests/test_math_f64.py, tests/test_gpu_parity.py).
BS_HD double price_f64_fast(double s, double k, double r, double v, double t, int otype, bool *ok, const double *tab)
{
    const double vv = v * v;
    const double z = vv * t;                     // xDen^2 = v^2 t
    const double rden = rsqrt_f64(z);            // 1/(v sqrt t)   (v > 0: checked below)
    const double den = z * rden;                 // xDen = v sqrt t               :224,:238
    const double sk = s * rcp_f64(k);            // s/k
    const double lg = BS_F64_LOG(sk, tab);       // log(s/k)                      :226   (tab: BS_F64_FILL_TABLES)
    const double rt = r * t;
    const double d1 = (fma(0.5, z, rt) + lg) * rden;  // ((r + v^2/2) t + log(s/k)) / xDen   :231-239
    const double d2 = d1 - den;                  //                               :240
    const double fv = k * BS_F64_EXP(-rt, tab);  // strike exp(-r t)              :248
    const double a1 = fma(fabs(d1), kd(K_CNDF_C), 1.0), a2 = fma(fabs(d2), kd(K_CNDF_C), 1.0);
    const double a12 = a1 * a2;
    const double rab = rcp_f64(a12);             // one reciprocal serves both CNDF arguments   :156-158
    const double q1 = (-0.5 * d1) * d1;          // -d1^2/2
    const double g = s * BS_F64_EXP(q1, tab);    // s exp(-d1^2/2): s n(d1) = fv n(d2), up to 1/sqrt(2 pi)
    const double P1 = cndf_poly_f64(rab * a2), P2 = cndf_poly_f64(rab * a1);
    // N(x) = tail w for x < 0 and 1 - w otherwise; a put needs N(-x): the tail itself is wanted when sign(d) != put.   :249-255
    // Sign bit of x_j = high word of d_j XOR put flag: set <=> the tail itself is wanted.
    const uint32_t SIGN = 0x80000000u;
    const uint32_t puth = otype != 0 ? SIGN : 0u;
    const uint32_t x1 = (uint32_t)(to_bits(d1) >> 32) ^ puth, x2 = (uint32_t)(to_bits(d2) >> 32) ^ puth;
    const double t1 = flip_sign_unless(P1, x1);             // +P1 for the tail, -P1 for 1 - tail
    const double t2 = flip_sign_unless(P2, x2);
    const double bs = zero_if_negative(s, (int)x1);         // s when N1 = 1 - w1, else 0
    const double bf = zero_if_negative(fv, (int)x2);        // fv when N2 = 1 - w2, else 0
    const double c = fma(g, t1 - t2, bs - bf);
    // Domain of the blocks above (integer range checks on the high words; everything else -- t = 0, v = 0, s <= 0,
    // NaN, inf, denormals, overflow -- is left to the IEEE-order path, which reproduces the reference's inf/NaN results):
    //   v, v^2 t and s/k positive normal numbers in [2^-255, 2^256) (v^2 stays normal; t > 0 follows; a k the reciprocal
    //              block cannot invert -- zero, denormal, >= 2^1022, inf, NaN -- yields s/k = 0, inf or NaN);
    //   |d1| < 37: exp(-d1^2/2) has not underflowed, so expressing the second tail through it loses nothing
    //              (inside the inputgen range |d1| <= 32);
    //   a1 a2 = (1 + c|d1|)(1 + c|d2|) finite: the shared reciprocal is not 0 * inf -- implied by the two above;
    //   |r t| < 512: the exponent arithmetic of exp_core_f64 cannot wrap (the reference returns inf / NaN there).
    // (a1 a2 needs no check of its own: |d1| < 37 and xDen = sqrt(v^2 t) < 2^128 bound it by 2^130.)
    // The three range checks share one compare: x - LO < SPAN for all three <=> the OR of the differences is < SPAN
    // (SPAN is a power of two).  |d1| < 37 is tested on the exponential's argument q = -d1^2/2, whose sign bit is always
    // set: high word - 0x80000000 < high word of 684.5 (a NaN fails with either sign); |r t| < 512 on the high word
    // shifted left by one (drops the sign).
    const uint32_t LO = 0x30000000u, SPAN = 0x20000000u;
    const uint32_t span3 = ((uint32_t)(to_bits(v) >> 32) - LO) | ((uint32_t)(to_bits(z) >> 32) - LO) | ((uint32_t)(to_bits(sk) >> 32) - LO);
    *ok = (span3 < SPAN) && ((uint32_t)(to_bits(q1) >> 32) - 0x80000000u < 0x40856400u) && ((uint32_t)(to_bits(rt) >> 32) << 1 < 0x81000000u);
    return from_bits(to_bits(c) ^ ((uint64_t)puth << 32));   // put: fv N(-d2) - s N(-d1) = -(s N(-d1) - fv N(-d2))
}

}  // namespace bsm
