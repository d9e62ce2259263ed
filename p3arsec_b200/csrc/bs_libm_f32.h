/*
 * bs_libm_f32.h -- expf / logf that return what the reference's libm returns, bit for bit.
 *
 * The reference's fp32 build calls glibc's expf (3x per option) and logf (1x) -- blackscholes.c:152,226,248; the
 * only arithmetic of the hot path that lives in a third-party dependency.  BS_MATH_REFERENCE
 * reproduces the reference's own roundings everywhere else, so these two calls decide whether the GPU prices are
 * IDENTICAL to the CPU prices or merely close: a last-bit difference in logf moves a price by up to ~5e-5, because
 * the float evaluation of CNDF is chaotic at that level (measured).
 *
 * Dependency restated: GNU libc 2.39 (the image's libm; sysdeps/ieee754/flt-32/e_expf.c, e_logf.c, the algorithms
 * of ARM's optimized-routines adopted in glibc 2.28): both evaluate a short polynomial in DOUBLE around a table
 *   expf: x N/ln2 = k + r, N = 32;  2^(k/N) from a 32-entry table of bit patterns, degree-3 polynomial in r
 *   logf: x = 2^k z, z in [0x1.66p-1, 0x1.66p0), 16-entry table {1/c, log c}, degree-3 polynomial in r = z/c - 1
 * and round the double once to float.  Multiply-adds are fused as in the FMA build glibc's ifunc selects on every
 * x86-64 CPU with FMA3 (including r = fma(InvLn2N, x, -kd) in expf -- the one contraction that is visible in the
 * float result: 2 of 2.2e9 arguments).  PINNED: tools/libm_f32_host_check.c compares this header, compiled for the
 * host, with the running libm over EVERY float (4.3e9 arguments per function; tests/test_libm_f32.py runs it):
 * 0 mismatches on glibc 2.39.  On the device the same double operations are IEEE-exact (DFMA/DMUL/DADD), so the
 * device results are the host results.
 *
 * Host-or-device code, like bs_math_f64.h.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BSL_HD __device__ __forceinline__
#define BSL_TABLE __device__ const
#else
#define BSL_HD static inline
#define BSL_TABLE static const
#endif

namespace bsl {

// 2^(i/32) as a double bit pattern, minus (i << 52) / 32 so that adding ki << 47 yields 2^(ki/32)   (e_exp2f_data.c)
BSL_TABLE uint64_t EXP2F_TAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

// {1/c, log c} for the 16 intervals of z (e_logf_data.c), as double bit patterns
BSL_TABLE uint64_t LOGF_TAB[16][2] = {
    {0x3ff661ec79f8f3beull, 0xbfd57bf7808caadeull}, {0x3ff571ed4aaf883dull, 0xbfd2bef0a7c06ddbull},
    {0x3ff49539f0f010b0ull, 0xbfd01eae7f513a67ull}, {0x3ff3c995b0b80385ull, 0xbfcb31d8a68224e9ull},
    {0x3ff30d190c8864a5ull, 0xbfc6574f0ac07758ull}, {0x3ff25e227b0b8ea0ull, 0xbfc1aa2bc79c8100ull},
    {0x3ff1bb4a4a1a343full, 0xbfba4e76ce8c0e5eull}, {0x3ff12358f08ae5baull, 0xbfb1973c5a611cccull},
    {0x3ff0953f419900a7ull, 0xbfa252f438e10c1eull}, {0x3ff0000000000000ull, 0x0000000000000000ull},
    {0x3fee608cfd9a47acull, 0x3faaa5aa5df25984ull}, {0x3feca4b31f026aa0ull, 0x3fbc5e53aa362eb4ull},
    {0x3feb2036576afce6ull, 0x3fc526e57720db08ull}, {0x3fe9c2d163a1aa2dull, 0x3fcbc2860d224770ull},
    {0x3fe886e6037841edull, 0x3fd1058bc8a07ee1ull}, {0x3fe767dcf5534862ull, 0x3fd4043057b6ee09ull},
};

BSL_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
BSL_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
BSL_HD uint64_t d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
BSL_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
// individually rounded double operations (no contraction beyond the fma() calls written out below)
BSL_HD double dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host builds of this header use -ffp-contract=off
#endif
}
BSL_HD double dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
BSL_HD float d2f(double y)
{
#if defined(__CUDA_ARCH__)
    return __double2float_rn(y);
#else
    return (float)y;
#endif
}

// glibc 2.39 expf
BSL_HD float expf_glibc(float x)
{
    const uint32_t ix = f2u(x);
    const uint32_t abstop = (ix >> 20) & 0x7ff;
    if (abstop >= 0x42b) {  // |x| >= 88 or NaN: top12(88.0f) = 0x42b
        if (ix == 0xff800000u) return 0.0f;                      // exp(-inf)
        if (abstop >= 0x7f8) return x + x;                       // inf, NaN
        if (x > u2f(0x42b17217u)) return u2f(0x7f800000u);       // > 0x1.62e42ep6f = log(2^128): overflow
        if (x < u2f(0xc2cff1b4u)) return 0.0f;                   // < -0x1.9fe368p6f = log(2^-150): underflow
        if (x < u2f(0xc2ce8ecfu)) return u2f(1u);                // < -0x1.9d1d9ep6f = log(2^-149): 1.5625 2^-150 rounds to 2^-149
    }
    // constants as bit patterns: no decimal conversion between this file and glibc's hex literals
    const double INVLN2N = u2d(0x40471547652b82feull);  // 0x1.71547652b82fep+0 * 32
    const double SHIFT = u2d(0x4338000000000000ull);    // 0x1.8p52
    const double C0 = u2d(0x3ebc6af84b912394ull);       // 0x1.c6af84b912394p-5 / 32^3
    const double C1 = u2d(0x3f2ebfce50fac4f3ull);       // 0x1.ebfce50fac4f3p-3 / 32^2
    const double C2 = u2d(0x3f962e42ff0c52d6ull);       // 0x1.62e42ff0c52d6p-1 / 32
    const double xd = (double)x;
    const double z = dmul(INVLN2N, xd);
    double kd = dadd(z, SHIFT);
    const uint64_t ki = d2u(kd);
    kd = dadd(kd, -SHIFT);
    const double r = fma(INVLN2N, xd, -kd);  // z - kd, contracted by the FMA build
    uint64_t t = EXP2F_TAB[ki & 31];
    t += ki << 47;
    const double s = u2d(t);
    const double zz = fma(C0, r, C1);
    const double r2 = dmul(r, r);
    double y = fma(C2, r, 1.0);
    y = fma(zz, r2, y);
    y = dmul(y, s);
    return d2f(y);
}

// glibc 2.39 logf
BSL_HD float logf_glibc(float x)
{
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return u2f(0xff800000u);                      // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                               // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return u2f(0x7fc00000u);  // negative or NaN: NaN
        ix = f2u(x * 8388608.0f);                                      // subnormal: normalise
        ix -= 23u << 23;
    }
    const double LN2 = u2d(0x3fe62e42fefa39efull);      // 0x1.62e42fefa39efp-1
    const double A0 = u2d(0xbfd00ea348b88334ull);       // -0x1.00ea348b88334p-2
    const double A1 = u2d(0x3fd5575b0be00b6aull);       // 0x1.5575b0be00b6ap-2
    const double A2 = u2d(0xbfdffffef20a4123ull);       // -0x1.ffffef20a4123p-2
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = u2d(LOGF_TAB[i][0]), logc = u2d(LOGF_TAB[i][1]);
    const double z = (double)u2f(iz);
    const double r = fma(z, invc, -1.0);
    const double y0 = fma((double)k, LN2, logc);
    const double r2 = dmul(r, r);
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, dadd(y0, r));
    return d2f(y);
}

}  // namespace bsl
