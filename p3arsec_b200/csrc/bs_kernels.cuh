/*
 * bs_kernels.cuh -- sm_100a kernels of the blackscholes Map.
 *
 * What is computed is the reference's per-option body
 *     prices[i] = BlkSchlsEqEuroNoDiv(sptprice[i], strike[i], rate[i], volatility[i], otime[i], otype[i], 0)
 * (/root/reference/parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c:328-331, kernel :190-258,
 * CNDF :128-184), plus the optional ERR_CHK compare (:333-340).  How it is computed is B200-first:
 *
 *   - HBM-bound streaming Map, 28 B/option (fp32) or 52 B/option (fp64): six read streams and one
 *     write stream, every access a coalesced 128-bit vector transaction (LDG.E.128 / STG.E.128),
 *     read-only non-coherent path, no L1 allocation (nothing is reused inside a run).
 *   - persistent grid: blocks = SMs x resident CTAs/SM, interleaved grid-stride over 16-byte groups
 *     so that every SM owns the same number of groups (+-1) and no tail wave exists; UNROLL
 *     independent groups per thread per trip keep >= 96*UNROLL bytes in flight per thread.
 *   - no tensor cores (nothing is a contraction); shared memory only for the fp64 lookup tables and, in the
 *     alternative bs_map_tma kernel, for the bulk-copy ring.
 *   - the put/call select is branch-free; CNDF's sign handling uses N(-x) = 1 - N(x).
 *
 * Kernels: bs_map<FP, MATH, UNROLL, CHK, PIPE> (the default, LDG.128 streams; PIPE = software-pipelined loads),
 * bs_map_tma<FP, MATH> (inputs by cp.async.bulk into a shared-memory ring; measured slower, kept as an option),
 * bs_fill_synthetic<FP>.  Math flavours (bs_gpu_math in include/bs_gpu.h), both precisions:
 *   BS_MATH_IEEE  libdevice exp/log/sqrt and IEEE-rounded divisions in the reference's operation order (fp32: pure
 *                 fp32, the reference's double-literal promotions are not imitated; fp64: explicitly rounded, never
 *                 contracted operations, so only the last ulp of exp()/log() can differ from the fp64 CPU build).
 *   BS_MATH_FAST  fp32: 9 MUFU ops per option (sqrt, rcp x3, lg2 x2, ex2 x3) with log2(e)/ln(2) and 1/sqrt(2 pi)
 *                 folded into constants, Horner form of the degree-5 polynomial, one Newton step on each
 *                 reciprocal.  fp64: bs_math_f64.h.
 *   MATH_PROBE    diagnostic only: no pricing, same streams (the bandwidth ceiling of the traffic pattern).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bs_libm_f32.h"
#include "bs_libm_f64.h"
#include "bs_math_f64.h"

namespace bsk {

enum { MATH_PROBE = 0, MATH_IEEE = 1, MATH_FAST = 2, MATH_REFERENCE = 3 };  // MATH_PROBE: no pricing, traffic only (diagnostic)

// ---------------------------------------------------------------------------------------------
// 128-bit streaming accessors
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ld_stream(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream(const int2 *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
// The same loads under a predicate (the destination keeps its value when `on` is false): the software-pipelined loop issues
// the next trip's loads without a branch, so they stay where the program puts them -- ahead of the current trip's math.
__device__ __forceinline__ void ld_stream_if(float4 &r, const float4 *p, int on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w) : "l"(p), "r"(on));
}
__device__ __forceinline__ void ld_stream_if(int4 &r, const int4 *p, int on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"(on));
}
__device__ __forceinline__ void ld_stream_if(double2 &r, const double2 *p, int on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n\t}"
                 : "+d"(r.x), "+d"(r.y) : "l"(p), "r"(on));
}
__device__ __forceinline__ void ld_stream_if(int2 &r, const int2 *p, int on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];\n\t}"
                 : "+r"(r.x), "+r"(r.y) : "l"(p), "r"(on));
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(double2 *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------------------
// MUFU wrappers (one SASS MUFU.* each)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 1/x to <= 1 ulp: MUFU.RCP plus one Newton step.  The raw approximation's last-bit errors are amplified by the
// degree-5 CNDF polynomial and by d1 = num / den: over 20M random options the step brings the largest distance to the
// reference output from 8.3e-5 to 7.3e-5 for 0.3 us per 10M-option pass (profiles/r01_fp32_accuracy_knobs.txt).
// For x = 0 or inf the correction term is NaN (0 * inf) and the raw result (inf or 0) is kept, so the degenerate
// inputs (t = 0, v = 0) still follow the reference's inf arithmetic.
__device__ __forceinline__ float rcp_refined(float x)
{
    const float r = mufu_rcp(x);
    const float c = fmaf(r, fmaf(-x, r, 1.0f), r);
    return (c == c) ? c : r;
}

// ---------------------------------------------------------------------------------------------
// fp32, BS_MATH_FAST
// ---------------------------------------------------------------------------------------------
// 1 - N(|d|) = n(d) * poly(k),  k = 1/(1 + 0.2316419|d|)      (CNDF, blackscholes.c:152-174)
// The five A&S coefficients are pre-multiplied by 1/sqrt(2 pi) (:126,:154); exp(-d*d/2) is one ex2.
__device__ __forceinline__ float cndf_tail_fast(float d)
{
    const float A1 = 0.319381530f * 0.39894228040143270286f;
    const float A2 = -0.356563782f * 0.39894228040143270286f;
    const float A3 = 1.781477937f * 0.39894228040143270286f;
    const float A4 = -1.821255978f * 0.39894228040143270286f;
    const float A5 = 1.330274429f * 0.39894228040143270286f;
    const float NEG_HALF_LOG2E = -0.72134752044448170368f;

    float k = rcp_refined(fmaf(fabsf(d), 0.2316419f, 1.0f));
    float e = mufu_ex2((d * NEG_HALF_LOG2E) * d);
    float p = fmaf(k, A5, A4);
    p = fmaf(k, p, A3);
    p = fmaf(k, p, A2);
    p = fmaf(k, p, A1);
    return (p * k) * e;
}

__device__ __forceinline__ float price_fast(float s, float k, float r, float v, float t, int otype)
{
    const float LN2 = 0.69314718055994530942f;
    const float NEG_LOG2E = -1.44269504088896340736f;

    float sq = mufu_sqrt(t);                         // xSqrtTime            :224
    float den = v * sq;                              // xDen                 :238
    float rden = rcp_refined(den);
    float lg = mufu_lg2(s) - mufu_lg2(k);            // log2(s/k)            :226
    float drift = fmaf(0.5f * v, v, r);              // r + v*v/2            :231-234
    float num = fmaf(lg, LN2, drift * t);            // (..)*t + ln(s/k)     :235-236
    float d1 = num * rden;                           //                      :239
    float d2 = d1 - den;                             //                      :240
    float fv = k * mufu_ex2((r * NEG_LOG2E) * t);    // strike*exp(-r t)     :248

    float w1 = cndf_tail_fast(d1);
    float w2 = cndf_tail_fast(d2);
    // call: s N(d1) - fv N(d2);  put: fv N(-d2) - s N(-d1) = -(s N(-d1) - fv N(-d2))   :249-255
    // N(x) = w for x<0, 1-w otherwise; N(-x) the other way round.
    bool put = (otype != 0);
    float x1 = ((d1 < 0.0f) != put) ? w1 : 1.0f - w1;
    float x2 = ((d2 < 0.0f) != put) ? w2 : 1.0f - w2;
    float c = fmaf(s, x1, -(fv * x2));
    return put ? -c : c;
}

// ---------------------------------------------------------------------------------------------
// fp32, BS_MATH_IEEE: the reference's operation order with libdevice expf/logf/sqrtf and
// IEEE-rounded divisions (nvcc default: -prec-div=true -prec-sqrt=true, no fast-math).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float cndf_ieee(float x)
{
    bool neg = x < 0.0f;
    float ax = fabsf(x);
    float npx = expf(-0.5f * ax * ax) * 0.39894228040143270286f;
    float k1 = 1.0f / (1.0f + 0.2316419f * ax);
    float k2 = k1 * k1, k3 = k2 * k1, k4 = k3 * k1, k5 = k4 * k1;
    float lead = k1 * 0.319381530f;
    float acc = k2 * (-0.356563782f);
    acc = acc + k3 * 1.781477937f;
    acc = acc + k4 * (-1.821255978f);
    acc = acc + k5 * 1.330274429f;
    float out = 1.0f - (acc + lead) * npx;
    return neg ? 1.0f - out : out;
}

__device__ __forceinline__ float price_ieee(float s, float k, float r, float v, float t, int otype)
{
    float sq = sqrtf(t);
    float lg = logf(s / k);
    float d1 = (r + v * v * 0.5f) * t + lg;
    float den = v * sq;
    d1 = d1 / den;
    float d2 = d1 - den;
    float n1 = cndf_ieee(d1);
    float n2 = cndf_ieee(d2);
    float fv = k * expf(-r * t);
    float call = s * n1 - fv * n2;
    float put = fv * (1.0f - n2) - s * (1.0f - n1);
    return otype == 0 ? call : put;
}

// ---------------------------------------------------------------------------------------------
// fp32, BS_MATH_REFERENCE: the reference's fp32 build AS COMPILED -- its constants are double literals, so every
// expression that touches one is evaluated in double and rounded back to float on assignment (blackscholes.c:154,
// 156-158,164-170,175,180,232,252-253; verified by disassembly), while literal-free expressions stay in
// float.  Every operation below is individually rounded (no FMA contraction, like the x86-64 build), and expf/logf
// are glibc 2.39's own algorithms restated in bs_libm_f32.h (bit-identical to the host libm over all 2^32 floats).
// The result is the reference CPU output BIT FOR BIT; this is the validation mode: it separates "rounding of the
// reference" from "kernel defect" for the fast modes, at several times their cost.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ref_expf(float x) { return bsl::expf_glibc(x); }
__device__ __forceinline__ float ref_logf(float x) { return bsl::logf_glibc(x); }
__device__ __forceinline__ float ref_mul_lit(float a, double lit) { return __double2float_rn(__dmul_rn((double)a, lit)); }

__device__ __noinline__ float cndf_reference(float x)
{
    const bool neg = x < 0.0f;                                             // :144-148
    if (neg) x = -x;
    const float ax = x;
    float npx = ref_expf(__fmul_rn(__fmul_rn(-0.5f, x), x));               // :152  float literal: stays in fptype
    npx = ref_mul_lit(npx, 0.39894228040143270286);                        // :154
    float k1 = ref_mul_lit(ax, 0.2316419);                                 // :156
    k1 = __double2float_rn(__dadd_rn(1.0, (double)k1));                    // :157
    k1 = __double2float_rn(__ddiv_rn(1.0, (double)k1));                    // :158
    const float k2 = __fmul_rn(k1, k1);                                    // :159-162
    const float k3 = __fmul_rn(k2, k1);
    const float k4 = __fmul_rn(k3, k1);
    const float k5 = __fmul_rn(k4, k1);
    float lead = ref_mul_lit(k1, 0.319381530);                             // :164
    float acc = ref_mul_lit(k2, -0.356563782);                             // :165
    acc = __fadd_rn(acc, ref_mul_lit(k3, 1.781477937));                    // :166-167
    acc = __fadd_rn(acc, ref_mul_lit(k4, -1.821255978));                   // :168-169
    acc = __fadd_rn(acc, ref_mul_lit(k5, 1.330274429));                    // :170-171
    lead = __fadd_rn(acc, lead);                                           // :173
    float out = __fmul_rn(lead, npx);                                      // :174
    out = __double2float_rn(__dsub_rn(1.0, (double)out));                  // :175
    if (neg) out = __double2float_rn(__dsub_rn(1.0, (double)out));         // :179-181
    return out;
}

__device__ __forceinline__ float price_reference(float s, float k, float r, float v, float t, int otype)
{
    const float sq = __fsqrt_rn(t);                                        // :224
    const float lg = ref_logf(__fdiv_rn(s, k));                            // :226
    float pw = __fmul_rn(v, v);                                            // :231
    pw = ref_mul_lit(pw, 0.5);                                             // :232
    float d1 = __fadd_rn(r, pw);                                           // :234
    d1 = __fmul_rn(d1, t);                                                 // :235
    d1 = __fadd_rn(d1, lg);                                                // :236
    const float den = __fmul_rn(v, sq);                                    // :238
    d1 = __fdiv_rn(d1, den);                                               // :239
    const float d2 = __fsub_rn(d1, den);                                   // :240
    const float n1 = cndf_reference(d1);                                   // :245
    const float n2 = cndf_reference(d2);                                   // :246
    const float fv = __fmul_rn(k, ref_expf(__fmul_rn(-r, t)));             // :248
    if (otype == 0) return __fsub_rn(__fmul_rn(s, n1), __fmul_rn(fv, n2)); // :250
    const float m1 = __double2float_rn(__dsub_rn(1.0, (double)n1));        // :252
    const float m2 = __double2float_rn(__dsub_rn(1.0, (double)n2));        // :253
    return __fsub_rn(__fmul_rn(fv, m2), __fmul_rn(s, m1));                 // :254
}

template <int MATH>
__device__ __forceinline__ float price_f32(float s, float k, float r, float v, float t, int otype)
{
    if (MATH == MATH_REFERENCE) return price_reference(s, k, r, v, t, otype);
    // MATH_PROBE is a measurement aid, not a pricing mode: same seven streams, same access pattern, five adds.
    // Its bandwidth is the ceiling of THIS traffic pattern (6 read streams : 1 write stream) on the device.
    if (MATH == MATH_PROBE) return s + k + r + v + t + (float)otype;
    if (MATH == MATH_FAST) return price_fast(s, k, r, v, t, otype);
    return price_ieee(s, k, r, v, t, otype);
}

// ---------------------------------------------------------------------------------------------
// fp64: reference operation order, every operation individually rounded (no FMA contraction).
// GLIBC = false (BS_MATH_IEEE): libdevice exp/log -- the last ulp of those two can differ from the CPU build's libm.
// GLIBC = true (BS_MATH_REFERENCE): glibc 2.39's own exp/log restated bit-exactly (bs_libm_f64.h) -- the prices are the
// reference CPU prices bit for bit (sqrt and the divisions are correctly rounded on both sides).
// ---------------------------------------------------------------------------------------------
template <bool GLIBC> __device__ __forceinline__ double exp_ref64(double x) { return GLIBC ? bsl64::exp_glibc(x) : exp(x); }
template <bool GLIBC> __device__ __forceinline__ double log_ref64(double x) { return GLIBC ? bsl64::log_glibc(x) : log(x); }

template <bool GLIBC>
__device__ __forceinline__ double cndf_f64(double x)
{
    bool neg = x < 0.0;
    double ax = neg ? -x : x;
    double npx = exp_ref64<GLIBC>(__dmul_rn(__dmul_rn(-0.5, ax), ax));
    npx = __dmul_rn(npx, 0.39894228040143270286);
    double k1 = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(0.2316419, ax)));
    double k2 = __dmul_rn(k1, k1);
    double k3 = __dmul_rn(k2, k1);
    double k4 = __dmul_rn(k3, k1);
    double k5 = __dmul_rn(k4, k1);
    double lead = __dmul_rn(k1, 0.319381530);
    double acc = __dmul_rn(k2, -0.356563782);
    acc = __dadd_rn(acc, __dmul_rn(k3, 1.781477937));
    acc = __dadd_rn(acc, __dmul_rn(k4, -1.821255978));
    acc = __dadd_rn(acc, __dmul_rn(k5, 1.330274429));
    double out = __dsub_rn(1.0, __dmul_rn(__dadd_rn(acc, lead), npx));
    return neg ? __dsub_rn(1.0, out) : out;
}

template <bool GLIBC = false>
__device__ __forceinline__ double price_f64(double s, double k, double r, double v, double t, int otype)
{
    double sq = __dsqrt_rn(t);
    double lg = log_ref64<GLIBC>(__ddiv_rn(s, k));
    double pw = __dmul_rn(__dmul_rn(v, v), 0.5);
    double d1 = __dadd_rn(__dmul_rn(__dadd_rn(r, pw), t), lg);
    double den = __dmul_rn(v, sq);
    d1 = __ddiv_rn(d1, den);
    double d2 = __dsub_rn(d1, den);
    double n1 = cndf_f64<GLIBC>(d1);
    double n2 = cndf_f64<GLIBC>(d2);
    double fv = __dmul_rn(k, exp_ref64<GLIBC>(__dmul_rn(-r, t)));
    double call = __dsub_rn(__dmul_rn(s, n1), __dmul_rn(fv, n2));
    double put = __dsub_rn(__dmul_rn(fv, __dsub_rn(1.0, n2)), __dmul_rn(s, __dsub_rn(1.0, n1)));
    return otype == 0 ? call : put;
}

// Out-of-line copy of the IEEE-order path for the (never hot) degenerate inputs of the fast kernel: keeps the hot
// loop compact and its register allocation free of libdevice's exp/log/div slow paths.
__device__ __noinline__ double price_f64_cold(double s, double k, double r, double v, double t, int otype)
{
    return price_f64<false>(s, k, r, v, t, otype);
}

// fp64 dispatch.  MATH_FAST: bs_math_f64.h (about half the instructions; <= ~2 ulp per building block, measured
// 4e-13 worst absolute distance to the fp64 CPU build on the goldens); inputs it cannot handle (v sqrt(t) not a
// positive normal number) take the IEEE-order path, so degenerate options behave exactly as in the reference.
template <int MATH>
__device__ __forceinline__ double price_f64_any(double s, double k, double r, double v, double t, int otype, const double *tab)
{
    if (MATH == MATH_PROBE) return s + k + r + v + t + (double)otype;
    if (MATH == MATH_FAST) {
        bool ok;
        const double p = bsm::price_f64_fast(s, k, r, v, t, otype, &ok, tab);
        if (__builtin_expect(ok, 1)) return p;
        return price_f64_cold(s, k, r, v, t, otype);
    }
    // MATH_IEEE and MATH_REFERENCE: with fptype=double nothing is promoted, the two differ in whose exp/log they call
    return price_f64<MATH == MATH_REFERENCE>(s, k, r, v, t, otype);
}

// ---------------------------------------------------------------------------------------------
// ERR_CHK (blackscholes.c:333-340): |DGrefval - price| >= 1e-4, the threshold being a double literal
// ---------------------------------------------------------------------------------------------
struct ErrChk {
    unsigned long long *count;  // running total over all runs (the reference's numError)
    unsigned int *list_count;   // offenders recorded by the recording run
    long long *list;            // their shard-local indices
    unsigned int list_cap;
    int record;                 // 1 on the run whose offenders are listed
    long long base;             // shard-local index of element 0 of this launch (chunked launches)
};

__device__ __forceinline__ bool err_bad(float price, float ref) { return (double)fabsf(ref - price) >= 1e-4; }
__device__ __forceinline__ bool err_bad(double price, double ref) { return fabs(__dsub_rn(ref, price)) >= 1e-4; }

__device__ __forceinline__ void err_note(const ErrChk &ec, size_t idx)
{
    if (ec.record) {
        unsigned int slot = atomicAdd(ec.list_count, 1u);
        if (slot < ec.list_cap) ec.list[slot] = ec.base + (long long)idx;
    }
}

__device__ __forceinline__ void err_flush(const ErrChk &ec, unsigned int local_bad)
{
    // warp-level sum, one atomic per warp that saw an error
    for (int off = 16; off > 0; off >>= 1) local_bad += __shfl_xor_sync(0xffffffffu, local_bad, off);
    if ((threadIdx.x & 31) == 0 && local_bad) atomicAdd(ec.count, (unsigned long long)local_bad);
}

// ---------------------------------------------------------------------------------------------
// The Map kernel, one template for both precisions.
//   group  = one 128-bit access per fp stream: 4 options (fp32, int4 of otype) or 2 options (fp64, int2)
//   trip   = UNROLL independent groups per thread, interleaved grid-stride (group g + u*stride)
//   PIPE   = software pipelining: the loads of trip i+1 are issued before the math of trip i, so the DRAM
//            latency hides behind a whole trip of arithmetic instead of behind other warps only.  It matters for
//            the instruction-heavy fp64 kernel (long_scoreboard was its top stall at 38 % active warps).
// ---------------------------------------------------------------------------------------------
template <typename FP> struct VT;
template <> struct VT<float> { typedef float4 vec; typedef int4 ivec; enum { LANES = 4, SHIFT = 2 }; };
template <> struct VT<double> { typedef double2 vec; typedef int2 ivec; enum { LANES = 2, SHIFT = 1 }; };

__device__ __forceinline__ float lane(const float4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ int lane(const int4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ double lane(const double2 &v, int i) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ int lane(const int2 &v, int i) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ void set_lane(float4 &v, int i, float x) { if (i == 0) v.x = x; else if (i == 1) v.y = x; else if (i == 2) v.z = x; else v.w = x; }
__device__ __forceinline__ void set_lane(double2 &v, int i, double x) { if (i == 0) v.x = x; else v.y = x; }

template <typename FP>
struct Streams {
    const FP *spt, *strike, *rate, *vol, *otime;
    const int *otype;
    FP *prices;
    const FP *refval;  // DGrefval stream, only read when CHK
};
typedef Streams<float> StreamsF32;
typedef Streams<double> StreamsF64;

template <typename FP>
struct Group {
    typename VT<FP>::vec s, k, r, v, t, ref;
    typename VT<FP>::ivec o;
};

template <int MATH> __device__ __forceinline__ float price_any(float s, float k, float r, float v, float t, int o, const double *) { return price_f32<MATH>(s, k, r, v, t, o); }
template <int MATH> __device__ __forceinline__ double price_any(double s, double k, double r, double v, double t, int o, const double *tab) { return price_f64_any<MATH>(s, k, r, v, t, o, tab); }

// One group (4 fp32 / 2 fp64 options) priced lane by lane.  BS_F64_GROUP_ILP = 1 makes the fp64 fast path defer its
// degenerate-input test until every lane has been evaluated: one branch per group instead of one per option, so that the
// lanes' dependent FP64 chains share a basic block and the compiler interleaves them.  Measured on B200 (profiles/
// r02_tune_fp64_ilp.txt): 110.9-111.1 G options/s sustained against 112.7-113.0 with one branch per option (six more
// registers and a longer loop; the warps of the other schedulers already fill the `wait` slots), so it is off.
#ifndef BS_F64_GROUP_ILP
#define BS_F64_GROUP_ILP 0
#endif
template <int MATH, typename FP>
__device__ __forceinline__ typename VT<FP>::vec price_group(const typename VT<FP>::vec &s, const typename VT<FP>::vec &k, const typename VT<FP>::vec &r,
                                                            const typename VT<FP>::vec &v, const typename VT<FP>::vec &t,
                                                            const typename VT<FP>::ivec &o, const double *tab)
{
    typename VT<FP>::vec p;
    if constexpr (sizeof(FP) == 8 && MATH == MATH_FAST && BS_F64_GROUP_ILP) {
        bool ok[VT<FP>::LANES];
        bool all = true;
#pragma unroll
        for (int l = 0; l < VT<FP>::LANES; l++) {
            set_lane(p, l, (FP)bsm::price_f64_fast((double)lane(s, l), (double)lane(k, l), (double)lane(r, l), (double)lane(v, l), (double)lane(t, l),
                                                   lane(o, l), &ok[l], tab));
            all = all && ok[l];
        }
        if (__builtin_expect(!all, 0)) {
#pragma unroll
            for (int l = 0; l < VT<FP>::LANES; l++)
                if (!ok[l])
                    set_lane(p, l, (FP)price_f64_cold((double)lane(s, l), (double)lane(k, l), (double)lane(r, l), (double)lane(v, l), (double)lane(t, l), lane(o, l)));
        }
    } else {
#pragma unroll
        for (int l = 0; l < VT<FP>::LANES; l++) set_lane(p, l, price_any<MATH>(lane(s, l), lane(k, l), lane(r, l), lane(v, l), lane(t, l), lane(o, l), tab));
    }
    return p;
}


template <typename FP, int MATH, int UNROLL, bool CHK, bool PIPE>
__global__ void __launch_bounds__(256) bs_map(Streams<FP> a, size_t n, ErrChk ec)
{
    typedef typename VT<FP>::vec vec;
    typedef typename VT<FP>::ivec ivec;
    enum { LANES = VT<FP>::LANES, SHIFT = VT<FP>::SHIFT };

    // the fp64 fast math reads two 64-entry tables (bs_math_f64.h) from shared memory; other variants carry 8 bytes
    enum { USE_TAB = (sizeof(FP) == 8 && MATH == MATH_FAST) ? 1 : 0 };
    __shared__ __align__(16) unsigned char s_tab_raw[USE_TAB ? bsm::BS_F64_TAB_DOUBLES * sizeof(double) + bsm::BS_F64_TAB_PAD : 16];
    double *s_tab = USE_TAB ? bsm::place_tables(s_tab_raw) : reinterpret_cast<double *>(s_tab_raw);
    if (USE_TAB) {
        bsm::BS_F64_FILL_TABLES(s_tab, (int)threadIdx.x, (int)blockDim.x);
        __syncthreads();
    }
    const double *tab = s_tab;

    const size_t groups = n >> SHIFT;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;

    const vec *p_s = reinterpret_cast<const vec *>(a.spt);
    const vec *p_k = reinterpret_cast<const vec *>(a.strike);
    const vec *p_r = reinterpret_cast<const vec *>(a.rate);
    const vec *p_v = reinterpret_cast<const vec *>(a.vol);
    const vec *p_t = reinterpret_cast<const vec *>(a.otime);
    const ivec *p_o = reinterpret_cast<const ivec *>(a.otype);
    const vec *p_ref = reinterpret_cast<const vec *>(a.refval);
    vec *p_out = reinterpret_cast<vec *>(a.prices);

    // all loads of a trip are issued back to back, before any of its math (group 0 of a trip is always in range)
    auto load_trip = [&](Group<FP> *dst, size_t g0) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g0 + u * stride;
            if (u == 0 || gi < groups) {
                dst[u].s = ld_stream(p_s + gi);
                dst[u].k = ld_stream(p_k + gi);
                dst[u].r = ld_stream(p_r + gi);
                dst[u].v = ld_stream(p_v + gi);
                dst[u].t = ld_stream(p_t + gi);
                dst[u].o = ld_stream(p_o + gi);
                if (CHK) dst[u].ref = ld_stream(p_ref + gi);
            }
        }
    };
    auto price_trip = [&](const Group<FP> *src, size_t g0) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g0 + u * stride;
            if (u == 0 || gi < groups) {
                const vec p = price_group<MATH, FP>(src[u].s, src[u].k, src[u].r, src[u].v, src[u].t, src[u].o, tab);
                st_stream(p_out + gi, p);
                if (CHK) {
#pragma unroll
                    for (int l = 0; l < LANES; l++)
                        if (err_bad(lane(p, l), lane(src[u].ref, l))) { bad++; err_note(ec, gi * LANES + l); }
                }
            }
        }
    };

    // Programmatic dependent launch (opt-in, BS_GPU_FLAG_PDL: run j+1 of the NUM_RUNS sequence carries the attribute): let the
    // next run's CTAs be scheduled as soon as all of ours are running, and order OUR first store after the
    // previous run's completion.  Our first trip of loads is issued before that wait, so the drain of run j and
    // the fill of run j+1 overlap instead of leaving the memory system idle between launches.  Every run still
    // reads every input and writes every price, in run order.  Both instructions are no-ops in a plain launch.
    asm volatile("griddepcontrol.launch_dependents;");
    if (!PIPE) {
        if (g < groups) {
            Group<FP> cur[UNROLL];
            load_trip(cur, g);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            price_trip(cur, g);
            g += (size_t)UNROLL * stride;
        } else {
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
        for (; g < groups; g += (size_t)UNROLL * stride) {
            Group<FP> cur[UNROLL];
            load_trip(cur, g);
            price_trip(cur, g);
        }
    } else if (g >= groups) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    } else {
        // two register sets used in turn (ping-pong), so no trip is ever copied from "next" to "current"
        Group<FP> bufA[UNROLL], bufB[UNROLL];
        const size_t step = (size_t)UNROLL * stride;
        load_trip(bufA, g);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // The next trip's loads are predicated, not branched around: with a conditional block the compiler scheduled the
        // current trip's math ahead of it once the fp64 math had shrunk, and the loads lost 150 instructions of head start.
        auto load_trip_if = [&](Group<FP> *dst, size_t g0, bool on) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const size_t gi = g0 + u * stride;
                const int p = (on && (u == 0 || gi < groups)) ? 1 : 0;
                const size_t gs = p ? gi : 0;  // a predicated-off load still forms its address
                ld_stream_if(dst[u].s, p_s + gs, p);
                ld_stream_if(dst[u].k, p_k + gs, p);
                ld_stream_if(dst[u].r, p_r + gs, p);
                ld_stream_if(dst[u].v, p_v + gs, p);
                ld_stream_if(dst[u].t, p_t + gs, p);
                ld_stream_if(dst[u].o, p_o + gs, p);
                if (CHK) ld_stream_if(dst[u].ref, p_ref + gs, p);
            }
        };
        for (;;) {
            bool more = g + step < groups;
            load_trip_if(bufB, g + step, more);  // in flight while trip A is priced
            price_trip(bufA, g);
            if (!more) break;
            g += step;
            more = g + step < groups;
            load_trip_if(bufA, g + step, more);  // in flight while trip B is priced
            price_trip(bufB, g);
            if (!more) break;
            g += step;
        }
    }
    // ragged tail: the last n % LANES options, one scalar option per thread of block 0
    const size_t tail0 = groups << SHIFT;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < n) {
        const size_t i = tail0 + threadIdx.x;
        FP p = price_any<MATH>(a.spt[i], a.strike[i], a.rate[i], a.vol[i], a.otime[i], a.otype[i], tab);
        a.prices[i] = p;
        if (CHK && err_bad(p, a.refval[i])) { bad++; err_note(ec, i); }
    }
    if (CHK) err_flush(ec, bad);
}

// ---------------------------------------------------------------------------------------------
// TMA variant (bs_gpu_config.variant bit 2): the same Map with the input streams moved by the bulk-copy engine.
//   * one persistent CTA per SM slot; warp 8 is the PRODUCER: one elected lane issues, per tile, six
//     cp.async.bulk global->shared copies (one per input stream, TILE options each) that complete on the stage's
//     "full" mbarrier (expect_tx = bytes of the tile);
//   * warps 0-7 are CONSUMERS: wait on "full", LDS.128 their group out of every stream tile, release the stage
//     ("empty" mbarrier, one arrival per warp) BEFORE doing the math, price, STG.128 the prices;
//   * STAGES tiles are in flight per CTA (STAGES x 24 KB fp32), so no registers are tied up by loads in flight and
//     the LSU issues six LDS instead of six LDG + their 64-bit address arithmetic per group.
// Whole tiles only; the last n % TILE options are priced by the consumers of block 0 with plain loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// SHAPE 0: 16 consumer warps, 4 stages.  SHAPE 1 (fp64 only; the fp32 entry is an alias of shape 0): 24 consumer warps, 3 stages --
// the fp64 math is issue-bound with long dependent DFMA chains, so more resident warps per scheduler hide more of its
// fixed-latency stalls; registers are then capped at 80 per thread by the launch bounds.
#ifndef BS_TMA_ARRIVE_ALL
#define BS_TMA_ARRIVE_ALL 1  /* 1 (default): every consumer thread arrives on the "empty" barrier itself; 0: one elected lane per warp */
#endif
template <typename FP, int SHAPE> struct TmaCfg;
template <int SHAPE> struct TmaCfg<float, SHAPE> { enum { TILE = 2048, STAGES = 4, CONSUMERS = 512 }; };  // 6 x 8 KB per stage; 192 KB per CTA, 1 CTA per SM
template <> struct TmaCfg<double, 0> { enum { TILE = 1024, STAGES = 4, CONSUMERS = 512 }; };              // 5 x 8 KB + 4 KB per stage; 176 KB per CTA
template <> struct TmaCfg<double, 1> { enum { TILE = 1536, STAGES = 3, CONSUMERS = 768 }; };              // 66 KB per stage; 198 KB per CTA
// SHAPE 2 (fp64): two groups (four options) per consumer thread and tile -- twice the independent work per warp for the
// dependent DFMA chains, and the per-tile overhead (barrier wait, addresses, loop) paid once per four options; the same
// 176 KB in flight as two stages of 88 KB.
template <> struct TmaCfg<double, 2> { enum { TILE = 2048, STAGES = 2, CONSUMERS = 512 }; };

template <typename FP, int SHAPE> __host__ __device__ constexpr size_t tma_stage_bytes() { return (size_t)TmaCfg<FP, SHAPE>::TILE * (5 * sizeof(FP) + sizeof(int)); }
template <typename FP, int SHAPE> __host__ __device__ constexpr size_t tma_smem_bytes()
{
    return tma_stage_bytes<FP, SHAPE>() * TmaCfg<FP, SHAPE>::STAGES + 2 * TmaCfg<FP, SHAPE>::STAGES * sizeof(uint64_t) + 128 +
           ((sizeof(FP) == 8) ? bsm::BS_F64_TAB_DOUBLES * sizeof(double) + bsm::BS_F64_TAB_PAD : 0);
}
template <typename FP, int SHAPE> __host__ __device__ constexpr int tma_threads() { return TmaCfg<FP, SHAPE>::CONSUMERS + 32; }

template <typename FP, int MATH, int SHAPE>
__global__ void __launch_bounds__(TmaCfg<FP, SHAPE>::CONSUMERS + 32, 1) bs_map_tma(Streams<FP> a, size_t n, ErrChk ec)
{
    typedef typename VT<FP>::vec vec;
    typedef typename VT<FP>::ivec ivec;
    enum { LANES = VT<FP>::LANES, TILE = TmaCfg<FP, SHAPE>::TILE, STAGES = TmaCfg<FP, SHAPE>::STAGES, GROUPS = TILE / LANES,
           TMA_CONSUMERS = TmaCfg<FP, SHAPE>::CONSUMERS };
    constexpr unsigned FP_TILE_BYTES = TILE * sizeof(FP), OT_TILE_BYTES = TILE * sizeof(int);
    constexpr unsigned STAGE_BYTES = 5 * FP_TILE_BYTES + OT_TILE_BYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage_base = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGE_BYTES * STAGES);
    uint64_t *empty = full + STAGES;
    enum { USE_TAB = (sizeof(FP) == 8 && MATH == MATH_FAST) ? 1 : 0 };
    double *s_tab = USE_TAB ? bsm::place_tables(empty + STAGES + 2) : reinterpret_cast<double *>(empty + STAGES + 2);
    (void)ec;

    const size_t tiles = n / TILE;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int st = 0; st < STAGES; st++) {
            mbar_init(&full[st], 1);                      // the producer's expect_tx arrival
            mbar_init(&empty[st], BS_TMA_ARRIVE_ALL ? TMA_CONSUMERS : TMA_CONSUMERS / 32);  // one arrival per consumer warp (or thread)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (USE_TAB) bsm::BS_F64_FILL_TABLES(s_tab, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();

    if (warp == TMA_CONSUMERS / 32) {
        // ---- producer: one lane keeps STAGES tiles in flight
        if ((threadIdx.x & 31) == 0) {
            unsigned it = 0;
            for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
                const int st = it % STAGES;
                const unsigned round = it / STAGES;
                if (round > 0) mbar_wait(&empty[st], (round - 1) & 1);   // consumers have drained this stage
                unsigned char *dst = stage_base + (size_t)st * STAGE_BYTES;
                const size_t o0 = t * TILE;
                mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                bulk_g2s(dst + 0 * FP_TILE_BYTES, a.spt + o0, FP_TILE_BYTES, &full[st]);
                bulk_g2s(dst + 1 * FP_TILE_BYTES, a.strike + o0, FP_TILE_BYTES, &full[st]);
                bulk_g2s(dst + 2 * FP_TILE_BYTES, a.rate + o0, FP_TILE_BYTES, &full[st]);
                bulk_g2s(dst + 3 * FP_TILE_BYTES, a.vol + o0, FP_TILE_BYTES, &full[st]);
                bulk_g2s(dst + 4 * FP_TILE_BYTES, a.otime + o0, FP_TILE_BYTES, &full[st]);
                bulk_g2s(dst + 5 * FP_TILE_BYTES, a.otype + o0, OT_TILE_BYTES, &full[st]);
            }
        }
    } else {
        // ---- consumers
        vec *p_out = reinterpret_cast<vec *>(a.prices);
        unsigned it = 0;
        for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
            const int st = it % STAGES;
            mbar_wait(&full[st], (it / STAGES) & 1);
            const unsigned char *src = stage_base + (size_t)st * STAGE_BYTES;
            // GROUPS groups per tile, GPT of them per consumer thread (one, or two for shape 2): thread i owns groups i, i + CONSUMERS
            enum { GPT = GROUPS / TMA_CONSUMERS };
            static_assert(GROUPS == GPT * TMA_CONSUMERS, "whole groups per consumer thread and tile");
            vec vs[GPT], vk[GPT], vr[GPT], vv[GPT], vt[GPT];
            ivec vo[GPT];
#pragma unroll
            for (int u = 0; u < GPT; u++) {
                const int gi = threadIdx.x + u * TMA_CONSUMERS;
                vs[u] = reinterpret_cast<const vec *>(src + 0 * FP_TILE_BYTES)[gi];
                vk[u] = reinterpret_cast<const vec *>(src + 1 * FP_TILE_BYTES)[gi];
                vr[u] = reinterpret_cast<const vec *>(src + 2 * FP_TILE_BYTES)[gi];
                vv[u] = reinterpret_cast<const vec *>(src + 3 * FP_TILE_BYTES)[gi];
                vt[u] = reinterpret_cast<const vec *>(src + 4 * FP_TILE_BYTES)[gi];
                vo[u] = reinterpret_cast<const ivec *>(src + 5 * FP_TILE_BYTES)[gi];
            }
            // The stage is free again (its refill overlaps our math).  Every consumer thread releases its own reads.
            // (One arrival per warp behind a __syncwarp is equally correct -- the lanes' reads are ordered before lane 0's
            // release -- and equally fast, 88-90 us either way, but compute-sanitizer racecheck does not follow that
            // transitive ordering and reports the next round's bulk copy as a hazard against lanes 1-31: 77 reports with
            // the elected-lane form, none with this one; profiles/r02_compute_sanitizer_racecheck.txt.)
            if (BS_TMA_ARRIVE_ALL) {
                mbar_arrive(&empty[st]);
            } else {
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
            }
#pragma unroll
            for (int u = 0; u < GPT; u++) {
                const vec p = price_group<MATH, FP>(vs[u], vk[u], vr[u], vv[u], vt[u], vo[u], s_tab);
                st_stream(p_out + t * GROUPS + threadIdx.x + u * TMA_CONSUMERS, p);
            }
        }
        // the last n % TILE options (no whole tile): plain loads, block 0 only
        if (blockIdx.x == 0) {
            for (size_t i = tiles * TILE + threadIdx.x; i < n; i += TMA_CONSUMERS)
                a.prices[i] = price_any<MATH>(a.spt[i], a.strike[i], a.rate[i], a.vol[i], a.otime[i], a.otype[i], s_tab);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic fill: option i of the shard = table[(first + i) % rows]   (inputgen's cyclic replay)
// ---------------------------------------------------------------------------------------------
template <typename FP>
struct TableDev {
    const FP *spt, *strike, *rate, *vol, *otime, *refval;
    const int *otype;
    int rows;
};

template <typename FP>
__global__ void __launch_bounds__(256) bs_fill_synthetic(TableDev<FP> tab, unsigned long long first, size_t n,
                                                          FP *spt, FP *strike, FP *rate, FP *vol, FP *otime,
                                                          int *otype, FP *refval)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int row = (int)((first + i) % (unsigned long long)tab.rows);
        spt[i] = tab.spt[row];
        strike[i] = tab.strike[row];
        rate[i] = tab.rate[row];
        vol[i] = tab.vol[row];
        otime[i] = tab.otime[row];
        otype[i] = tab.otype[row];
        if (refval) refval[i] = tab.refval[row];
    }
}

// ---------------------------------------------------------------------------------------------
// AoS -> SoA gather for the CAF Map's message layout (blackscholes.c:482-489): records of
//     struct DataCont { int otype; float sptprice, strike, rate, volatility, otime; }   (24 bytes)
// A CTA moves AOS_TILE consecutive records: coalesced 128-bit loads of the tile into shared memory (the tile is
// 6 KB = 384 x 16 B), then thread t reads the six words of record t (stride 6 words: 2-way bank conflict at worst)
// and writes one element of each SoA stream, so every global access of the kernel is coalesced.  `aos` must be
// 16-byte aligned and readable up to a whole tile past the last record (the arena is padded accordingly).
// ---------------------------------------------------------------------------------------------
enum { AOS_TILE = 256, AOS_WORDS = 6 };

__global__ void __launch_bounds__(AOS_TILE) bs_aos_to_soa(const uint4 *aos, size_t n, float *spt, float *strike, float *rate,
                                                          float *vol, float *otime, int *otype)
{
    __shared__ __align__(16) uint32_t tile[AOS_TILE * AOS_WORDS];
    const size_t tiles = (n + AOS_TILE - 1) / AOS_TILE;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint4 *src = aos + t * (AOS_TILE * AOS_WORDS / 4);
        for (int q = threadIdx.x; q < AOS_TILE * AOS_WORDS / 4; q += AOS_TILE) reinterpret_cast<uint4 *>(tile)[q] = src[q];
        __syncthreads();
        const size_t i = t * AOS_TILE + threadIdx.x;
        if (i < n) {
            const uint32_t *rec = tile + threadIdx.x * AOS_WORDS;
            otype[i] = (int)rec[0];
            spt[i] = __uint_as_float(rec[1]);
            strike[i] = __uint_as_float(rec[2]);
            rate[i] = __uint_as_float(rec[3]);
            vol[i] = __uint_as_float(rec[4]);
            otime[i] = __uint_as_float(rec[5]);
        }
        __syncthreads();
    }
}

}  // namespace bsk
