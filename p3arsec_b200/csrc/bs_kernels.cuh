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
 *   - no shared memory, no tensor cores (nothing is reused, nothing is a contraction).
 *   - the put/call select is branch-free; CNDF's sign handling uses N(-x) = 1 - N(x).
 *
 * Two fp32 math flavours (bs_gpu_math in include/bs_gpu.h):
 *   BS_MATH_IEEE  expf/logf/sqrtf and IEEE-rounded divisions in the reference's operation order
 *                 (pure fp32; the reference's double-literal promotions are not imitated).
 *   BS_MATH_FAST  9 MUFU ops per option (sqrt, rcp x3, lg2 x2, ex2 x3) with log2(e)/ln(2) and
 *                 1/sqrt(2 pi) folded into constants, Horner form of the degree-5 polynomial.
 * fp64 always follows the reference's operation order with explicitly rounded (never contracted)
 * IEEE operations so that the only possible difference to the fp64 CPU build is the last ulp of
 * exp()/log().
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bs_math_f64.h"

namespace bsk {

enum { MATH_IEEE = 1, MATH_FAST = 2 };

// ---------------------------------------------------------------------------------------------
// 128-bit streaming accessors
// ---------------------------------------------------------------------------------------------
template <typename V> struct Vec16;  // 16-byte vector of T

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

    float k = mufu_rcp(fmaf(fabsf(d), 0.2316419f, 1.0f));
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
    float rden = mufu_rcp(den);
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

template <int MATH>
__device__ __forceinline__ float price_f32(float s, float k, float r, float v, float t, int otype)
{
    if (MATH == MATH_FAST) return price_fast(s, k, r, v, t, otype);
    return price_ieee(s, k, r, v, t, otype);
}

// ---------------------------------------------------------------------------------------------
// fp64: reference operation order, every operation individually rounded (no FMA contraction)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cndf_f64(double x)
{
    bool neg = x < 0.0;
    double ax = neg ? -x : x;
    double npx = exp(__dmul_rn(__dmul_rn(-0.5, ax), ax));
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

__device__ __forceinline__ double price_f64(double s, double k, double r, double v, double t, int otype)
{
    double sq = __dsqrt_rn(t);
    double lg = log(__ddiv_rn(s, k));
    double pw = __dmul_rn(__dmul_rn(v, v), 0.5);
    double d1 = __dadd_rn(__dmul_rn(__dadd_rn(r, pw), t), lg);
    double den = __dmul_rn(v, sq);
    d1 = __ddiv_rn(d1, den);
    double d2 = __dsub_rn(d1, den);
    double n1 = cndf_f64(d1);
    double n2 = cndf_f64(d2);
    double fv = __dmul_rn(k, exp(__dmul_rn(-r, t)));
    double call = __dsub_rn(__dmul_rn(s, n1), __dmul_rn(fv, n2));
    double put = __dsub_rn(__dmul_rn(fv, __dsub_rn(1.0, n2)), __dmul_rn(s, __dsub_rn(1.0, n1)));
    return otype == 0 ? call : put;
}

// fp64 dispatch.  MATH_FAST: bs_math_f64.h (about half the instructions; <= ~2 ulp per building block, measured
// 4e-13 worst absolute distance to the fp64 CPU build on the goldens); inputs it cannot handle (v sqrt(t) not a
// positive normal number) take the IEEE-order path, so degenerate options behave exactly as in the reference.
template <int MATH>
__device__ __forceinline__ double price_f64_any(double s, double k, double r, double v, double t, int otype)
{
    if (MATH == MATH_FAST) {
        bool ok;
        const double p = bsm::price_f64_fast(s, k, r, v, t, otype, &ok);
        if (__builtin_expect(ok, 1)) return p;
    }
    return price_f64(s, k, r, v, t, otype);
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
// The fp32 Map kernel.  One "group" = 4 consecutive options = one 128-bit access per stream.
// ---------------------------------------------------------------------------------------------
struct StreamsF32 {
    const float *spt, *strike, *rate, *vol, *otime;
    const int *otype;
    float *prices;
    const float *refval;  // DGrefval stream, only read when CHK
};

template <int MATH, int UNROLL, bool CHK>
__global__ void __launch_bounds__(256) bs_map_f32(StreamsF32 a, size_t n, ErrChk ec)
{
    const size_t groups = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;

    const float4 *p_s = reinterpret_cast<const float4 *>(a.spt);
    const float4 *p_k = reinterpret_cast<const float4 *>(a.strike);
    const float4 *p_r = reinterpret_cast<const float4 *>(a.rate);
    const float4 *p_v = reinterpret_cast<const float4 *>(a.vol);
    const float4 *p_t = reinterpret_cast<const float4 *>(a.otime);
    const int4 *p_o = reinterpret_cast<const int4 *>(a.otype);
    const float4 *p_ref = reinterpret_cast<const float4 *>(a.refval);
    float4 *p_out = reinterpret_cast<float4 *>(a.prices);

    // main trips: UNROLL independent groups per thread, all loads issued before any math
    for (; g + (UNROLL - 1) * stride < groups; g += UNROLL * stride) {
        float4 s[UNROLL], k[UNROLL], r[UNROLL], v[UNROLL], t[UNROLL], ref[UNROLL];
        int4 o[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g + u * stride;
            s[u] = ld_stream(p_s + gi);
            k[u] = ld_stream(p_k + gi);
            r[u] = ld_stream(p_r + gi);
            v[u] = ld_stream(p_v + gi);
            t[u] = ld_stream(p_t + gi);
            o[u] = ld_stream(p_o + gi);
            if (CHK) ref[u] = ld_stream(p_ref + gi);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g + u * stride;
            float4 p;
            p.x = price_f32<MATH>(s[u].x, k[u].x, r[u].x, v[u].x, t[u].x, o[u].x);
            p.y = price_f32<MATH>(s[u].y, k[u].y, r[u].y, v[u].y, t[u].y, o[u].y);
            p.z = price_f32<MATH>(s[u].z, k[u].z, r[u].z, v[u].z, t[u].z, o[u].z);
            p.w = price_f32<MATH>(s[u].w, k[u].w, r[u].w, v[u].w, t[u].w, o[u].w);
            st_stream(p_out + gi, p);
            if (CHK) {
                if (err_bad(p.x, ref[u].x)) { bad++; err_note(ec, gi * 4 + 0); }
                if (err_bad(p.y, ref[u].y)) { bad++; err_note(ec, gi * 4 + 1); }
                if (err_bad(p.z, ref[u].z)) { bad++; err_note(ec, gi * 4 + 2); }
                if (err_bad(p.w, ref[u].w)) { bad++; err_note(ec, gi * 4 + 3); }
            }
        }
    }
    // leftover whole groups (fewer than UNROLL per thread)
    for (; g < groups; g += stride) {
        float4 s = ld_stream(p_s + g), k = ld_stream(p_k + g), r = ld_stream(p_r + g);
        float4 v = ld_stream(p_v + g), t = ld_stream(p_t + g);
        int4 o = ld_stream(p_o + g);
        float4 p;
        p.x = price_f32<MATH>(s.x, k.x, r.x, v.x, t.x, o.x);
        p.y = price_f32<MATH>(s.y, k.y, r.y, v.y, t.y, o.y);
        p.z = price_f32<MATH>(s.z, k.z, r.z, v.z, t.z, o.z);
        p.w = price_f32<MATH>(s.w, k.w, r.w, v.w, t.w, o.w);
        st_stream(p_out + g, p);
        if (CHK) {
            float4 ref = ld_stream(p_ref + g);
            if (err_bad(p.x, ref.x)) { bad++; err_note(ec, g * 4 + 0); }
            if (err_bad(p.y, ref.y)) { bad++; err_note(ec, g * 4 + 1); }
            if (err_bad(p.z, ref.z)) { bad++; err_note(ec, g * 4 + 2); }
            if (err_bad(p.w, ref.w)) { bad++; err_note(ec, g * 4 + 3); }
        }
    }
    // ragged tail: the last n % 4 options, one scalar option per thread of block 0
    const size_t tail0 = groups << 2;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < n) {
        const size_t i = tail0 + threadIdx.x;
        float p = price_f32<MATH>(a.spt[i], a.strike[i], a.rate[i], a.vol[i], a.otime[i], a.otype[i]);
        a.prices[i] = p;
        if (CHK && err_bad(p, a.refval[i])) { bad++; err_note(ec, i); }
    }
    if (CHK) err_flush(ec, bad);
}

// ---------------------------------------------------------------------------------------------
// The fp64 Map kernel.  One group = 2 consecutive options = one 128-bit access per fp stream
// (64-bit for otype).
// ---------------------------------------------------------------------------------------------
struct StreamsF64 {
    const double *spt, *strike, *rate, *vol, *otime;
    const int *otype;
    double *prices;
    const double *refval;
};

template <int MATH, int UNROLL, bool CHK>
__global__ void __launch_bounds__(256) bs_map_f64(StreamsF64 a, size_t n, ErrChk ec)
{
    const size_t groups = n >> 1;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;

    const double2 *p_s = reinterpret_cast<const double2 *>(a.spt);
    const double2 *p_k = reinterpret_cast<const double2 *>(a.strike);
    const double2 *p_r = reinterpret_cast<const double2 *>(a.rate);
    const double2 *p_v = reinterpret_cast<const double2 *>(a.vol);
    const double2 *p_t = reinterpret_cast<const double2 *>(a.otime);
    const int2 *p_o = reinterpret_cast<const int2 *>(a.otype);
    const double2 *p_ref = reinterpret_cast<const double2 *>(a.refval);
    double2 *p_out = reinterpret_cast<double2 *>(a.prices);

    for (; g + (UNROLL - 1) * stride < groups; g += UNROLL * stride) {
        double2 s[UNROLL], k[UNROLL], r[UNROLL], v[UNROLL], t[UNROLL], ref[UNROLL];
        int2 o[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g + u * stride;
            s[u] = ld_stream(p_s + gi);
            k[u] = ld_stream(p_k + gi);
            r[u] = ld_stream(p_r + gi);
            v[u] = ld_stream(p_v + gi);
            t[u] = ld_stream(p_t + gi);
            o[u] = ld_stream(p_o + gi);
            if (CHK) ref[u] = ld_stream(p_ref + gi);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t gi = g + u * stride;
            double2 p;
            p.x = price_f64_any<MATH>(s[u].x, k[u].x, r[u].x, v[u].x, t[u].x, o[u].x);
            p.y = price_f64_any<MATH>(s[u].y, k[u].y, r[u].y, v[u].y, t[u].y, o[u].y);
            st_stream(p_out + gi, p);
            if (CHK) {
                if (err_bad(p.x, ref[u].x)) { bad++; err_note(ec, gi * 2 + 0); }
                if (err_bad(p.y, ref[u].y)) { bad++; err_note(ec, gi * 2 + 1); }
            }
        }
    }
    for (; g < groups; g += stride) {
        double2 s = ld_stream(p_s + g), k = ld_stream(p_k + g), r = ld_stream(p_r + g);
        double2 v = ld_stream(p_v + g), t = ld_stream(p_t + g);
        int2 o = ld_stream(p_o + g);
        double2 p;
        p.x = price_f64_any<MATH>(s.x, k.x, r.x, v.x, t.x, o.x);
        p.y = price_f64_any<MATH>(s.y, k.y, r.y, v.y, t.y, o.y);
        st_stream(p_out + g, p);
        if (CHK) {
            double2 ref = ld_stream(p_ref + g);
            if (err_bad(p.x, ref.x)) { bad++; err_note(ec, g * 2 + 0); }
            if (err_bad(p.y, ref.y)) { bad++; err_note(ec, g * 2 + 1); }
        }
    }
    const size_t tail0 = groups << 1;
    if (blockIdx.x == 0 && tail0 + threadIdx.x < n) {
        const size_t i = tail0 + threadIdx.x;
        double p = price_f64_any<MATH>(a.spt[i], a.strike[i], a.rate[i], a.vol[i], a.otime[i], a.otype[i]);
        a.prices[i] = p;
        if (CHK && err_bad(p, a.refval[i])) { bad++; err_note(ec, i); }
    }
    if (CHK) err_flush(ec, bad);
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

}  // namespace bsk
