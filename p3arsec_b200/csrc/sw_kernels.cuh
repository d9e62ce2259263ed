/*
 * sw_kernels.cuh -- sm_100a kernels of the swaptions Map (HJM Monte-Carlo, fp64).
 *
 * Reference body: parsec-ff/pkgs/apps/swaptions/src/HJM_Swaption_Blocking.cpp:20-222 (cited as HSB: below) and the
 * PARSEC-owned leaves it calls (RanUnif, CumNormalInv, HJM_SimPath_Forward_Blocking, Discount_Factors_Blocking;
 * absent from the P3ARSEC overlay: their published PARSEC 3.0 semantics are followed).
 *
 * Work decomposition.  A trial is independent of every other trial: RanUnif's state is a counter and draw k is a
 * pure function of (seed + k), so trial t of a swaption reads draws [30 t, 30 t + 30) (iN = 11, iFactors = 3).
 * A work item = (swaption, chunk of THREADS x trials_per_thread consecutive-strided trials); persistent CTAs take
 * items round-robin (all items cost the same, so the static schedule is balanced to one item), every thread
 * simulates its trials one after the other and keeps two fp64 sums; one block reduction per item writes
 * partial[item] = {sum, sum of squares}; sw_finalize adds the partials of each swaption in a fixed tree and applies
 * HSB:212-214.  No atomics on the data path: the result is deterministic for a given geometry.
 *
 * These are FP64-pipe-bound kernels (no tensor cores: nothing is a contraction; ~1.2 KB of parameters per swaption
 * and 16 bytes of output per item, so HBM is idle).  Kernels, all built from the same device functions:
 *   sw_sim_fast<LEAN>  iN = 11, iFactors = 3; all (swaption, chunk) items of a device in ONE launch, the swaption's
 *                      tables in shared memory.  Per-trial state in registers, the trial's normals in shared memory
 *                      ([draw][thread], conflict-free).  LEAN = only what the price depends on.
 *   sw_sim_one<LEAN>   the same for ONE swaption per launch, its tables in the kernel-parameter constant bank, so that
 *                      the FP64 pipe gets them as uniform-register operands (a DFMA with three register-pair operands
 *                      issues every 3.07 cycles on B200, one with a constant operand every 2.07); the host uses it
 *                      for the full kernel when a swaption has enough trials to fill the GPU.
 *   sw_sim_generic     any shape, the reference's operation order (generic_trial), also the out-of-range fallback.
 *   sw_finalize        partial sums -> mean and standard error.
 * Building blocks: normals() = phase A (30 central-branch CumNormalInv, staged three at a time) + the deferred tail pass
 * (Moro's tail branch only for the draws that need it, through the composite table of sw_tail.h);
 * path_and_payoff<LEAN, START> = phase B, specialised on the swap start index.  DESIGN.md section 9 has the numbers.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "bs_math_f64.h"
#include "sw_tail.h"

// The per-trial arithmetic below is host-or-device source, like bs_math_f64.h: nvcc builds the device code; a plain C++
// compiler (g++ -ffp-contract=off, MUFU seeds emulated) builds the same functions for tools/sw_fast_host_check.cpp, which
// the CPU tests run against the oracle.  Kernels and warp-level code exist under nvcc only.
#if defined(__CUDACC__)
#define SW_HD __device__ __forceinline__
#define SW_HD_NOINLINE __device__ __noinline__
#define SW_CONST static __device__ __constant__
#define SW_HOST_DEVICE __host__ __device__
#else
#include <algorithm>
#define SW_HD inline
#define SW_HD_NOINLINE inline
#define SW_CONST static const
#define SW_HOST_DEVICE
#endif

namespace swk {

#if defined(__CUDACC__)
SW_HD int ffs32(uint32_t v) { return __ffs((int)v); }
SW_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
SW_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
SW_HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
#else
SW_HD int ffs32(uint32_t v) { return __builtin_ffs((int)v); }
SW_HD double mul_rn(double a, double b) { return a * b; }  // no contraction: the host build uses -ffp-contract=off
SW_HD double add_rn(double a, double b) { return a + b; }
SW_HD double div_rn(double a, double b) { return a / b; }
#endif

constexpr int MAXN = 32;  // SW_GPU_MAX_N
constexpr int MAXF = 8;   // SW_GPU_MAX_FACTORS
constexpr int THREADS = 128;

// Everything HJM_Swaption_Blocking derives from its arguments before the trial loop (HSB:48-61,116-148), computed
// on the host with the reference's own expressions, one record per swaption.
struct SwParams {
    double fwd[MAXN];            // HJM_Yield_to_Forward (HSB:143)
    double driftdt[MAXN];        // pdTotalDrift[l] * ddelt (HSB:148 + the product inside HJM_SimPath_Forward_Blocking)
    double fac[MAXF][MAXN];      // ppdFactors
    double pay[MAXN];            // pdSwapPayoffs, zero beyond iSwapVectorLength (HSB:132-140)
    double xd_path[MAXN];        // -(cumulative drift of path entry [j][0]) * ddelt       (fast kernels, see path_and_payoff)
    double xd_swap[MAXN];        // -(cumulative drift of path entry [start][i]) * swap_ddelt
    double ddelt;                // HSB:48
    double sqrt_ddelt;           // HJM_SimPath_Forward_Blocking
    double swap_ddelt;           // dSwapVectorYears / iSwapVectorLength (Discount_Factors_Blocking at HSB:184)
    double reserved;
    long long seed;              // swaption_seed + i (HJM_Securities.cpp:319)
    long long trials;            // lTrials
    long long sims;              // ceil(lTrials / BLOCKSIZE) * BLOCKSIZE (HSB:156)
    int start;                   // iSwapStartTimeIndex (HSB:125)
    int len;                     // iSwapVectorLength (HSB:116)
    int last_pay;                // iSwapTimePoints (HSB:126)
    int pad;
};

struct Geom {
    int iN, iFactors;
    int chunks;            // work items per swaption
    int items;             // chunks * swaptions on this device
    int tpt;               // trials per thread per item
    long long chunk_trials;  // THREADS * tpt
};

// ---- RanUnif (PARSEC RanUnif.c as restated in oracle/sw_absent/sw_leaves.c), literal 64-bit arithmetic ----------
SW_HD double ranunif_literal(long long ctr)
{
    long long ix = ctr;
    ix *= 1513517LL;
    ix %= 2147483647LL;
    long long k1 = ix / 127773LL;
    ix = 16807LL * (ix - k1 * 127773LL) - k1 * 2836LL;
    if (ix < 0) ix = ix + 2147483647LL;
    return (double)ix * 4.656612875e-10;
}

// Same draw for counters in [0, 2^40): x = ctr * 1513517 mod (2^31 - 1) kept as a 32-bit residue that advances by
// +1513517 per draw, and Schrage's step 16807 x mod (2^31 - 1) (exactly what the k1/127773/2836 dance computes
// for x in [0, 2^31 - 1)) done with one 64-bit product and a Mersenne fold.
constexpr uint32_t RU_M = 2147483647u;
constexpr int FD_MAX_DRAWS = 30;  // draws per trial of the fast kernels, (iN - 1) * iFactors = FD below
SW_HD uint32_t mersenne31(uint64_t p)  // p < 2^62
{
    uint64_t s = (p & RU_M) + (p >> 31);
    s = (s & RU_M) + (s >> 31);
    uint32_t r = (uint32_t)s;
    return r >= RU_M ? r - RU_M : r;
}
SW_HD uint32_t ru_residue(long long ctr) { return mersenne31((uint64_t)ctr * 1513517ull); }
SW_HD uint32_t ru_next(uint32_t x)
{
    x += 1513517u;
    return x >= RU_M ? x - RU_M : x;
}
SW_HD uint32_t ru_int(uint32_t x)  // the 31-bit integer of the draw; u = ru_int * 4.656612875e-10
{
    uint64_t p = (uint64_t)x * 16807ull;  // < 2^46
    uint32_t s = (uint32_t)(p & RU_M) + (uint32_t)(p >> 31);
    return s >= RU_M ? s - RU_M : s;
}

// Moro's coefficients; deliberately not const (see bs_tables_f64.h: const __constant__ doubles come back as
// 64-bit immediates that cost two UMOVs per use).
SW_CONST double MORO_A[4] = {2.50662823884, -18.61500062529, 41.39119773534, -25.44106049637};
SW_CONST double MORO_B[4] = {-8.47351093090, 23.08336743743, -21.06224101826, 3.13082909833};
SW_CONST double MORO_C[9] = {0.3374754822726147, 0.9761690190917186, 0.1607979714918209,
                                                   0.0276438810333863, 0.0038405729373609, 0.0003951896511919,
                                                   0.0000321767881768, 0.0000002888167364, 0.0000003960315187};

// ---- CumNormalInv in the reference's operation order, every operation rounded on its own --------------------------

SW_HD_NOINLINE double cumnormalinv_ieee(double u)
{
    double x = add_rn(u, -0.5);
    if (fabs(x) < 0.42) {
        double r = mul_rn(x, x);
        double num = add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(MORO_A[3], r), MORO_A[2]), r), MORO_A[1]), r), MORO_A[0]);
        double den = add_rn(
            mul_rn(add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(MORO_B[3], r), MORO_B[2]), r), MORO_B[1]), r), MORO_B[0]), r),
            1.0);
        return div_rn(mul_rn(x, num), den);
    }
    double r = u;
    if (x > 0.0) r = add_rn(1.0, -u);
    r = log(-log(r));
    double p = MORO_C[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) p = add_rn(MORO_C[i], mul_rn(r, p));
    return x < 0.0 ? -p : p;
}

// exp for the fast kernel.  exp_core = bsm::exp_f64 without its underflow clamp and without the second range-reduction step
// (9 FP64 operations), used for |x| < 2; everything else (rates of several hundred per cent, inf, NaN -- the latter reached
// only after a draw of exactly 0, see the tail pass) goes to libdevice, out of line, with the whole trial (generic_trial).  The range check is done on the high word with integer instructions: the FP64 pipe is
// the bottleneck of this kernel and a DSETP would cost it two issue cycles per exponential.
// Where exp_core reads 2^(j/64) from.  The index j is the low six bits of a rounded product, i.e. random per lane, so a
// plain 64-entry table in shared memory serves a half-warp's 64-bit loads in 2-4 wavefronts (16 bank pairs, 16 lanes
// with unrelated j).  REP = 16 copies interleaved [j][lane & 15] give every lane of a half-warp its own bank pair
// whatever j is: always one wavefront (8 KB per CTA).  REP = 1 is the plain table (host checker, generic fallback).
// scaled(n) = 2^(n >> 6) 2^((n & 63)/64) for n = round(x 64/ln2).  The replicated copy is BIASED (load_exp_replicated): entry
// j carries j * 2^14 less in its high word, so that adding n * 2^14 = (64 k + j) * 2^14 puts k into the exponent field and
// takes the bias out again -- one IMAD instead of shift, mask and add (twenty exponentials per trial).
template <int REP_SHIFT>
struct ExpTab {
    const double *t;
    int lane;
    SW_HD double at(int j) const { return t[(j << REP_SHIFT) + lane]; }
    SW_HD double scaled(int n) const
    {
#if defined(__CUDA_ARCH__)
        if (REP_SHIFT) {
            const double tb = at(n & 63);
            return __hiloint2double(n * 16384 + __double2hiint(tb), __double2loint(tb));
        }
#endif
        return bsm::scale_by_pow2(at(n & 63), n >> 6);
    }
};
constexpr int EXP_REP_SHIFT = 4, EXP_REP_DOUBLES = 64 << EXP_REP_SHIFT;
// The same for log_f64's {1/c, log c} pairs, read as one 128-bit load: a quarter-warp (8 lanes) per wavefront, so 8 copies
// interleaved [i][lane & 7] are conflict-free (8 KB per CTA).
template <int REP_SHIFT>
struct LogTabRep {
    const double *t;
    int lane;
    SW_HD void pair(int i, double &rc, double &lc) const
    {
#if defined(__CUDA_ARCH__)
        const double2 v = reinterpret_cast<const double2 *>(t)[(i << REP_SHIFT) + lane];
        rc = v.x;
        lc = v.y;
#else
        rc = t[2 * ((i << REP_SHIFT) + lane)];
        lc = t[2 * ((i << REP_SHIFT) + lane) + 1];
#endif
    }
};
constexpr int LOG_REP_SHIFT = 3, LOG_REP_DOUBLES = (64 << LOG_REP_SHIFT) * 2;
template <class ET>
SW_HD double exp_core(double x, const ET &et)
{
    using namespace bsm;
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51
    double nd = fma(x, kd(K_EXP_INV), MAGIC);
    const int n = (int)(uint32_t)to_bits(nd);  // round(x * 64/ln2) = 64 k + j
    nd -= MAGIC;
    // x - n ln2/64 with the 53-bit constant alone: the FMA rounds once, and what the low part of ln2/64 (3.6e-19) would add is
    // below 185 * 3.6e-19 = 6.7e-17 for the |x| < 2 this block is used for (EXP_HI_LIMIT) -- a third of an ulp of the result
    // at the very edge, 1e-17 for the |x| < 0.2 of real rates; bsm::exp_f64 keeps the second step for its |x| < 700.
    const double r = fma(nd, -kd(K_EXP_HI), x);
    double q = fma(kd(K_EXP_C5), r, kd(K_EXP_C4));
    q = fma(q, r, kd(K_EXP_C3));
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);
    const double T = et.scaled(n);   // 2^k 2^(j/64): a normal number, so scaling T first changes no bit of the result
    return fma(T, em1, T);
}
// |x| >= 2, inf or NaN <=> (high word of x, sign cleared) >= EXP_HI_LIMIT.  The fast kernel evaluates every exponential with
// exp_core, branch-free, and only tracks the high words of the arguments (integer pipes); a trial that exceeded the limit -- a
// rate above 100 % a year and more, or the inf / NaN that follow a draw of exactly 0 -- is redone by generic_trial().
// `worst` accumulates the high words with OR (a three-input LOP3 takes two exponentials at a time; a running maximum of the
// sign-cleared words cost a LOP3 and a VIMNMX each): the limit is a single bit, so the OR exceeds it iff one word does.
constexpr uint32_t EXP_HI_LIMIT = 0x40000000u;
template <class ET>
SW_HD double exp_tracked(double x, const ET &et, uint32_t &worst)
{
    worst |= (uint32_t)(bsm::to_bits(x) >> 32);
    return exp_core(x, et);
}
SW_HD bool exp_range_left(uint32_t worst) { return (worst & 0x7fffffffu) >= EXP_HI_LIMIT; }

// CumNormalInv takes its central branch iff fabs(u - 0.5) < 0.42 with u = s * 4.656612875e-10 (s the 31-bit draw).
// Both roundings are monotone in s, so the test is an integer range check: central <=> S_LO <= s <= S_HI
// (found by bisection over the double arithmetic and verified around both edges by tests/test_sw_oracle.py).
constexpr uint32_t S_LO = 171798692u, S_HI = 1975684955u;

// CumNormalInv's tail branch for draw k of the trial whose first residue is x0, branch-free so that several of them can
// be interleaved: z = -/+ P8(log(-log(min(u, 1 - u)))).  A draw of exactly 0 (counter a multiple of 2^31 - 1) gives
// log(-log(0)) = +inf in the reference, hence z = -inf.
// TWO_LOGS: P8(log(-log r)) as the reference writes it -- two logarithms and Moro's own coefficients (uniform operands),
// 20 FP64 operations and two table loads; otherwise one logarithm + the composite table of sw_tail.h, 9 FP64 operations and
// ten per-lane table loads after the logarithm's one.  The full-work kernel is bound by instruction issue and, with the
// composite table, by shared-memory bank conflicts of those per-lane loads (30 % of all wavefronts, profiles/
// r01_sw_ncu_one_full.txt); see SW_FULL_TWO_LOGS.
template <bool TWO_LOGS, class LT>
SW_HD double tail_normal(uint32_t x0, int k, const LT &lt, const double *tailtab)
{
    const uint32_t xk = x0 + (uint32_t)k * 1513517u;  // < 2^31 + 2^26, left unreduced: xk * 16807 < 2^46 and ru_int's fold is
    const uint32_t s = ru_int(xk);                    // below 2 (2^31 - 1), so its one conditional subtraction gives the exact draw
    const double u = (double)(int)s * 4.656612875e-10;
    const bool upper = s > S_HI;
    const double r = upper ? 1.0 - u : u;
    double p;
    if (TWO_LOGS) {
        const double w = bsm::log_f64_t(-bsm::log_f64_t(r, lt), lt);
        p = fma(w, MORO_C[8], MORO_C[7]);
        p = fma(w, p, MORO_C[6]);
        p = fma(w, p, MORO_C[5]);
        p = fma(w, p, MORO_C[4]);
        p = fma(w, p, MORO_C[3]);
        p = fma(w, p, MORO_C[2]);
        p = fma(w, p, MORO_C[1]);
        p = fma(w, p, MORO_C[0]);
    } else {
        p = swt::moro_tail_t(r, lt, tailtab);  // P8(log(-log r)): one logarithm + the composite table (sw_tail.h)
    }
    // (a draw of exactly 0 -- log(-log 0) = +inf in the reference -- never gets here with a meaning: trial_draws_zero() sends
    // such a trial to generic_trial() as a whole; what this function returns for it is finite garbage.)
    // sign: -p for the lower tail.  S_HI - s is negative exactly for the upper tail (both below 2^31): its sign bit cancels the flip.
    return bsm::from_bits(bsm::to_bits(p) ^ ((uint64_t)(~(S_HI - s) & 0x80000000u) << 32));
}

// Does one of the trial's FD draws come out as exactly 0?  Draw k is 0 iff its residue x0 + k c is a multiple of 2^31 - 1
// (c = 1513517, x0 = ru_residue(first counter) in [0, 2^31 - 1)), i.e. iff 2^31 - 1 - x0 (or x0 itself, for k = 0) is one of the
// first FD multiples of c: a range test that fails for 98 % of the trials, then one remainder.  The reference carries the
// resulting -inf through the path; the fast kernels hand the whole trial to generic_trial(), once per trial instead of a
// compare and a 64-bit select per tail draw.
SW_HD bool trial_draws_zero(uint32_t x0)
{
    const uint32_t d = x0 == 0 ? 0u : RU_M - x0;
    if (d > (uint32_t)(FD_MAX_DRAWS - 1) * 1513517u) return false;
    return d % 1513517u == 0;
}

#if defined(__CUDACC__)
// The exp/log tables and the tail table, [tail | tab], as the kernels' shared-memory blocks start with them.  They are
// expanded once per device into global memory by sw_fill_tables; every CTA then copies them with coalesced loads (the
// per-thread-indexed reads of the __constant__ originals serialise in the constant cache: ~10 us per launch, which is
// what a one-swaption launch or a small portfolio cannot afford).
constexpr int TABLE_DOUBLES = swt::TAIL_DOUBLES + bsm::TAB_DOUBLES;
__global__ void sw_fill_tables(double *__restrict__ g)
{
    swt::fill_tail(g, threadIdx.x, blockDim.x);
    bsm::fill_tables(g + swt::TAIL_DOUBLES, threadIdx.x, blockDim.x);
}
SW_HD void load_tables(double *__restrict__ smem_tail_then_tab, const double *__restrict__ g, int tid)
{
    for (int i = tid; i < TABLE_DOUBLES; i += THREADS) smem_tail_then_tab[i] = g[i];
}
SW_HD void load_exp_replicated(double *__restrict__ xexp, const double *__restrict__ g, int tid)
{
    for (int i = tid; i < EXP_REP_DOUBLES; i += THREADS) {  // biased by j * 2^14 in the high word (ExpTab::scaled)
        const int j = i >> EXP_REP_SHIFT;
        xexp[i] = bsm::from_bits(bsm::to_bits(g[swt::TAIL_DOUBLES + bsm::TAB_EXP + j]) - ((uint64_t)j << 46));
    }
}
SW_HD void load_log_replicated(double *__restrict__ xlog, const double *__restrict__ g, int tid)
{
    // element e = 2 * ((i << LOG_REP_SHIFT) + lane) + c  <-  table pair i, component c
    for (int e = tid; e < LOG_REP_DOUBLES; e += THREADS)
        xlog[e] = g[swt::TAIL_DOUBLES + bsm::TAB_LOG + 2 * (e >> (LOG_REP_SHIFT + 1)) + (e & 1)];
}

// Block-wide sum of two doubles; result valid in thread 0.
SW_HD void block_sum2(double &a, double &b, double (*red)[THREADS / 32])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[0][w] = a;
        red[1][w] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = red[0][0];
        b = red[1][0];
#pragma unroll
        for (int i = 1; i < THREADS / 32; ++i) {
            a += red[0][i];
            b += red[1][i];
        }
    }
}
#endif  // __CUDACC__

// =====================================================================================================================
// One trial in the reference's operation order: libdevice exp/log, IEEE divide, every operation rounded on its own
// (no FMA contraction), any iN <= 32 and iFactors <= 8, path rows in local memory.  Body of sw_sim_generic and the
// out-of-line fallback of sw_sim_fast for trials whose exponentials leave the range of its fast arithmetic.
// Returns the discounted payoff of trial t (HSB:198).
// =====================================================================================================================
SW_HD_NOINLINE double generic_trial(const SwParams &P, int iN, int nF, long long t)
{
    long long ctr = P.seed + t * ((long long)(iN - 1) * nF);
    double row[MAXN], srow[MAXN], z[MAXF];
    for (int l = 0; l < iN; ++l) {
        row[l] = P.fwd[l];
        srow[l] = row[l];
    }
    double run = 1.0, pay_df = 1.0;
    for (int j = 1; j <= iN - 1; ++j) {
        run = mul_rn(run, exp(mul_rn(-row[0], P.ddelt)));
        if (j == P.start) pay_df = run;
        for (int i = 0; i < nF; ++i) z[i] = cumnormalinv_ieee(ranunif_literal(ctr++));
        for (int l = 0; l <= iN - 1 - j; ++l) {
            double shock = 0.0;
            for (int i = 0; i < nF; ++i) shock = add_rn(shock, mul_rn(P.fac[i][l], z[i]));
            row[l] = add_rn(add_rn(row[l + 1], P.driftdt[l]), mul_rn(P.sqrt_ddelt, shock));
        }
        row[iN - j] = 0.0;
        if (j == P.start)
            for (int l = 0; l < iN; ++l) srow[l] = row[l];
    }
    double df = 1.0, fixed = 0.0;
    for (int i = 0; i <= P.len - 1; ++i) {
        if (i >= 1) df = mul_rn(df, exp(mul_rn(-srow[i - 1], P.swap_ddelt)));
        fixed = add_rn(fixed, mul_rn(P.pay[i], df));
    }
    const double fm1 = add_rn(fixed, -1.0);
    const double payoff = fm1 > 0.0 ? fm1 : 0.0;   // dMax (HSB:196)
    return mul_rn(payoff, pay_df);                 // HSB:198
}

// =====================================================================================================================
// sw_sim_fast: iN = 11, iFactors = 3 (the only shape the reference drivers create: HJM_Securities.cpp:56,58).
// =====================================================================================================================
constexpr int FN = 11, FF = 3, FD = (FN - 1) * FF;
static_assert(FD == FD_MAX_DRAWS, "trial_draws_zero() covers the draws of one trial");

struct FastShared {
    double tail[swt::TAIL_DOUBLES];  // composite table of Moro's tail branch (5 KB; rows of 80 B, 16-byte aligned)
    double tab[bsm::TAB_DOUBLES];
    double4 fd[FN - 1];        // per maturity l: {fac0, fac1, fac2} * sqrt_ddelt and pdTotalDrift[l] * ddelt (two LDS.128)
    double fwd[FN];
    double pay[FN];
    double xdp[FN];            // SwParams::xd_path
    double xds[FN];            // SwParams::xd_swap
    double red[2][THREADS / 32];
    double xexp[EXP_REP_DOUBLES];  // 2^(j/64) replicated [j][lane & 15]: conflict-free whatever the lanes' j (ExpTab)
    double xlog[LOG_REP_DOUBLES];  // {1/c, log c} replicated [i][lane & 7] (LogTabRep)
};
// Behind FastShared in dynamic shared memory: z[z_rows][THREADS], the trial's normals [draw][thread] (conflict-free) --
// all 30 draws for the full kernel, 3 x (largest swap start index of the launch) for the lean one.
static_assert(sizeof(FastShared) % 16 == 0, "z starts 16-byte aligned");
SW_HOST_DEVICE constexpr size_t fast_shared_bytes(int z_rows) { return sizeof(FastShared) + (size_t)z_rows * THREADS * sizeof(double); }

// ---- phase A + tail pass -----------------------------------------------------------------------------------------------
// The trial's normals into sh.z[draw][tid].  Central branch of CumNormalInv for every draw, G draws at a time written stage
// by stage so that G independent Horner chains are in flight (the FP64 pipe has a long dependent-issue latency and only
// four warps per scheduler to hide it).  Draws that belong to the tail branch are noted in a mask and redone afterwards,
// three per trip: a warp pays for max-over-lanes(#tail draws) tail evaluations, not for 30.
// The phase needs no exact residues: x0 + k c is left unreduced (< 2^31 + 2^26, so x * 16807 < 2^47 and the fold still
// gives a value congruent to the draw) and so is the folded sum, s in [0, 2^31 - 1 + 2^16): an unreduced s >= 2^31 - 1
// stands for a draw below 2^16, i.e. a tail draw, fails the range check like one and is recomputed exactly by the tail pass.
// Tuning knobs of the full kernel (measured on B200, 128 x 1M trials: every combination of 2/3/6 tail draws per trip,
// 3/6/10 draws per phase-A group, row-/coefficient-major tail table or two logarithms lands within 11.0-11.3 G trials/s;
// profiles/r01_sw_variants.txt).
#ifndef SW_TAIL_TRIP
#define SW_TAIL_TRIP 6
#endif
#ifndef SW_PAIRED_RCP
#define SW_PAIRED_RCP 0  /* 1: phase A shares one reciprocal (one MUFU seed + Newton step) per pair of central draws.  Measured
                            neutral on B200 in the first half of round 2 (12.56-12.62 against 12.62 G trials/s native: 15 MUFU.RCP64H
                            fewer per trial, the same FP64 count, the same time -- the seeds were not what the dispatch port was
                            waiting for) and 2.5 % slower with the final kernel (13.37 against 13.71), so off. */
#endif
#ifndef SW_PHASE_A_GROUP
#define SW_PHASE_A_GROUP 3  /* draws evaluated side by side in phase A: 3, 6, 10, 15 or 30.  Measured with the final kernel on one board,
                               alternating (profiles/r02_sw_tune_driftfree.txt): 13.84 (3) / 13.71 (6) / 13.72 (10) / 12.38 (15, spills)
                               G trials/s; round 1 had 3 / 6 / 10 within noise of each other. */
#endif
constexpr int TAIL_TRIP = SW_TAIL_TRIP;

// Measured on B200, PARSEC native (profiles/r02_sw_tail_variants.txt): composite table 12.62 G trials/s with 30 % of the
// shared-memory wavefronts being bank conflicts (the ten per-lane coefficient loads of each tail: lanes hold unrelated
// rows); two logarithms 12.12 G trials/s with 0.09 % conflicts (+60 FP64 instructions per trial, longer dependent
// chains: `wait` stalls 1.09 -> 1.52 per issue).  Throughput is the metric, so the composite table stays the default; the
// exp and log tables, which could be replicated per lane in 16 KB, are conflict-free in both.
#ifndef SW_FULL_TWO_LOGS
#define SW_FULL_TWO_LOGS 0  /* 1: full-work kernels evaluate Moro's tail with two logarithms (no composite-table loads) */
#endif
template <bool LEAN, class LT>
SW_HD void normals(const LT &tab, const double *__restrict__ tailtab, double *__restrict__ z, int tid, uint32_t x0,
                                        int steps)
{
    constexpr bool TWO_LOGS = !LEAN && SW_FULL_TWO_LOGS;
    constexpr int G = LEAN ? FF : SW_PHASE_A_GROUP;
    uint32_t tail = 0;
#pragma unroll
    for (int k0 = 0; k0 < FD; k0 += G) {
        if (LEAN && k0 >= FF * steps) break;
        uint32_t sg[G];
        double xc[G], r[G], num[G], den[G], q[G], rx[G], e[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const uint64_t p = (uint64_t)(x0 + (uint32_t)(k0 + i) * 1513517u) * 16807ull;
            sg[i] = (uint32_t)(p & RU_M) + (uint32_t)(p >> 31);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) xc[i] = fma((double)(int)sg[i], 4.656612875e-10, -0.5);  // u - 0.5, one rounding
#pragma unroll
        for (int i = 0; i < G; ++i) r[i] = xc[i] * xc[i];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            num[i] = fma(MORO_A[3], r[i], MORO_A[2]);
            den[i] = fma(MORO_B[3], r[i], MORO_B[2]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            num[i] = fma(num[i], r[i], MORO_A[1]);
            den[i] = fma(den[i], r[i], MORO_B[1]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            num[i] = fma(num[i], r[i], MORO_A[0]);
            den[i] = fma(den[i], r[i], MORO_B[0]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            den[i] = fma(den[i], r[i], 1.0);
            q[i] = xc[i] * num[i];
        }
        if (SW_PAIRED_RCP && (G % 2) == 0) {
            // one reciprocal per PAIR of draws: 1/den_a = r den_b, 1/den_b = r den_a with r = 1/(den_a den_b).  The MUFU seed
            // is the expensive instruction here (6-8 dispatch cycles when dense, profiles/r01_dispatch_cost_microbench.txt);
            // a pair costs the same eight FP64 instructions either way and one seed less.  den is in [0.11, 1] on the
            // central branch (and garbage-but-finite for tail draws, whose z is overwritten below).
            double dd[G / 2], rr[G / 2];
#pragma unroll
            for (int i = 0; i < G / 2; ++i) dd[i] = den[2 * i] * den[2 * i + 1];
#pragma unroll
            for (int i = 0; i < G / 2; ++i) rr[i] = bsm::seed_rcp(dd[i]);    // MUFU.RCP64H
#pragma unroll
            for (int i = 0; i < G / 2; ++i) e[i] = fma(-dd[i], rr[i], 1.0);  // one cubic Newton step: x (1 + e + e^2)
#pragma unroll
            for (int i = 0; i < G / 2; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
            for (int i = 0; i < G / 2; ++i) rr[i] = fma(rr[i], e[i], rr[i]);
#pragma unroll
            for (int i = 0; i < G / 2; ++i) {
                rx[2 * i] = rr[i] * den[2 * i + 1];
                rx[2 * i + 1] = rr[i] * den[2 * i];
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                z[(k0 + i) * THREADS + tid] = q[i] * rx[i];  // garbage for tail draws: overwritten below
                if ((sg[i] - S_LO) > (S_HI - S_LO)) tail |= 1u << (k0 + i);
            }
        } else {
#pragma unroll
            for (int i = 0; i < G; ++i) rx[i] = bsm::seed_rcp(den[i]);       // MUFU.RCP64H
#pragma unroll
            for (int i = 0; i < G; ++i) e[i] = fma(-den[i], rx[i], 1.0);     // one cubic Newton step: x (1 + e + e^2)
#pragma unroll
            for (int i = 0; i < G; ++i) {
                e[i] = fma(e[i], e[i], e[i]);
                q[i] = q[i] * rx[i];
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                z[(k0 + i) * THREADS + tid] = fma(q[i], e[i], q[i]);  // garbage for tail draws: overwritten below
                if ((sg[i] - S_LO) > (S_HI - S_LO)) tail |= 1u << (k0 + i);
            }
        }
    }
    while (tail) {
        const int k0 = ffs32(tail) - 1;
        tail &= tail - 1;
        if (LEAN) {  // three or six draws per trial: rarely more than one tail draw per lane (two per trip measured 8 % slower)
            z[k0 * THREADS + tid] = tail_normal<TWO_LOGS>(x0, k0, tab, tailtab);
        } else {
            // TRIP draws per trip with interleaved chains; with fewer left the last trip repeats draw k0
            int kk[TAIL_TRIP];
            double zz[TAIL_TRIP];
            kk[0] = k0;
#pragma unroll
            for (int i = 1; i < TAIL_TRIP; ++i) {
                kk[i] = tail ? ffs32(tail) - 1 : k0;
                tail &= tail - 1;  // (0 & anything == 0)
            }
#pragma unroll
            for (int i = 0; i < TAIL_TRIP; ++i) zz[i] = tail_normal<TWO_LOGS>(x0, kk[i], tab, tailtab);
#pragma unroll
            for (int i = 0; i < TAIL_TRIP; ++i) z[kk[i] * THREADS + tid] = zz[i];
        }
    }
}

// ---- phase B --------------------------------------------------------------------------------------------------------------
// The forward-rate path row by row in registers (HJM_SimPath_Forward_Blocking), column 0 feeding the payoff discount
// factor (HSB:167-172), row `start` feeding the swap leg (HSB:179-195); returns the discounted payoff (HSB:198).
// The rows are kept WITHOUT their drift: entry [j][l] of the reference's path is R_j[l] + cd_j[l] with the shocks-only recursion
// R_j[l] = R_{j-1}[l+1] + sqrt(ddelt) sum_i fac_i[l] z_{j,i} (three FMAs on the previous row's entry, no addition of its own)
// and the cumulative drift cd_j[l] = cd_{j-1}[l+1] + pdTotalDrift[l] ddelt, which is the same for every trial.  The path is
// only ever read as an argument of exp(-rate * dt) -- column 0 of every row and the start row -- so prepare() folds cd into
// those twenty arguments (xd_path, xd_swap) and the multiplication -rate * dt becomes an FMA: 55 FP64 additions less per trial.
// START >= 0: the swap start index as a compile-time constant -- the row snapshot is then a register renaming and the
// whole phase is one basic block; START < 0: taken from start_rt.  Every exponential is evaluated branch-free; `worst`
// remembers whether one of them left the fast range.
// SRC = where the swaption's tables (fd, fwd, pay) are read from: FastShared (shared memory, any swaption per work item)
// or OneSwaption (kernel-parameter constant bank, one swaption per launch).
template <bool LEAN, int START, class SRC, class ET>
SW_HD double path_and_payoff(const SRC &sh, const ET &tab, const double *__restrict__ z, int tid, double ddelt,
                                                  double swap_ddelt, int start_rt, int swap_end, uint32_t &worst)
{
    const int start = START >= 0 ? START : start_rt;
    const int steps = LEAN ? start : FN - 1;  // time steps whose shocks are needed
    double row[FN], srow[FN];
#pragma unroll
    for (int l = 0; l < FN; ++l) {
        row[l] = sh.fwd[l];
        srow[l] = row[l];  // start == 0
    }
    double run = 1.0, pay_df = 1.0;
#pragma unroll
    for (int j = 1; j <= FN - 1; ++j) {
        if (!LEAN || j <= steps) {
            // Discount_Factors_Blocking: DF[j] = DF[j-1] exp(-(rate of row j-1, column 0) * ddelt)
            run *= exp_tracked(fma(-row[0], ddelt, sh.xdp[j - 1]), tab, worst);
            const double z0 = z[(FF * (j - 1) + 0) * THREADS + tid];
            const double z1 = z[(FF * (j - 1) + 1) * THREADS + tid];
            const double z2 = z[(FF * (j - 1) + 2) * THREADS + tid];
            // lean: of row j only the entries that can still reach column 0 by step `start` or the swap leg's rates
            // are needed: l <= (start - j) + swap_end - 1
            const int need = LEAN ? start - j + swap_end - 1 : FN;
#pragma unroll
            for (int l = 0; l <= FN - 1 - j; ++l) {
                if (!LEAN || l <= need) {
                    const double4 c = sh.fd[l];
                    double v = fma(c.x, z0, row[l + 1]);
                    v = fma(c.y, z1, v);
                    row[l] = fma(c.z, z2, v);
                }
            }
            row[FN - j] = 0.0;  // the reference's path matrix is zero beyond the triangle
            if (j == start) {
                pay_df = run;
#pragma unroll
                for (int l = 0; l < FN; ++l) srow[l] = row[l];
            }
        }
    }
    // swap leg: DF[i] = prod_{k<i} exp(-srow[k] swap_ddelt); fixed leg = sum pay[i] DF[i] (HSB:184-195).  pay[i] is zero
    // beyond the last payment and srow is zero beyond the row's length, so the full kernel sums all ten terms without a
    // branch (0 * finite = 0; a non-finite factor sends the trial to generic_trial).
    double df = 1.0, fixed = 0.0;
#pragma unroll
    for (int i = 1; i <= FN - 1; ++i) {
        if (!LEAN || i <= swap_end) {
            df *= exp_tracked(fma(-srow[i - 1], swap_ddelt, sh.xds[i - 1]), tab, worst);
            fixed = fma(sh.pay[i], df, fixed);
        }
    }
    const double payoff = fixed - 1.0 > 0.0 ? fixed - 1.0 : 0.0;  // dMax (HSB:196)
    return payoff * pay_df;                                       // HSB:198
}

// Four CTAs (16 warps) per SM: 128 registers per thread.  Builds bounded for 5 and 6 CTAs (96 / 80 registers) spill
// the path rows and measured 4-12 % slower.
#ifndef SW_LEAN_MINB
#define SW_LEAN_MINB 5  /* lean kernel: 5 CTAs per SM (96 registers) measured best: 46.6 / 48.7 / 45.9 G trials/s for 4 / 5 / 8 */
#endif
#if defined(__CUDACC__)
template <bool LEAN>
__global__ void __launch_bounds__(THREADS, LEAN ? SW_LEAN_MINB : 4)
sw_sim_fast(const SwParams *__restrict__ params, const Geom g, double2 *__restrict__ partials, const double *__restrict__ tables)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastShared &sh = *reinterpret_cast<FastShared *>(smem_raw);
    double *const z = reinterpret_cast<double *>(smem_raw + sizeof(FastShared));
    const int tid = threadIdx.x;
    static_assert(offsetof(FastShared, tail) == 0 && offsetof(FastShared, tab) == sizeof(double) * swt::TAIL_DOUBLES, "[tail | tab] first");
    load_tables(sh.tail, tables, tid);
    load_exp_replicated(sh.xexp, tables, tid);
    load_log_replicated(sh.xlog, tables, tid);
    const ExpTab<EXP_REP_SHIFT> et = {sh.xexp, tid & ((1 << EXP_REP_SHIFT) - 1)};
    const LogTabRep<LOG_REP_SHIFT> lt = {sh.xlog, tid & ((1 << LOG_REP_SHIFT) - 1)};

    int cur = -1;
    double ddelt = 0, swap_ddelt = 0;
    long long seed = 0, sims = 0;
    int start = 0, len = 0, last_pay = 0;

    for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
        const int sw = item / g.chunks, chunk = item - sw * g.chunks;
        __syncthreads();  // previous item's reduction and parameter reads are done
        if (sw != cur) {
            const SwParams &P = params[sw];
            if (tid < FN) {
                sh.fwd[tid] = P.fwd[tid];
                sh.pay[tid] = P.pay[tid];
                sh.xdp[tid] = P.xd_path[tid];
                sh.xds[tid] = P.xd_swap[tid];
            }
            if (tid >= 32 && tid < 32 + FN - 1) {
                const int l = tid - 32;
                const double sq = P.sqrt_ddelt;
                sh.fd[l] = make_double4(P.fac[0][l] * sq, P.fac[1][l] * sq, P.fac[2][l] * sq, P.driftdt[l]);
            }
            ddelt = P.ddelt;
            swap_ddelt = P.swap_ddelt;
            seed = P.seed;
            sims = P.sims;
            start = P.start;
            len = P.len;
            last_pay = P.last_pay;
            cur = sw;
        }
        __syncthreads();

        const int steps = LEAN ? start : FN - 1;         // time steps whose shocks are needed
        const int swap_end = LEAN ? last_pay : len - 1;  // last swap discount factor needed
        double sum = 0.0, sumsq = 0.0;

        for (int m = 0; m < g.tpt; ++m) {
            const long long t = (long long)chunk * g.chunk_trials + (long long)m * THREADS + tid;
            if (t >= sims) break;

            // ---- phase A + tail pass: the trial's normals into sh.z (see normals())
            const uint32_t x0 = ru_residue(seed + t * FD);
            normals<LEAN>(lt, sh.tail, z, tid, x0, steps);

            // ---- phase B: path, discount factors, payoff; specialised on the swap start index (1..3 covers every
            // swaption the reference drivers create: dMaturity = 1, dYears in [5, 20))
            uint32_t worst = trial_draws_zero(x0) ? EXP_HI_LIMIT : 0u;
            double disc;
            switch (start) {
                case 1: disc = path_and_payoff<LEAN, 1>(sh, et, z, tid, ddelt, swap_ddelt, start, swap_end, worst); break;
                case 2: disc = path_and_payoff<LEAN, 2>(sh, et, z, tid, ddelt, swap_ddelt, start, swap_end, worst); break;
                case 3: disc = path_and_payoff<LEAN, 3>(sh, et, z, tid, ddelt, swap_ddelt, start, swap_end, worst); break;
                default: disc = path_and_payoff<LEAN, -1>(sh, et, z, tid, ddelt, swap_ddelt, start, swap_end, worst); break;
            }
            if (exp_range_left(worst)) disc = generic_trial(params[sw], FN, FF, t);
            sum += disc;                                                  // HSB:203
            sumsq = fma(disc, disc, sumsq);                               // HSB:204
        }

        block_sum2(sum, sumsq, sh.red);
        if (tid == 0) partials[item] = make_double2(sum, sumsq);
    }
}
#endif  // __CUDACC__

// =====================================================================================================================
// sw_sim_one: the same trial code for ONE swaption per launch, its tables in the kernel-parameter constant bank.
// A DFMA whose three operands are three different register pairs issues every 3.07 scheduler cycles on B200, one with a
// constant-bank operand every 2.07 (profiles/r01_dfma_operands_microbench.txt).  In sw_sim_fast the factor table comes from
// shared memory, so the 165 shock FMAs of a trial are of the slow kind and need 110 LDS.128 on top; here fd[l], fwd[l]
// and pay[i] are immediates of the instruction (c[0x0][imm]).  Used when a swaption has enough trials to fill the GPU.
// =====================================================================================================================
struct OneSwaption {
    double4 fd[FN - 1];  // {fac0, fac1, fac2} * sqrt_ddelt and pdTotalDrift * ddelt per maturity
    double fwd[FN];
    double pay[FN];
    double xdp[FN];      // SwParams::xd_path
    double xds[FN];      // SwParams::xd_swap
    double ddelt, swap_ddelt;
    long long seed, sims, chunk_trials;
    int start, len, last_pay, tpt;
    int chunks, partial_base, sw_index, pad;
};

struct OneShared {
    double tail[swt::TAIL_DOUBLES];
    double tab[bsm::TAB_DOUBLES];
    double red[2][THREADS / 32];
    double xexp[EXP_REP_DOUBLES];  // 2^(j/64) replicated [j][lane & 15] (ExpTab)
    double xlog[LOG_REP_DOUBLES];  // {1/c, log c} replicated [i][lane & 7] (LogTabRep)
};
static_assert(sizeof(OneShared) % 16 == 0, "z starts 16-byte aligned");
SW_HOST_DEVICE constexpr size_t one_shared_bytes(int z_rows) { return sizeof(OneShared) + (size_t)z_rows * THREADS * sizeof(double); }

#if defined(__CUDACC__)
template <bool LEAN>
__global__ void __launch_bounds__(THREADS, LEAN ? SW_LEAN_MINB : 4)
sw_sim_one(const __grid_constant__ OneSwaption P, const SwParams *__restrict__ params, double2 *__restrict__ partials,
           const double *__restrict__ tables)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OneShared &sh = *reinterpret_cast<OneShared *>(smem_raw);
    double *const z = reinterpret_cast<double *>(smem_raw + sizeof(OneShared));
    const int tid = threadIdx.x;
    static_assert(offsetof(OneShared, tail) == 0 && offsetof(OneShared, tab) == sizeof(double) * swt::TAIL_DOUBLES, "[tail | tab] first");
    load_tables(sh.tail, tables, tid);
    load_exp_replicated(sh.xexp, tables, tid);
    load_log_replicated(sh.xlog, tables, tid);
    const ExpTab<EXP_REP_SHIFT> et = {sh.xexp, tid & ((1 << EXP_REP_SHIFT) - 1)};
    const LogTabRep<LOG_REP_SHIFT> lt = {sh.xlog, tid & ((1 << LOG_REP_SHIFT) - 1)};
    const int steps = LEAN ? P.start : FN - 1;
    const int swap_end = LEAN ? P.last_pay : P.len - 1;

    for (int chunk = blockIdx.x; chunk < P.chunks; chunk += gridDim.x) {
        __syncthreads();  // tables filled / previous chunk's reduction read
        double sum = 0.0, sumsq = 0.0;
        for (int m = 0; m < P.tpt; ++m) {
            const long long t = (long long)chunk * P.chunk_trials + (long long)m * THREADS + tid;
            if (t >= P.sims) break;
            const uint32_t x0 = ru_residue(P.seed + t * FD);
            normals<LEAN>(lt, sh.tail, z, tid, x0, steps);
            uint32_t worst = trial_draws_zero(x0) ? EXP_HI_LIMIT : 0u;
            double disc;
            switch (P.start) {
                case 1: disc = path_and_payoff<LEAN, 1>(P, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
                case 2: disc = path_and_payoff<LEAN, 2>(P, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
                case 3: disc = path_and_payoff<LEAN, 3>(P, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
                default: disc = path_and_payoff<LEAN, -1>(P, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
            }
            if (exp_range_left(worst)) disc = generic_trial(params[P.sw_index], FN, FF, t);
            sum += disc;                     // HSB:203
            sumsq = fma(disc, disc, sumsq);  // HSB:204
        }
        block_sum2(sum, sumsq, sh.red);
        if (tid == 0) partials[P.partial_base + chunk] = make_double2(sum, sumsq);
    }
}

// =====================================================================================================================
// sw_sim_generic: any iN <= 32, iFactors <= 8, one generic_trial() per thread at a time.
// =====================================================================================================================
__global__ void __launch_bounds__(THREADS)
sw_sim_generic(const SwParams *__restrict__ params, const Geom g, double2 *__restrict__ partials)
{
    __shared__ double red[2][THREADS / 32];
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
        const int sw = item / g.chunks, chunk = item - sw * g.chunks;
        const SwParams &P = params[sw];
        __syncthreads();
        double sum = 0.0, sumsq = 0.0;
        for (int m = 0; m < g.tpt; ++m) {
            const long long t = (long long)chunk * g.chunk_trials + (long long)m * THREADS + tid;
            if (t >= P.sims) break;
            const double disc = generic_trial(P, g.iN, g.iFactors, t);
            sum = add_rn(sum, disc);                      // HSB:203
            sumsq = add_rn(sumsq, mul_rn(disc, disc));    // HSB:204
        }
        block_sum2(sum, sumsq, red);
        if (tid == 0) partials[item] = make_double2(sum, sumsq);
    }
}

// One CTA per swaption: add its partials in a fixed order, then HSB:212-214.
__global__ void __launch_bounds__(THREADS)
sw_finalize(const SwParams *__restrict__ params, const Geom g, const double2 *__restrict__ partials,
            double *__restrict__ mean, double *__restrict__ err)
{
    __shared__ double red[2][THREADS / 32];
    const int sw = blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int c = threadIdx.x; c < g.chunks; c += THREADS) {
        const double2 p = partials[(size_t)sw * g.chunks + c];
        a += p.x;
        b += p.y;
    }
    block_sum2(a, b, red);
    if (threadIdx.x == 0) {
        const double n = (double)params[sw].trials;
        mean[sw] = __ddiv_rn(a, n);
        const double var = __ddiv_rn(__dadd_rn(b, -__ddiv_rn(__dmul_rn(a, a), n)), __dadd_rn(n, -1.0));
        err[sw] = __ddiv_rn(__dsqrt_rn(var), __dsqrt_rn(n));
    }
}

#endif  // __CUDACC__

}  // namespace swk
