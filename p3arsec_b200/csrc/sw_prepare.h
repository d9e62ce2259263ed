/*
 * sw_prepare.h -- host side of the swaptions Map: everything HJM_Swaption_Blocking computes before its trial loop
 * (parsec-ff/pkgs/apps/swaptions/src/HJM_Swaption_Blocking.cpp:48-61,116-148, cited as HSB:), evaluated with the
 * reference's own expressions and packed into one SwParams record per swaption.  Plain C++ (no CUDA): included by
 * sw_gpu.cu and by the host checker tools/sw_fast_host_check.cpp.
 */
#pragma once
#include <cmath>
#include <cstring>

#include "../../include/sw_gpu.h"
#include "sw_kernels.cuh"

namespace swk {

// Everything HJM_Swaption_Blocking derives before the trial loop.  Returns false where the reference would index
// outside its vectors (it has no checks of its own).
inline bool prepare(SwParams &P, const sw_gpu_swaption &s, int iN, int iFactors, const double *pdYield,
             const double *ppdFactors, long seed, long lTrials, int BLOCKSIZE)
{
    memset(&P, 0, sizeof(P));
    const double ddelt = (double)(s.dYears / iN);                                  // HSB:48
    if (!(ddelt > 0.0) || !std::isfinite(ddelt)) return false;
    const double fr = s.dPaymentInterval / ddelt + 0.5, st = s.dMaturity / ddelt + 0.5, tp = s.dTenor / ddelt + 0.5;
    const double vl = iN - s.dMaturity / ddelt + 0.5;
    if (!(fabs(fr) < 1e6 && fabs(st) < 1e6 && fabs(tp) < 1e6 && fabs(vl) < 1e6)) return false;
    const int iFreqRatio = (int)fr;                                                // HSB:50
    double dStrikeCont;
    if (s.dCompounding == 0) dStrikeCont = s.dStrike;                              // HSB:56-57
    else dStrikeCont = (1 / s.dCompounding) * log(1 + s.dStrike * s.dCompounding);  // HSB:61
    const int iSwapVectorLength = (int)vl;                                         // HSB:116
    const int iSwapStartTimeIndex = (int)st;                                       // HSB:125
    const int iSwapTimePoints = (int)tp;                                           // HSB:126
    const double dSwapVectorYears = (double)(iSwapVectorLength * ddelt);           // HSB:127
    if (iSwapVectorLength < 1 || iSwapVectorLength > iN || iSwapStartTimeIndex < 0 || iSwapStartTimeIndex > iN - 1 ||
        iFreqRatio < 1 || iSwapTimePoints > iSwapVectorLength - 1)
        return false;

    for (int i = iFreqRatio; i <= iSwapTimePoints; i += iFreqRatio) {              // HSB:134-140
        if (i != iSwapTimePoints) P.pay[i] = exp(dStrikeCont * s.dPaymentInterval) - 1;
        if (i == iSwapTimePoints) P.pay[i] = exp(dStrikeCont * s.dPaymentInterval);
    }
    // HJM_Yield_to_Forward (HSB:143): f(0) = y(0), f(i) = (i+1) y(i) - i y(i-1)
    P.fwd[0] = pdYield[0];
    for (int i = 1; i <= iN - 1; ++i) P.fwd[i] = (i + 1) * pdYield[i] - i * pdYield[i - 1];
    // HJM_Drifts (HSB:148): per-factor no-arbitrage drifts, summed over the factors in factor order
    double drifts[MAXF][MAXN];
    for (int i = 0; i < iFactors; ++i) {
        const double *f = ppdFactors + (size_t)i * (iN - 1);
        drifts[i][0] = 0.5 * ddelt * (f[0]) * (f[0]);
        for (int j = 1; j <= iN - 2; ++j) {
            double d = 0;
            for (int l = 0; l <= j - 1; ++l) d -= drifts[i][l];
            double dSumVol = 0;
            for (int l = 0; l <= j; ++l) dSumVol += f[l];
            d += 0.5 * ddelt * (dSumVol) * (dSumVol);
            drifts[i][j] = d;
        }
        for (int l = 0; l <= iN - 2; ++l) P.fac[i][l] = f[l];
    }
    for (int l = 0; l <= iN - 2; ++l) {
        double tot = 0;
        for (int i = 0; i < iFactors; ++i) tot += drifts[i][l];
        P.driftdt[l] = tot * ddelt;  // the product inside HJM_SimPath_Forward_Blocking's path update
    }
    P.ddelt = ddelt;
    P.sqrt_ddelt = sqrt(ddelt);
    P.swap_ddelt = (double)(dSwapVectorYears / iSwapVectorLength);
    // Cumulative drift of the path entries the payoff reads (path_and_payoff keeps its rows drift-free): row j of the path is
    // row j-1 shifted by one maturity plus driftdt (HJM_SimPath_Forward_Blocking), so cd_j[l] = cd_{j-1}[l+1] + driftdt[l],
    // zero beyond the triangle like the path itself.
    {
        double cd[MAXN + 1] = {0};
        for (int j = 0; j <= iN - 1; ++j) {
            if (j <= iN - 2) P.xd_path[j] = -cd[0] * ddelt;                       // read by time step j + 1 (HSB:167-172)
            if (j == iSwapStartTimeIndex)
                for (int i = 0; i < iN; ++i) P.xd_swap[i] = -cd[i] * P.swap_ddelt;  // read by the swap leg (HSB:184)
            for (int l = 0; l <= iN - 2 - j; ++l) cd[l] = cd[l + 1] + P.driftdt[l];
            for (int l = iN - 1 - j < 0 ? 0 : iN - 1 - j; l < iN; ++l) cd[l] = 0;
        }
    }
    P.seed = seed;
    P.trials = lTrials;
    P.sims = lTrials <= 0 ? 0 : ((lTrials + BLOCKSIZE - 1) / BLOCKSIZE) * (long long)BLOCKSIZE;   // HSB:156
    P.start = iSwapStartTimeIndex;
    P.len = iSwapVectorLength;
    P.last_pay = iSwapTimePoints;
    return true;
}


// The tables of one swaption as the one-swaption kernel takes them (kernel-parameter constant bank).
inline void to_one_swaption(OneSwaption &P, const SwParams &H)
{
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < FN - 1; ++l)
        P.fd[l] = make_double4(H.fac[0][l] * H.sqrt_ddelt, H.fac[1][l] * H.sqrt_ddelt, H.fac[2][l] * H.sqrt_ddelt, H.driftdt[l]);
    for (int l = 0; l < FN; ++l) {
        P.fwd[l] = H.fwd[l];
        P.pay[l] = H.pay[l];
        P.xdp[l] = H.xd_path[l];
        P.xds[l] = H.xd_swap[l];
    }
    P.ddelt = H.ddelt;
    P.swap_ddelt = H.swap_ddelt;
    P.seed = H.seed;
    P.sims = H.sims;
    P.start = H.start;
    P.len = H.len;
    P.last_pay = H.last_pay;
}

}  // namespace swk
