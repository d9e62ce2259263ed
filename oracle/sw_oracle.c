/* oracle/sw_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker for the swaptions Map).
 *
 * Plain-C restatement of the swaptions hot path of P3ARSEC:
 *   - the per-swaption Map body HJM_Swaption_Blocking()   parsec-ff/pkgs/apps/swaptions/src/HJM_Swaption_Blocking.cpp:20-222
 *   - the portfolio set-up of the driver                   parsec-ff/pkgs/apps/swaptions/src/HJM_Securities.cpp:198,231-297
 *   - the Map over swaptions (FastFlow / SkePU / pthreads) HJM_Securities.cpp:311-323, HJM_Securities_skepu.cpp:43-57
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of the bench tools may
 * load it; nothing under p3arsec_b200/ does.
 *
 * Formulation: one trial at a time (the reference simulates BLOCKSIZE = 16 trials side by side, which is a
 * cache blocking and changes no arithmetic), flat arrays instead of dmatrix/dvector, sums accumulated in trial
 * order -- so the result is bit-identical to the reference's.  The leaf routines come from
 * oracle/sw_absent/sw_leaves.c (RanUnif, CumNormalInv; exp/log/sqrt from libm).
 *
 * PINNING: HJM_Swaption_Blocking.cpp and HJM_Securities.cpp exist in /root/reference and are compiled
 * unmodified into oracle/_ref/sw_ref_{serial,ff} (oracle/Makefile); tests/test_sw_oracle.py requires this
 * restatement to reproduce their "%.10lf" output byte for byte (committed goldens under tests/golden/sw_*).
 * The leaves under oracle/sw_absent/ are PARSEC-owned files ABSENT from /root/reference: their bodies are
 * "parity unpinned" (restated from the published PARSEC 3.0 package, nothing here to diff against), and that
 * caveat is inherited by every swaptions number in this repo.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#include "sw_absent/HJM.h"

#define SW_MAX_N 64

typedef struct sw_oracle_swaption {   /* the scalar parm fields the Map passes (HJM_Securities.cpp:314-318) */
    double dStrike, dCompounding, dMaturity, dTenor, dPaymentInterval, dYears;
} sw_oracle_swaption;

/* HJM_Swaption_Blocking.cpp:20-222.  factors: iFactors x (iN-1), row-major.  Returns 1 like the reference. */
int sw_oracle_price_one(double out[2], const sw_oracle_swaption *p, int iN, int iFactors, const double *pdYield,
                        const double *factors, long iRndSeed, long lTrials, int BLOCKSIZE)
{
    if (iN < 2 || iN > SW_MAX_N || iFactors < 1 || iFactors > SW_MAX_N || BLOCKSIZE < 1) return 0;
    double ddelt = (double)(p->dYears / iN);                                   /* :48 */
    int iFreqRatio = (int)(p->dPaymentInterval / ddelt + 0.5);                 /* :50 */
    double dStrikeCont;
    if (p->dCompounding == 0) dStrikeCont = p->dStrike;                        /* :56-57 */
    else dStrikeCont = (1 / p->dCompounding) * log(1 + p->dStrike * p->dCompounding);   /* :61 */

    int iSwapVectorLength = (int)(iN - p->dMaturity / ddelt + 0.5);            /* :116 */
    int iSwapStartTimeIndex = (int)(p->dMaturity / ddelt + 0.5);               /* :125 */
    int iSwapTimePoints = (int)(p->dTenor / ddelt + 0.5);                      /* :126 */
    double dSwapVectorYears = (double)(iSwapVectorLength * ddelt);             /* :127 */
    /* the reference indexes its heap vectors with these without checks; refuse what would run off them */
    if (iSwapVectorLength < 1 || iSwapVectorLength > iN || iSwapStartTimeIndex < 0 || iSwapStartTimeIndex > iN - 1 ||
        iFreqRatio < 1 || iSwapTimePoints > iSwapVectorLength - 1)
        return 0;

    double pdSwapPayoffs[SW_MAX_N];
    for (int i = 0; i <= iSwapVectorLength - 1; ++i) pdSwapPayoffs[i] = 0.0;    /* :132-133 */
    for (int i = iFreqRatio; i <= iSwapTimePoints; i += iFreqRatio) {           /* :134-140 */
        if (i != iSwapTimePoints) pdSwapPayoffs[i] = exp(dStrikeCont * p->dPaymentInterval) - 1;
        if (i == iSwapTimePoints) pdSwapPayoffs[i] = exp(dStrikeCont * p->dPaymentInterval);
    }

    /* :143 HJM_Yield_to_Forward, :148 HJM_Drifts -- through the leaves (they take NR matrices) */
    double pdForward[SW_MAX_N], pdTotalDrift[SW_MAX_N];
    double *drift_store = (double *)malloc(sizeof(double) * SW_MAX_N * SW_MAX_N);
    double *ppdDrifts[SW_MAX_N], *ppdFactors[SW_MAX_N];
    for (int i = 0; i < iFactors; ++i) {
        ppdDrifts[i] = drift_store + (size_t)i * SW_MAX_N;
        ppdFactors[i] = (double *)factors + (size_t)i * (iN - 1);
    }
    HJM_Yield_to_Forward(pdForward, iN, (double *)pdYield);
    HJM_Drifts(pdTotalDrift, ppdDrifts, iN, iFactors, p->dYears, ppdFactors);
    free(drift_store);

    double sqrt_ddelt = sqrt(ddelt);
    double dSwapDelt = (double)(dSwapVectorYears / iSwapVectorLength);   /* Discount_Factors_Blocking's own ddelt at :184 */
    double dSum = 0.0, dSumSquare = 0.0;                                   /* :152-153 */
    long draws_per_trial = (long)(iN - 1) * iFactors;

    double(*path)[SW_MAX_N] = (double(*)[SW_MAX_N])malloc(sizeof(double) * SW_MAX_N * SW_MAX_N);
    double(*z)[SW_MAX_N] = (double(*)[SW_MAX_N])malloc(sizeof(double) * SW_MAX_N * SW_MAX_N);   /* z[factor][step] */

    /* :156  for (l = 0; l <= lTrials-1; l += BLOCKSIZE): every block simulates BLOCKSIZE trials, also the last
     * one when lTrials is not a multiple of BLOCKSIZE (the sums then hold more than lTrials terms). */
    long nSim = ((lTrials + BLOCKSIZE - 1) / BLOCKSIZE) * BLOCKSIZE;
    if (lTrials <= 0) nSim = 0;
    for (long t = 0; t < nSim; ++t) {
        long ctr = iRndSeed + t * draws_per_trial;
        /* HJM_SimPath_Forward_Blocking: row 0 = forward curve, the rest zero */
        for (int j = 0; j <= iN - 1; ++j) {
            path[0][j] = pdForward[j];
            for (int i = 1; i <= iN - 1; ++i) path[i][j] = 0;
        }
        for (int j = 1; j <= iN - 1; ++j)
            for (int l = 0; l <= iFactors - 1; ++l)
                z[l][j] = CumNormalInv(RanUnif(&ctr));
        for (int j = 1; j <= iN - 1; ++j)
            for (int l = 0; l <= iN - (j + 1); ++l) {
                double dTotalShock = 0;
                for (int i = 0; i <= iFactors - 1; ++i) dTotalShock += factors[(size_t)i * (iN - 1) + l] * z[i][j];
                path[j][l] = path[j - 1][l + 1] + pdTotalDrift[l] * ddelt + sqrt_ddelt * dTotalShock;
            }
        /* :167-172 discount factors along column 0 of the path; only index iSwapStartTimeIndex is read (:198) */
        double dPayoffDF = 1.0;
        for (int j = 0; j <= iSwapStartTimeIndex - 1; ++j) dPayoffDF *= exp(-path[j][0] * ddelt);
        /* :179-184 discount factors along the swap row; :193-195 fixed leg */
        double dFixedLegValue = 0.0;
        double df = 1.0;
        for (int i = 0; i <= iSwapVectorLength - 1; ++i) {
            if (i >= 1) df *= exp(-path[iSwapStartTimeIndex][i - 1] * dSwapDelt);
            dFixedLegValue += pdSwapPayoffs[i] * df;
        }
        double dSwaptionPayoff = (dFixedLegValue - 1.0 > 0) ? dFixedLegValue - 1.0 : 0;   /* :196 dMax */
        double dDisc = dSwaptionPayoff * dPayoffDF;                                       /* :198 */
        dSum += dDisc;                                                                    /* :203 */
        dSumSquare += dDisc * dDisc;                                                      /* :204 */
    }
    free(path);
    free(z);
    out[0] = dSum / lTrials;                                                              /* :212 */
    out[1] = sqrt((dSumSquare - dSum * dSum / lTrials) / (lTrials - 1.0)) / sqrt((double)lTrials);   /* :213-214 */
    return 1;
}

/* The Map over swaptions: HJM_Securities.cpp:312-323 (swaption i gets seed swaption_seed + i). */
int sw_oracle_map(int nSwaptions, const sw_oracle_swaption *s, int iN, int iFactors, const double *yields,
                  const double *factors, long swaption_seed, long lTrials, int BLOCKSIZE, double *mean, double *err,
                  int nthreads)
{
    int ok = 1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = 0; i < nSwaptions; ++i) {
        double out[2] = {0, 0};
        int r = sw_oracle_price_one(out, &s[i], iN, iFactors, yields + (size_t)i * iN,
                                    factors + (size_t)i * iFactors * (iN - 1), swaption_seed + i, lTrials, BLOCKSIZE);
        if (r != 1) {
#pragma omp atomic write
            ok = 0;
        }
        mean[i] = out[0];
        err[i] = out[1];
    }
    return ok;
}

/* The driver's portfolio: HJM_Securities.cpp:198 (swaption_seed), :231-265 (factor table), :276-296. */
static const double sw_factor_table[3][10] = {
    {.01, .01, .01, .01, .01, .01, .01, .01, .01, .01},
    {.009048, .008187, .007408, .006703, .006065, .005488, .004966, .004493, .004066, .003679},
    {.001000, .000750, .000500, .000250, .000000, -.000250, -.000500, -.000750, -.001000, -.001250}};

long sw_oracle_portfolio(int nSwaptions, long seed, sw_oracle_swaption *s, double *yields /* n x 11 */,
                         double *factors /* n x 3 x 10 */)
{
    const int iN = 11, iFactors = 3;
    long swaption_seed = (long)(2147483647L * RanUnif(&seed));                 /* :198 */
    for (int i = 0; i < nSwaptions; i++) {
        s[i].dYears = 5.0 + ((int)(60 * RanUnif(&seed))) * 0.25;               /* :279 */
        s[i].dStrike = 0.1 + ((int)(49 * RanUnif(&seed))) * 0.1;               /* :281 */
        s[i].dCompounding = 0;
        s[i].dMaturity = 1.0;
        s[i].dTenor = 2.0;
        s[i].dPaymentInterval = 1.0;
        double *y = yields + (size_t)i * iN;
        y[0] = .1;                                                             /* :288 */
        for (int j = 1; j <= iN - 1; ++j) y[j] = y[j - 1] + .005;              /* :289-290 */
        for (int k = 0; k < iFactors; ++k)
            for (int j = 0; j <= iN - 2; ++j) factors[((size_t)i * iFactors + k) * (iN - 1) + j] = sw_factor_table[k][j];
    }
    return swaption_seed;
}

double sw_oracle_ranunif(long *s) { return RanUnif(s); }
double sw_oracle_cumnormalinv(double u) { return CumNormalInv(u); }
