/* oracle/sw_absent/sw_leaves.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the PARSEC-owned leaf routines of the swaptions package that P3ARSEC's overlay does NOT
 * ship (see HJM_type.h in this directory for the full story): RanUnif.c, CumNormalInv.c, MaxFunction.c,
 * nr_routines.c, HJM.cpp and HJM_SimPath_Forward_Blocking.cpp of PARSEC 3.0 pkgs/apps/swaptions/src.
 * Written from the published algorithm (Park-Miller minimal standard generator driven by a counter;
 * Moro's inverse normal; the HJM forward-rate recursion of Broadie/Dewanwala) -- PARITY UNPINNED for the
 * bodies in this file.  Compiles as C (into libsw_oracle.so) and as C++ (into oracle/_ref/sw_ref_*, linked
 * with the reference's own HJM_Securities.cpp and HJM_Swaption_Blocking.cpp).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "HJM.h"
#include "HJM_Securities.h"
#include "nr_routines.h"

/* ---- RanUnif: counter-driven Park-Miller draw.  The state is a plain counter (advanced by one per
 * call); the draw is one minimal-standard step (a = 16807, m = 2^31-1, Schrage split q = 127773,
 * r = 2836) applied to counter * 1513517 mod m, scaled by 1/m.  Because the k-th draw depends only on
 * seed + k, any trial's normals can be generated independently -- this is what the GPU kernel uses. ---- */
FTYPE RanUnif(long *s)
{
    long ix, k1;
    FTYPE dRes;

    ix = *s;
    *s = ix + 1;
    ix *= 1513517L;
    ix %= 2147483647L;
    k1 = ix / 127773L;
    ix = 16807L * (ix - k1 * 127773L) - k1 * 2836L;
    if (ix < 0) ix = ix + 2147483647L;
    dRes = (ix * 4.656612875e-10);
    return dRes;
}

/* ---- CumNormalInv: Moro (1995) inverse cumulative normal: rational approximation for |u-0.5| < 0.42,
 * degree-8 polynomial in log(-log(tail)) outside. ---- */
static const FTYPE sw_a[4] = {2.50662823884, -18.61500062529, 41.39119773534, -25.44106049637};
static const FTYPE sw_b[4] = {-8.47351093090, 23.08336743743, -21.06224101826, 3.13082909833};
static const FTYPE sw_c[9] = {0.3374754822726147, 0.9761690190917186, 0.1607979714918209,
                              0.0276438810333863, 0.0038405729373609, 0.0003951896511919,
                              0.0000321767881768, 0.0000002888167364, 0.0000003960315187};

FTYPE CumNormalInv(FTYPE u)
{
    FTYPE x, r;

    x = u - 0.5;
    if (fabs(x) < 0.42) {
        r = x * x;
        r = x * (((sw_a[3] * r + sw_a[2]) * r + sw_a[1]) * r + sw_a[0]) /
            ((((sw_b[3] * r + sw_b[2]) * r + sw_b[1]) * r + sw_b[0]) * r + 1.0);
        return r;
    }
    r = u;
    if (x > 0.0) r = 1.0 - u;
    r = log(-log(r));
    r = sw_c[0] + r * (sw_c[1] + r * (sw_c[2] + r * (sw_c[3] + r * (sw_c[4] + r * (sw_c[5] +
        r * (sw_c[6] + r * (sw_c[7] + r * sw_c[8])))))));
    if (x < 0.0) r = -r;
    return r;
}

FTYPE dMax(FTYPE dA, FTYPE dB)
{
    return (dA > dB ? dA : dB);
}

/* ---- Numerical-Recipes allocators (index ranges [nl..nh]; one contiguous block per matrix) ---- */
FTYPE *dvector(long nl, long nh)
{
    FTYPE *v = (FTYPE *)malloc((size_t)((nh - nl + 2) * sizeof(FTYPE)));
    if (!v) { fprintf(stderr, "allocation failure in dvector()\n"); exit(1); }
    return v - nl + 1;
}

void free_dvector(FTYPE *v, long nl, long nh)
{
    (void)nh;
    free((char *)(v + nl - 1));
}

FTYPE **dmatrix(long nrl, long nrh, long ncl, long nch)
{
    long i, nrow = nrh - nrl + 1, ncol = nch - ncl + 1;
    FTYPE **m = (FTYPE **)malloc((size_t)((nrow + 1) * sizeof(FTYPE *)));
    if (!m) { fprintf(stderr, "allocation failure 1 in dmatrix()\n"); exit(1); }
    m += 1;
    m -= nrl;
    m[nrl] = (FTYPE *)malloc((size_t)((nrow * ncol + 1) * sizeof(FTYPE)));
    if (!m[nrl]) { fprintf(stderr, "allocation failure 2 in dmatrix()\n"); exit(1); }
    m[nrl] += 1;
    m[nrl] -= ncl;
    for (i = nrl + 1; i <= nrh; i++) m[i] = m[i - 1] + ncol;
    return m;
}

void free_dmatrix(FTYPE **m, long nrl, long nrh, long ncl, long nch)
{
    (void)nrh; (void)nch;
    free((char *)(m[nrl] + ncl - 1));
    free((char *)(m + nrl - 1));
}

/* ---- HJM.cpp: forward curve from yields, f(0) = y(0), f(i) = (i+1) y(i) - i y(i-1) ---- */
int HJM_Yield_to_Forward(FTYPE *pdForward, int iN, FTYPE *pdYield)
{
    int i;
    pdForward[0] = pdYield[0];
    for (i = 1; i <= iN - 1; ++i)
        pdForward[i] = (i + 1) * pdYield[i] - i * pdYield[i - 1];
    return 1;
}

/* ---- HJM.cpp: no-arbitrage drift per factor and maturity, then summed over factors ---- */
int HJM_Drifts(FTYPE *pdTotalDrift, FTYPE **ppdDrifts, int iN, int iFactors, FTYPE dYears, FTYPE **ppdFactors)
{
    int i, j, l;
    FTYPE ddelt = (FTYPE)(dYears / iN);
    FTYPE dSumVol;

    for (i = 0; i <= iFactors - 1; ++i)
        ppdDrifts[i][0] = 0.5 * ddelt * (ppdFactors[i][0]) * (ppdFactors[i][0]);

    for (i = 0; i <= iFactors - 1; ++i)
        for (j = 1; j <= iN - 2; ++j) {
            ppdDrifts[i][j] = 0;
            for (l = 0; l <= j - 1; ++l)
                ppdDrifts[i][j] -= ppdDrifts[i][l];
            dSumVol = 0;
            for (l = 0; l <= j; ++l)
                dSumVol += ppdFactors[i][l];
            ppdDrifts[i][j] += 0.5 * ddelt * (dSumVol) * (dSumVol);
        }

    for (i = 0; i <= iN - 2; ++i) {
        pdTotalDrift[i] = 0;
        for (j = 0; j <= iFactors - 1; ++j)
            pdTotalDrift[i] += ppdDrifts[j][i];
    }
    return 1;
}

/* ---- HJM.cpp: discount factors along a rate path, BLOCKSIZE trials side by side:
 * DF[i] = prod_{j<i} exp(-rate[j] * ddelt), every exponential rounded on its own ---- */
int Discount_Factors_Blocking(FTYPE *pdDiscountFactors, int iN, FTYPE dYears, FTYPE *pdRatePath, int BLOCKSIZE)
{
    int i, j, b;
    FTYPE ddelt = (FTYPE)(dYears / iN);
    FTYPE *pdexpRes = dvector(0, (iN - 1) * BLOCKSIZE - 1);

    for (j = 0; j <= (iN - 1) * BLOCKSIZE - 1; ++j) pdexpRes[j] = -pdRatePath[j] * ddelt;
    for (j = 0; j <= (iN - 1) * BLOCKSIZE - 1; ++j) pdexpRes[j] = exp(pdexpRes[j]);

    for (i = 0; i < iN * BLOCKSIZE; ++i) pdDiscountFactors[i] = 1.0;

    for (i = 1; i <= iN - 1; ++i)
        for (b = 0; b < BLOCKSIZE; b++)
            for (j = 0; j <= i - 1; ++j)
                pdDiscountFactors[i * BLOCKSIZE + b] *= pdexpRes[j * BLOCKSIZE + b];

    free_dvector(pdexpRes, 0, (iN - 1) * BLOCKSIZE - 1);
    return 1;
}

/* ---- HJM_SimPath_Forward_Blocking.cpp: BLOCKSIZE forward-rate paths.
 * Draw order: trial b, then time step j = 1..iN-1, then factor l -- so trial t of a swaption consumes
 * draws [t*(iN-1)*iFactors, (t+1)*(iN-1)*iFactors) of that swaption's counter stream. ---- */
static void sw_serialB(FTYPE **pdZ, FTYPE **randZ, int BLOCKSIZE, int iN, int iFactors)
{
    int l, b, j;
    for (l = 0; l <= iFactors - 1; ++l)
        for (b = 0; b < BLOCKSIZE; b++)
            for (j = 1; j <= iN - 1; ++j)
                pdZ[l][BLOCKSIZE * j + b] = CumNormalInv(randZ[l][BLOCKSIZE * j + b]);
}

int HJM_SimPath_Forward_Blocking(FTYPE **ppdHJMPath, int iN, int iFactors, FTYPE dYears, FTYPE *pdForward,
                                 FTYPE *pdTotalDrift, FTYPE **ppdFactors, long *lRndSeed, int BLOCKSIZE)
{
    int i, j, l, b;
    FTYPE **pdZ, **randZ;
    FTYPE dTotalShock, ddelt, sqrt_ddelt;

    ddelt = (FTYPE)(dYears / iN);
    sqrt_ddelt = sqrt(ddelt);

    pdZ = dmatrix(0, iFactors - 1, 0, iN * BLOCKSIZE - 1);
    randZ = dmatrix(0, iFactors - 1, 0, iN * BLOCKSIZE - 1);

    for (b = 0; b < BLOCKSIZE; b++)
        for (j = 0; j <= iN - 1; j++) {
            ppdHJMPath[0][BLOCKSIZE * j + b] = pdForward[j];
            for (i = 1; i <= iN - 1; ++i) ppdHJMPath[i][BLOCKSIZE * j + b] = 0;
        }

    for (b = 0; b < BLOCKSIZE; b++)
        for (j = 1; j <= iN - 1; ++j)
            for (l = 0; l <= iFactors - 1; ++l)
                randZ[l][BLOCKSIZE * j + b] = RanUnif(lRndSeed);

    sw_serialB(pdZ, randZ, BLOCKSIZE, iN, iFactors);

    for (b = 0; b < BLOCKSIZE; b++)
        for (j = 1; j <= iN - 1; ++j)
            for (l = 0; l <= iN - (j + 1); ++l) {
                dTotalShock = 0;
                for (i = 0; i <= iFactors - 1; ++i)
                    dTotalShock += ppdFactors[i][l] * pdZ[i][BLOCKSIZE * j + b];
                ppdHJMPath[j][BLOCKSIZE * l + b] =
                    ppdHJMPath[j - 1][BLOCKSIZE * (l + 1) + b] + pdTotalDrift[l] * ddelt + sqrt_ddelt * dTotalShock;
            }

    free_dmatrix(pdZ, 0, iFactors - 1, 0, iN * BLOCKSIZE - 1);
    free_dmatrix(randZ, 0, iFactors - 1, 0, iN * BLOCKSIZE - 1);
    return 1;
}
