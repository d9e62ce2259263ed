/* oracle/sw_absent/HJM.h -- TEST INFRASTRUCTURE ONLY; restated, see HJM_type.h.
 * Call sites: HJM_Swaption_Blocking.cpp:143 (HJM_Yield_to_Forward), :148 (HJM_Drifts),
 * :161 (HJM_SimPath_Forward_Blocking), :172,184 (Discount_Factors_Blocking);
 * HJM_Securities.cpp:198,279,281 (RanUnif). */
#ifndef SW_ABSENT_HJM_H
#define SW_ABSENT_HJM_H
#include "HJM_type.h"
FTYPE RanUnif(long *s);
FTYPE CumNormalInv(FTYPE u);
int HJM_Yield_to_Forward(FTYPE *pdForward, int iN, FTYPE *pdYield);
int HJM_Drifts(FTYPE *pdTotalDrift, FTYPE **ppdDrifts, int iN, int iFactors, FTYPE dYears, FTYPE **ppdFactors);
int HJM_SimPath_Forward_Blocking(FTYPE **ppdHJMPath, int iN, int iFactors, FTYPE dYears, FTYPE *pdForward,
                                 FTYPE *pdTotalDrift, FTYPE **ppdFactors, long *lRndSeed, int BLOCKSIZE);
int Discount_Factors_Blocking(FTYPE *pdDiscountFactors, int iN, FTYPE dYears, FTYPE *pdRatePath, int BLOCKSIZE);
#endif
