/* oracle/sw_absent/HJM_Securities.h -- TEST INFRASTRUCTURE ONLY; restated, see HJM_type.h.
 * Declares what HJM_Securities.cpp:314-319 calls and HJM_Swaption_Blocking.cpp:20-42 defines, plus dMax (:196). */
#ifndef SW_ABSENT_HJM_SECURITIES_H
#define SW_ABSENT_HJM_SECURITIES_H
#include "HJM_type.h"
int HJM_Swaption_Blocking(FTYPE *pdSwaptionPrice, FTYPE dStrike, FTYPE dCompounding, FTYPE dMaturity, FTYPE dTenor,
                          FTYPE dPaymentInterval, int iN, int iFactors, FTYPE dYears, FTYPE *pdYield,
                          FTYPE **ppdFactors, long iRndSeed, long lTrials, int BLOCKSIZE, int tid);
FTYPE dMax(FTYPE dA, FTYPE dB);
#endif
