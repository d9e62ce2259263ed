/* oracle/sw_absent/HJM_type.h -- TEST INFRASTRUCTURE ONLY.
 *
 * P3ARSEC is an overlay on the PARSEC 3.0 tarball (install.sh:59-73): for swaptions it ships only
 * HJM_Securities*.cpp, HJM_Swaption_Blocking.cpp and the Makefile.  The headers and leaf routines those
 * files include/call (HJM_type.h, HJM.h, HJM_Securities.h, nr_routines.h, RanUnif, CumNormalInv, dMax,
 * HJM_SimPath_Forward_Blocking, HJM_Yield_to_Forward, HJM_Drifts, Discount_Factors_Blocking, dvector/
 * dmatrix) are PARSEC-owned and ABSENT from /root/reference.  This directory restates them from the
 * published PARSEC 3.0 swaptions package (pkgs/apps/swaptions/src) so that the two files P3ARSEC does
 * ship can be compiled, unmodified, from where they lie (oracle/Makefile, target sw_ref_*).
 * PARITY STATUS: the names, signatures and call sites are pinned by the reference's own files (they
 * would not compile or link otherwise); the BODIES of the absent leaves are "parity unpinned" -- there
 * is no copy of PARSEC here to diff them against.
 *
 * Used by the reference at: parm fields HJM_Securities.cpp:276-296,322-323; FTYPE throughout;
 * BLOCK_SIZE :319; DEFAULT_NUM_TRIALS :53.
 */
#ifndef SW_ABSENT_HJM_TYPE_H
#define SW_ABSENT_HJM_TYPE_H

#include <assert.h>
#include <string.h>

#define FTYPE double
#define BLOCK_SIZE 16          /* trials simulated per HJM_SimPath_Forward_Blocking call */
#define DEFAULT_NUM_TRIALS 102400

typedef struct {
    int Id;
    FTYPE dSimSwaptionMeanPrice;
    FTYPE dSimSwaptionStdError;
    FTYPE dStrike;
    FTYPE dCompounding;
    FTYPE dMaturity;
    FTYPE dTenor;
    FTYPE dPaymentInterval;
    int iN;
    FTYPE dYears;
    int iFactors;
    FTYPE *pdYield;
    FTYPE **ppdFactors;
} parm;

#endif
