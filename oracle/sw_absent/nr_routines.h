/* oracle/sw_absent/nr_routines.h -- TEST INFRASTRUCTURE ONLY; restated, see HJM_type.h.
 * Numerical-Recipes style allocators used at HJM_Securities.cpp:231,287,292,360-361 and
 * HJM_Swaption_Blocking.cpp:81-84,110-122. */
#ifndef SW_ABSENT_NR_ROUTINES_H
#define SW_ABSENT_NR_ROUTINES_H
#include "HJM_type.h"
FTYPE *dvector(long nl, long nh);
void free_dvector(FTYPE *v, long nl, long nh);
FTYPE **dmatrix(long nrl, long nrh, long ncl, long nch);
void free_dmatrix(FTYPE **m, long nrl, long nrh, long ncl, long nch);
#endif
