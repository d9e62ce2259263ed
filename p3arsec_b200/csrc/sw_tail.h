/*
 * sw_tail.h -- |z| of Moro's tail branch of CumNormalInv (PARSEC CumNormalInv.c, called through
 * HJM_SimPath_Forward_Blocking from parsec-ff/pkgs/apps/swaptions/src/HJM_Swaption_Blocking.cpp:161) for the fast
 * swaptions kernel:  z = +/- P8(log(-log r)),  r = min(u, 1 - u) in [4.66e-10, 0.08].
 *
 * y = -log r comes from the table-driven log_f64 of bs_math_f64.h (12 FP64 operations); P8(log y) comes from the
 * composite table of sw_tail_table.h -- one degree-8 polynomial in y - centre per interval of y (64 intervals:
 * exponent x top four mantissa bits), 9 FP64 operations instead of the 20 of a second logarithm plus P8.
 * Host-or-device code like bs_math_f64.h: tools/sw_tail_host_check.cpp measures it against long-double libm on the CPU.
 */
#pragma once
#include "bs_math_f64.h"
#include "sw_tail_table.h"

namespace swt {

enum { TAIL_DOUBLES = TAIL_ROWS * TAIL_COLS };

#ifndef SW_TAIL_LAYOUT
#define SW_TAIL_LAYOUT 1  /* 0: row-major {centre, a0..a8} per interval (LDS.128); 1: coefficient-major (LDS.64, fewer bank conflicts) */
#endif

BS_HD void fill_tail(double *t, int first, int step)
{
    for (int i = first; i < TAIL_DOUBLES; i += step) {
#if SW_TAIL_LAYOUT == 0
        t[i] = bsm::from_bits(TAIL_BITS[i / TAIL_COLS][i % TAIL_COLS]);
#else
        t[i] = bsm::from_bits(TAIL_BITS[i % TAIL_ROWS][i / TAIL_ROWS]);
#endif
    }
}

// lt: where log_f64_t finds its table (bsm::PlainLogTab or a bank-replicated copy); tt: the block filled by fill_tail
// (16-byte aligned).
template <class LT>
BS_HD double moro_tail_t(double r, const LT &lt, const double *tt)
{
    const double y = -bsm::log_f64_t(r, lt);                            // in [2.52, 21.5]
    const int idx = ((int)(bsm::to_bits(y) >> 48) - 0x4000) & 63;       // 16 (exponent - 1) + top four mantissa bits
#if SW_TAIL_LAYOUT == 0
    const double *a = tt + idx * TAIL_COLS;
#define SW_A(k) a[k]
#else
    const double *a = tt + idx;
#define SW_A(k) a[(k) * TAIL_ROWS]
#endif
    const double d = y - SW_A(0);                                       // exact: same binade
    double p = fma(SW_A(9), d, SW_A(8));
    p = fma(p, d, SW_A(7));
    p = fma(p, d, SW_A(6));
    p = fma(p, d, SW_A(5));
    p = fma(p, d, SW_A(4));
    p = fma(p, d, SW_A(3));
    p = fma(p, d, SW_A(2));
    return fma(p, d, SW_A(1));
#undef SW_A
}
BS_HD double moro_tail(double r, const double *logtab, const double *tt)
{
    const bsm::PlainLogTab lt = {logtab};
    return moro_tail_t(r, lt, tt);
}

}  // namespace swt
