// Host-side accuracy check of p3arsec_b200/csrc/sw_tail.h (the composite table of Moro's tail branch).
//   g++ -O2 -std=c++17 -o sw_tail_host_check tools/sw_tail_host_check.cpp -lm
// Draws the 31-bit integers a tail draw can have, forms r = min(u, 1 - u) exactly as the kernel does and compares
// moro_tail(r) with P8(log(-log r)) evaluated in long double.  Prints "max_rel <value>" and "max_ulp <value>".
#include <cmath>
#include <cstdio>
#include <random>

#include "../p3arsec_b200/csrc/sw_tail.h"

static const long double C[9] = {0.3374754822726147, 0.9761690190917186, 0.1607979714918209, 0.0276438810333863, 0.0038405729373609,
                                 0.0003951896511919, 0.0000321767881768, 0.0000002888167364, 0.0000003960315187};

int main()
{
    static double logtab[bsm::TAB_DOUBLES];
    alignas(16) static double tt[swt::TAIL_DOUBLES];
    bsm::fill_tables(logtab, 0, 1);
    swt::fill_tail(tt, 0, 1);
    const unsigned S_LO = 171798692u, S_HI = 1975684955u, M = 2147483647u;
    std::mt19937_64 gen(20261017);
    double max_rel = 0, max_ulp = 0;
    auto check = [&](unsigned s) {
        const double u = (double)(int)s * 4.656612875e-10;
        const double r = s > S_HI ? 1.0 - u : u;
        const double got = swt::moro_tail(r, logtab, tt);
        const long double w = logl(-logl((long double)r));
        long double p = C[8];
        for (int k = 7; k >= 0; --k) p = C[k] + w * p;
        const double rel = (double)(fabsl(got - p) / fabsl(p));
        int e;
        frexpl(p, &e);
        const double ulp = (double)(fabsl(got - p) / ldexpl(1.0L, e - 53));
        if (rel > max_rel) max_rel = rel;
        if (ulp > max_ulp) max_ulp = ulp;
    };
    for (unsigned s : {1u, 2u, 3u, 100u, 65535u, 65536u, S_LO - 1, S_HI + 1, M - 1}) check(s);
    for (int i = 0; i < 2000000; ++i) {
        check(1 + (unsigned)(gen() % (S_LO - 1)));              // lower tail
        check(S_HI + 1 + (unsigned)(gen() % (M - 1 - S_HI)));   // upper tail
    }
    for (unsigned s = 1; s < 200000; ++s) check(s);               // the extreme tail, exhaustively
    printf("max_rel %.3e\nmax_ulp %.2f\n", max_rel, max_ulp);
    return 0;
}
