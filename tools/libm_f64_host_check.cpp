// libm_f64_host_check.cpp -- TEST INFRASTRUCTURE.  Pins p3arsec_b200/csrc/bs_libm_f64.h to the libm of the box it runs
// on: exp_glibc / log_glibc (the header compiled for the host) against exp / log.
//   g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -mfma -I p3arsec_b200/csrc tools/libm_f64_host_check.cpp -o X -lm
//   X [millions of random arguments per class, default 20]
// Argument classes (a 64-bit space cannot be enumerated, so every branch of both functions gets its own dense sample):
//   exp: uniform bit patterns; x uniform in [-746, 710] (all of the finite-result range, incl. 512 <= |x|: specialcase, and
//        subnormal results); x in [-40, 40]; tiny |x| around 2^-54; every value i/64 +- few ulps for i in [-47000, 46000]
//   log: uniform bit patterns; positive normal numbers with uniform exponent; x in [1 - 2^-4, 1 + 0x1.09p-4) (the near-1
//        branch) and around its two edges; subnormals; x in (0, 4]
// plus zeros, infinities, NaNs.  Prints "exp <mismatches> <checked>" and "log <mismatches> <checked>"; NaN == NaN.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "bs_libm_f64.h"

static inline uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t &s) { return (double)(splitmix(s) >> 11) * 0x1p-53; }
static inline bool same(double a, double b) { return (a != a && b != b) || bsl64::d2u(a) == bsl64::d2u(b); }

int main(int argc, char **argv)
{
    const long M = (argc > 1 ? atol(argv[1]) : 20) * 1000000L;
    long bad_e = 0, bad_l = 0, n_e = 0, n_l = 0;
    double first_e = 0, first_l = 0;
#pragma omp parallel reduction(+ : bad_e, bad_l, n_e, n_l)
    {
        uint64_t seed = 0x1234567ull;
#ifdef _OPENMP
        seed += 0x9e37ull * (uint64_t)omp_get_thread_num();
#endif
#pragma omp for schedule(static)
        for (long i = 0; i < M; i++) {
            double xs[8];
            xs[0] = bsl64::u2d(splitmix(seed));                                  // any bit pattern
            xs[1] = -746.0 + 1456.0 * u01(seed);                                 // the whole finite range of exp
            xs[2] = -40.0 + 80.0 * u01(seed);
            xs[3] = ldexp(u01(seed) - 0.5, -52 + (int)(splitmix(seed) % 6));     // around 2^-54
            xs[4] = (double)((long)(splitmix(seed) % 93000) - 47000) / 64.0;     // near table breakpoints
            xs[4] = bsl64::u2d(bsl64::d2u(xs[4]) + (splitmix(seed) % 9) - 4);
            xs[5] = -745.2 + 37.5 * u01(seed);                                   // subnormal results
            xs[6] = 700.0 + 10.0 * u01(seed);                                    // next to overflow
            xs[7] = -(512.0 + 233.0 * u01(seed));
            for (int j = 0; j < 8; j++) {
                n_e++;
                if (!same(exp(xs[j]), bsl64::exp_glibc(xs[j]))) { bad_e++; first_e = xs[j]; }
            }
            double ls[7];
            ls[0] = bsl64::u2d(splitmix(seed));
            ls[1] = ldexp(1.0 + u01(seed), (int)(splitmix(seed) % 2046) - 1022); // every binade
            ls[2] = 0.9375 + (1.064697265625 - 0.9375) * u01(seed);              // near-1 branch
            ls[3] = bsl64::u2d((splitmix(seed) & 1 ? 0x3fee000000000000ull : 0x3ff1090000000000ull) + (splitmix(seed) % 2001) - 1000);  // its edges
            ls[4] = bsl64::u2d(splitmix(seed) & 0x000fffffffffffffull);          // subnormals
            ls[5] = 4.0 * u01(seed);
            ls[6] = bsl64::u2d(0x3ff0000000000000ull + (splitmix(seed) % 200001) - 100000);  // 1 +- a few ulps
            for (int j = 0; j < 7; j++) {
                n_l++;
                if (!same(log(ls[j]), bsl64::log_glibc(ls[j]))) { bad_l++; first_l = ls[j]; }
            }
        }
    }
    const uint64_t specials[] = {0x0ull, 0x8000000000000000ull, 0x7ff0000000000000ull, 0xfff0000000000000ull, 0x7ff8000000000000ull,
                                 0xfff8000000000000ull, 0x1ull, 0x8000000000000001ull, 0x3ff0000000000000ull, 0xbff0000000000000ull,
                                 0x4080000000000000ull, 0xc080000000000000ull, 0x4090000000000000ull, 0xc090000000000000ull,
                                 0x40862e42fefa39efull, 0x40862e42fefa39f0ull, 0xc0874910d52d3051ull, 0xc0874910d52d3052ull, 0xc0874385446d71c3ull,
                                 0xc0874385446d71c4ull, 0x3c90000000000000ull, 0x3c8fffffffffffffull, 0x0010000000000000ull, 0x000fffffffffffffull,
                                 0x7fefffffffffffffull};
    for (uint64_t u : specials) {
        const double x = bsl64::u2d(u);
        n_e++;
        n_l++;
        if (!same(exp(x), bsl64::exp_glibc(x))) { bad_e++; first_e = x; }
        if (!same(log(x), bsl64::log_glibc(x))) { bad_l++; first_l = x; }
    }
    printf("exp %ld %ld %a\nlog %ld %ld %a\n", bad_e, n_e, first_e, bad_l, n_l, first_l);
    return (bad_e || bad_l) ? 1 : 0;
}
