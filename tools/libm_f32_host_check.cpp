// libm_f32_host_check.cpp -- TEST INFRASTRUCTURE.  Pins p3arsec_b200/csrc/bs_libm_f32.h to the libm of the box it
// runs on: expf_glibc / logf_glibc (the header compiled for the host) against expf / logf over floats.
//   g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -mfma -I p3arsec_b200/csrc tools/libm_f32_host_check.cpp -o X -lm
//   X [stride]      stride 1 = every one of the 2^32 bit patterns (about 20 s on 8 cores); default 64
// Prints "expf <mismatches> <checked>" and "logf <mismatches> <checked>"; NaN results count as equal to NaN.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "bs_libm_f32.h"

int main(int argc, char **argv)
{
    const long stride = argc > 1 ? atol(argv[1]) : 64;
    long bad_e = 0, bad_l = 0, n = 0;
    uint32_t first_e = 0, first_l = 0;
#pragma omp parallel for reduction(+ : bad_e, bad_l, n) schedule(static)
    for (long u = 0; u < 0x100000000L; u += stride) {
        const float x = bsl::u2f((uint32_t)u);
        const float e0 = expf(x), e1 = bsl::expf_glibc(x);
        const float l0 = logf(x), l1 = bsl::logf_glibc(x);
        n++;
        if (!(e0 != e0 && e1 != e1) && bsl::f2u(e0) != bsl::f2u(e1)) { bad_e++; first_e = (uint32_t)u; }
        if (!(l0 != l0 && l1 != l1) && bsl::f2u(l0) != bsl::f2u(l1)) { bad_l++; first_l = (uint32_t)u; }
    }
    printf("expf %ld %ld %08x\nlogf %ld %ld %08x\n", bad_e, n, first_e, bad_l, n, first_l);
    return (bad_e || bad_l) ? 1 : 0;
}
