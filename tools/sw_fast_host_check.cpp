// Host build of the fast swaptions kernels' per-trial arithmetic (p3arsec_b200/csrc/sw_kernels.cuh is host-or-device
// source: the same normals() / path_and_payoff() / exp_core / tail table the GPU runs, MUFU seeds emulated at 2^-20).
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -o sw_fast_host_check tools/sw_fast_host_check.cpp -lm
// stdin:  "<nSwaptions> <trials> <blocksize> <seed> <lean 0|1> <source 0|1>" then per swaption
//         dStrike dCompounding dMaturity dTenor dPaymentInterval dYears, 11 yields, 30 factors (row-major 3 x 10)
// stdout: per swaption "sum sumsq fallbacks" (%.17g): the two sums of HJM_Swaption_Blocking.cpp:203-204 accumulated in
//         trial order, and how many trials left the fast range and were redone by generic_trial().
// source 0: tables read from a FastShared block (the batched kernel's layout); 1: from a OneSwaption (constant bank).
// tests/test_sw_oracle.py compares sum / trials with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../p3arsec_b200/csrc/sw_prepare.h"

using namespace swk;

template <bool LEAN, class SRC>
static double one_trial(const SRC &src, const double *tab, const double *tail, double *z, const SwParams &P, long long t, int &fallbacks)
{
    const int tid = (int)(t % THREADS);  // any lane: exercises the [draw][thread] indexing of z
    const int steps = LEAN ? P.start : FN - 1;
    const int swap_end = LEAN ? P.last_pay : P.len - 1;
    const bsm::PlainLogTab plt = {tab};
    const uint32_t x0 = ru_residue(P.seed + t * FD);
    normals<LEAN>(plt, tail, z, tid, x0, steps);
    const ExpTab<0> et = {tab + bsm::TAB_EXP, 0};  // the plain table: bank layout is a device-only concern
    uint32_t worst = trial_draws_zero(x0) ? EXP_HI_LIMIT : 0u;
    double disc;
    switch (P.start) {
        case 1: disc = path_and_payoff<LEAN, 1>(src, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
        case 2: disc = path_and_payoff<LEAN, 2>(src, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
        case 3: disc = path_and_payoff<LEAN, 3>(src, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
        default: disc = path_and_payoff<LEAN, -1>(src, et, z, tid, P.ddelt, P.swap_ddelt, P.start, swap_end, worst); break;
    }
    if (exp_range_left(worst)) {
        ++fallbacks;
        disc = generic_trial(P, FN, FF, t);
    }
    return disc;
}

int main()
{
    int n, blocksize, lean, source;
    long trials, seed;
    if (scanf("%d %ld %d %ld %d %d", &n, &trials, &blocksize, &seed, &lean, &source) != 6) return 2;
    static FastShared sh;
    alignas(16) static double z[FD * THREADS];
    bsm::fill_tables(sh.tab, 0, 1);
    swt::fill_tail(sh.tail, 0, 1);
    for (int i = 0; i < n; ++i) {
        sw_gpu_swaption s;
        double y[FN], f[FF * (FN - 1)];
        if (scanf("%lf %lf %lf %lf %lf %lf", &s.dStrike, &s.dCompounding, &s.dMaturity, &s.dTenor, &s.dPaymentInterval, &s.dYears) != 6) return 2;
        for (double &v : y) if (scanf("%lf", &v) != 1) return 2;
        for (double &v : f) if (scanf("%lf", &v) != 1) return 2;
        SwParams P;
        if (!prepare(P, s, FN, FF, y, f, seed + i, trials, blocksize)) {
            printf("invalid\n");
            continue;
        }
        OneSwaption one;
        to_one_swaption(one, P);
        for (int l = 0; l < FN; ++l) {  // what sw_sim_fast stages into shared memory per work item
            sh.fwd[l] = one.fwd[l];
            sh.pay[l] = one.pay[l];
            sh.xdp[l] = one.xdp[l];
            sh.xds[l] = one.xds[l];
            if (l < FN - 1) sh.fd[l] = one.fd[l];
        }
        double sum = 0, sumsq = 0;
        int fallbacks = 0;
        for (long long t = 0; t < P.sims; ++t) {
            double d;
            if (lean) d = source ? one_trial<true>(one, sh.tab, sh.tail, z, P, t, fallbacks) : one_trial<true>(sh, sh.tab, sh.tail, z, P, t, fallbacks);
            else d = source ? one_trial<false>(one, sh.tab, sh.tail, z, P, t, fallbacks) : one_trial<false>(sh, sh.tab, sh.tail, z, P, t, fallbacks);
            sum += d;
            sumsq += d * d;
        }
        printf("%.17g %.17g %d\n", sum, sumsq, fallbacks);
    }
    return 0;
}
