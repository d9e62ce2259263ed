// Host-side accuracy check of p3arsec_b200/csrc/bs_math_f64.h (hardware seeds emulated at 2^-20).
//   g++ -O2 -std=c++17 -o math_f64_host_check tools/math_f64_host_check.cpp -lm
// Prints one "name max_err" line per building block (ulps of the result; for log: ulps of max(1,|log x|), the
// absolute accuracy its consumer d1 needs); `price` mode reads "s k r v t otype" rows on
// stdin and prints the fast-path price with %.17g (tests/test_math_f64.py compares them with the oracle).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../p3arsec_b200/csrc/bs_math_f64.h"

static double ulps(double got, long double ref)
{
    if (ref == 0) return got == 0 ? 0 : 1e300;
    int e;
    frexpl(ref, &e);
    long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - ref) / ulp);
}

static double TAB[bsm::TAB_DOUBLES];
static double TAB256[bsm::TAB256_DOUBLES];

// absolute error in units of ulp(max(1, |ref|)): what matters for log feeding d1
static double abs_units(double got, long double ref)
{
    long double scale = fabsl(ref) > 1 ? fabsl(ref) : 1.0L;
    int e;
    frexpl(scale, &e);
    return (double)(fabsl((long double)got - ref) / ldexpl(1.0L, e - 53));
}

int main(int argc, char **argv)
{
    bsm::fill_tables(TAB, 0, 1);
    bsm::fill_tables256(TAB256, 0, 1);
    if (argc > 1 && !strcmp(argv[1], "price")) {
        double s, k, r, v, t;
        int o;
        while (scanf("%lf %lf %lf %lf %lf %d", &s, &k, &r, &v, &t, &o) == 6) {
            bool ok;
            double p = bsm::price_f64_fast(s, k, r, v, t, o, &ok, TAB256);
            printf("%.17g %d\n", p, ok ? 1 : 0);
        }
        return 0;
    }
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    const int N = 2000000;
    double w_rcp = 0, w_rsq = 0, w_exp = 0, w_log = 0, w_exp_abs = 0;
    for (int i = 0; i < N; i++) {
        double b = exp((u01(rng) - 0.5) * 60.0);  // 1e-13 .. 1e13
        w_rcp = fmax(w_rcp, ulps(bsm::rcp_f64(b), 1.0L / (long double)b));
        w_rsq = fmax(w_rsq, ulps(bsm::rsqrt_f64(b), 1.0L / sqrtl((long double)b)));
        w_log = fmax(w_log, abs_units(bsm::log_f64(b, TAB), logl((long double)b)));
        double x = -u01(rng) * (i % 4 == 0 ? 700.0 : 40.0);
        w_exp = fmax(w_exp, ulps(bsm::exp_f64(x, TAB), expl((long double)x)));
        double xs = -u01(rng) * 1e-3;
        w_exp_abs = fmax(w_exp_abs, ulps(bsm::exp_f64(xs, TAB), expl((long double)xs)));
    }
    printf("rcp %.3f\nrsqrt %.3f\nlog %.3f\nexp %.3f\nexp_small %.3f\n", w_rcp, w_rsq, w_log, w_exp, w_exp_abs);
    // the unguarded exp core, both signs, |x| < 500
    double w_pp = 0, w_pn = 0;
    for (int i = 0; i < N; i++) {
        double x = (u01(rng) - 0.5) * (i % 4 == 0 ? 1000.0 : 1.0);
        double ep = bsm::exp_core_f64(x, TAB), en = bsm::exp_core_f64(-x, TAB);
        w_pp = fmax(w_pp, ulps(ep, expl((long double)x)));
        w_pn = fmax(w_pn, ulps(en, expl(-(long double)x)));
    }
    printf("exp_pair_pos %.3f\nexp_pair_neg %.3f\n", w_pp, w_pn);
    // the 256-entry blocks of the blackscholes fp64 kernel
    double w_e256 = 0, w_l256 = 0, w_l256_near1 = 0;
    for (int i = 0; i < N; i++) {
        double x = (u01(rng) - 0.5) * (i % 4 == 0 ? 1000.0 : 80.0);
        w_e256 = fmax(w_e256, ulps(bsm::exp256_core_f64(x, TAB256), expl((long double)x)));
        double b = exp((u01(rng) - 0.5) * 60.0);
        w_l256 = fmax(w_l256, abs_units(bsm::log256_f64(b, TAB256), logl((long double)b)));
        double c1 = 1.0 + (u01(rng) - 0.5) * 2e-3;
        if (c1 != 1.0) w_l256_near1 = fmax(w_l256_near1, abs_units(bsm::log256_f64(c1, TAB256), logl((long double)c1)));
    }
    printf("exp256 %.3f\nlog256 %.3f\nlog256_near_1_abs %.3f\nexp256_0 %.17g\nlog256_1 %.17g\n", w_e256, w_l256, w_l256_near1,
           bsm::exp256_core_f64(0.0, TAB256), bsm::log256_f64(1.0, TAB256));
    printf("exp_below_-708 %g\nexp_0 %.17g\nlog_1 %.17g\n", bsm::exp_f64(-709.5, TAB), bsm::exp_f64(0.0, TAB), bsm::log_f64(1.0, TAB));
    // log(x) for x within 1e-3 of 1, error relative to the result
    double w_log1 = 0;
    for (int i = 0; i < N; i++) {
        double c = 1.0 + (u01(rng) - 0.5) * 2e-3;
        if (c == 1.0) continue;
        w_log1 = fmax(w_log1, abs_units(bsm::log_f64(c, TAB), logl((long double)c)));
    }
    printf("log_near_1_abs %.3f\n", w_log1);
    return 0;
}
