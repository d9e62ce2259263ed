// swaptions_gpu.cpp -- drop-in driver for the swaptions Map on B200.
//
// Keeps the process surface of the reference driver, parsec-ff/pkgs/apps/swaptions/src/HJM_Securities.cpp
// (cited as SEC: below):
//   * command line  -ns <swaptions> -sm <simulations> -nt <threads> -sd <seed>, usage text and the error strings
//     (SEC:139-147,171-190,209-222); here -nt is the number of GPUs (clamped to the devices of the box);
//   * stdout: banner, "Number of Simulations: %d,  Number of threads: %d Number of swaptions: %d" (SEC:161-167,192);
//   * the portfolio set-up: swaption_seed from the first RanUnif draw, dYears / dStrike from the next two per
//     swaption, the 11-point yield curve and the 3 x 10 factor table (SEC:198,231-297);
//   * stderr: one "Swaption %d: [SwaptionPrice: %.10lf StdError: %.10lf] " line per swaption (SEC:357-358).
// What changes: the Map between the ROI markers (SEC:305-341) is one sw_gpu_price() call into libsw_gpu.so.
// The trial block size is the reference's BLOCK_SIZE (16): it only decides how many trials are simulated when
// -sm is not a multiple of it (HJM_Swaption_Blocking.cpp:156).
// Extra switches (not in the reference): -ieee, -lean select the kernel flavours of include/sw_gpu.h.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sw_gpu.h"

#ifdef ENABLE_PARSEC_HOOKS
#include <hooks.h>
#endif

namespace {

const int kBlockSize = 16;           // PARSEC HJM_type.h BLOCK_SIZE, passed at SEC:319
const int kDefaultTrials = 102400;   // PARSEC HJM_type.h DEFAULT_NUM_TRIALS, SEC:53
const int kN = 11, kFactors = 3;     // SEC:56,58

// PARSEC RanUnif.c: a counter-driven minimal-standard draw (state advances by one per call).
double ran_unif(long *s)
{
    long ix = *s;
    *s = ix + 1;
    ix *= 1513517L;
    ix %= 2147483647L;
    long k1 = ix / 127773L;
    ix = 16807L * (ix - k1 * 127773L) - k1 * 2836L;
    if (ix < 0) ix += 2147483647L;
    return ix * 4.656612875e-10;
}

void print_usage(const char *name)  // SEC:139-147
{
    fprintf(stderr, "Usage: %s OPTION [OPTIONS]...\n", name);
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "\t-ns [number of swaptions (should be > number of threads]\n");
    fprintf(stderr, "\t-sm [number of simulations]\n");
    fprintf(stderr, "\t-nt [number of threads]\n");
    fprintf(stderr, "\t-sd [random number seed]\n");
}

const double kFactorTable[kFactors][kN - 1] = {   // SEC:231-265
    {.01, .01, .01, .01, .01, .01, .01, .01, .01, .01},
    {.009048, .008187, .007408, .006703, .006065, .005488, .004966, .004493, .004066, .003679},
    {.001000, .000750, .000500, .000250, .000000, -.000250, -.000500, -.000750, -.001000, -.001250}};

}  // namespace

int main(int argc, char *argv[])
{
    int num_trials = kDefaultTrials, n_threads = 1, n_swaptions = 1;
    long seed = 1979;  // SEC:61
    unsigned flags = 0;

#ifdef PARSEC_VERSION
#define SW_STR_(x) #x
#define SW_STR(x) SW_STR_(x)
    printf("PARSEC Benchmark Suite Version " SW_STR(PARSEC_VERSION) "\n");
#else
    printf("PARSEC Benchmark Suite\n");
#endif
    fflush(NULL);
#ifdef ENABLE_PARSEC_HOOKS
    __parsec_bench_begin(__parsec_swaptions);
#endif

    if (argc == 1) {
        print_usage(argv[0]);
        exit(1);
    }
    for (int j = 1; j < argc; j++) {
        const bool has_value = j + 1 < argc;
        if (!strcmp("-sm", argv[j]) && has_value) num_trials = atoi(argv[++j]);
        else if (!strcmp("-nt", argv[j]) && has_value) n_threads = atoi(argv[++j]);
        else if (!strcmp("-ns", argv[j]) && has_value) n_swaptions = atoi(argv[++j]);
        else if (!strcmp("-sd", argv[j]) && has_value) seed = atoi(argv[++j]);
        else if (!strcmp("-ieee", argv[j])) flags |= SW_GPU_FLAG_IEEE;
        else if (!strcmp("-lean", argv[j])) flags |= SW_GPU_FLAG_LEAN;
        else {
            fprintf(stderr, "Error: Unknown option: %s\n", argv[j]);
            print_usage(argv[0]);
            exit(1);
        }
    }
    if (n_swaptions < n_threads) {  // SEC:186-190
        fprintf(stderr, "Error: Fewer swaptions than threads.\n");
        print_usage(argv[0]);
        exit(1);
    }
    printf("Number of Simulations: %d,  Number of threads: %d Number of swaptions: %d\n", num_trials, n_threads, n_swaptions);
    const long swaption_seed = (long)(2147483647L * ran_unif(&seed));  // SEC:198
    if (n_threads < 1 || n_threads > 1024) {                           // SEC:218-222 (MAX_THREAD)
        fprintf(stderr, "Number of threads must be between 1 and %d.\n", 1024);
        exit(1);
    }

    // the portfolio (SEC:276-296), held as the flat arrays the C ABI takes
    std::vector<sw_gpu_swaption> sw((size_t)n_swaptions);
    std::vector<double> yields((size_t)n_swaptions * kN), factors((size_t)n_swaptions * kFactors * (kN - 1));
    for (int i = 0; i < n_swaptions; i++) {
        sw[i].dYears = 5.0 + ((int)(60 * ran_unif(&seed))) * 0.25;
        sw[i].dStrike = 0.1 + ((int)(49 * ran_unif(&seed))) * 0.1;
        sw[i].dCompounding = 0;
        sw[i].dMaturity = 1.0;
        sw[i].dTenor = 2.0;
        sw[i].dPaymentInterval = 1.0;
        double *y = &yields[(size_t)i * kN];
        y[0] = .1;
        for (int j = 1; j <= kN - 1; ++j) y[j] = y[j - 1] + .005;
        memcpy(&factors[(size_t)i * kFactors * (kN - 1)], kFactorTable, sizeof(kFactorTable));
    }
    std::vector<double> mean((size_t)n_swaptions), err((size_t)n_swaptions);

    sw_gpu_ctx *ctx = nullptr;
    int st = sw_gpu_init(&ctx, n_threads, n_swaptions, kN, kFactors);
    if (st != SW_GPU_OK) {
        fprintf(stderr, "ERROR: sw_gpu_init: %s\n", sw_gpu_status_string(st));
        exit(1);
    }

#ifdef ENABLE_PARSEC_HOOKS
    __parsec_roi_begin();
#endif
    const auto t0 = std::chrono::steady_clock::now();
    st = sw_gpu_price(ctx, n_swaptions, sw.data(), yields.data(), factors.data(), swaption_seed, num_trials, kBlockSize, flags,
                      mean.data(), err.data());
    const double roi_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
#ifdef ENABLE_PARSEC_HOOKS
    __parsec_roi_end();
#endif
    if (st != SW_GPU_OK) {
        fprintf(stderr, "ERROR: sw_gpu_price: %s: %s\n", sw_gpu_status_string(st), sw_gpu_last_error(ctx));
        sw_gpu_fini(ctx);
        exit(1);
    }
#ifndef ENABLE_PARSEC_HOOKS
    sw_gpu_timing tm;
    sw_gpu_get_timing(ctx, &tm);
    printf("roi.time|%.9f\n", roi_s);  // the line the PARSEC hooks library prints (hooks.c:243)
    printf("gpu.kernels_ms|%.6f\ngpu.devices|%d\ngpu.trials|%llu\n", tm.roi_ms, sw_gpu_num_shards(ctx), tm.trials_simulated);
#endif

    for (int i = 0; i < n_swaptions; i++)  // SEC:356-360
        fprintf(stderr, "Swaption %d: [SwaptionPrice: %.10lf StdError: %.10lf] \n", i, mean[i], err[i]);

    sw_gpu_fini(ctx);
#ifdef ENABLE_PARSEC_HOOKS
    __parsec_bench_end();
#endif
    return 0;  // the reference returns main's own iSuccess, which is never assigned after its initialisation to 0 (SEC:152,377)
}
