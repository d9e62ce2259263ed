/*
 * hooks.h -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
 *
 * Header-only stand-in for PARSEC's hooks.h, which is owned by the PARSEC 3.0 tarball and is
 * therefore absent from the P3ARSEC overlay.  It lets the unmodified
 * reference driver (parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c, built with
 * -DENABLE_PARSEC_HOOKS) compile, and prints the ROI wall time in the same "roi.time|<seconds>"
 * line the real hooks library prints (parsec-hooks/pkgs/libs/hooks/src/hooks.c:185-246), so the
 * CPU baseline is timed at exactly the reference's own ROI boundary
 * (blackscholes.c:781-783 and :912-914).  No energy counters (Mammut is un-vendored).
 */
#ifndef BS_ORACLE_HOOKS_SHIM_H
#define BS_ORACLE_HOOKS_SHIM_H

#include <stdio.h>
#include <time.h>

enum __parsec_benchmark { __parsec_blackscholes = 1, __parsec_swaptions = 10 };

static double bs_shim_t_begin_;
static double bs_shim_t_end_;

static inline double bs_shim_now_(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline void __parsec_bench_begin(enum __parsec_benchmark b) {
    (void)b;
    printf("[HOOKS] shim (oracle/hooks_shim/hooks.h)\n");
    fflush(NULL);
}

static inline void __parsec_roi_begin(void) {
    printf("[HOOKS] Entering ROI\n");
    fflush(NULL);
    bs_shim_t_begin_ = bs_shim_now_();
}

static inline void __parsec_roi_end(void) {
    bs_shim_t_end_ = bs_shim_now_();
    printf("roi.time|%.9f\n", bs_shim_t_end_ - bs_shim_t_begin_);
    printf("[HOOKS] Leaving ROI\n");
    fflush(NULL);
}

static inline void __parsec_bench_end(void) {
    fflush(NULL);
    printf("[HOOKS] Total time spent in ROI: %.3fs\n", bs_shim_t_end_ - bs_shim_t_begin_);
    printf("[HOOKS] Terminating\n");
}

#endif
