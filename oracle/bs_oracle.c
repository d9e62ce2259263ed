/*
 * bs_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * CPU restatement of the P3ARSEC blackscholes Map hot path, used as the parity checker for the CUDA
 * implementation.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (p3arsec_b200/, libbs_gpu.so, blackscholes_gpu) never does
 * and has no CPU fallback.
 *
 * Follows /root/reference/parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c:
 *   CNDF                 :128-184     BlkSchlsEqEuroNoDiv  :190-258
 *   Map body             :328-331     ERR_CHK              :333-340, :949-951
 *   loader               :696-739     AoS->SoA, otype      :747-767
 *   writer               :923-947
 *
 * PINNING.  The reference tree holds no golden vectors for this path (its DGrefval fixtures live in
 * the absent PARSEC tarball).  This restatement is pinned instead against outputs of the reference
 * itself: oracle/_ref/bs_ref_* are the unmodified reference sources compiled by oracle/Makefile, and
 * tests/golden/ holds input/output pairs produced by those binaries (tests/golden/make_golden.py).
 * tests/test_oracle.py requires bit-identical "%.18f" output on every golden pair, fp32 and fp64.
 * libm note: expf/logf/exp/log come from the host glibc, so bit-exactness holds on the image the
 * goldens were made on (glibc 2.39); both boxes of this project run that image.
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>

#define BS_PASTE2(a, b) a##b
#define BS_PASTE(a, b) BS_PASTE2(a, b)

/* ---- fptype = float : the build every P3ARSEC configuration ships (blackscholes.c:85) ---- */
#define FP float
#define SFX(name) BS_PASTE(name, _f32)
#define FP_EXP expf
#define FP_LOG logf
#define FP_SQRT sqrtf
#define FP_ABS fabsf
#define FP_SCAN "%f"
#include "bs_oracle_impl.h"
#undef FP
#undef SFX
#undef FP_EXP
#undef FP_LOG
#undef FP_SQRT
#undef FP_ABS
#undef FP_SCAN

/* ---- fptype = double : BASELINE.json configs[2] ("fptype=double build") ---- */
#define FP double
#define SFX(name) BS_PASTE(name, _f64)
#define FP_EXP exp
#define FP_LOG log
#define FP_SQRT sqrt
#define FP_ABS fabs
#define FP_SCAN "%lf"
#include "bs_oracle_impl.h"
#undef FP
#undef SFX
#undef FP_EXP
#undef FP_LOG
#undef FP_SQRT
#undef FP_ABS
#undef FP_SCAN

int bs_oracle_abi_version(void) { return 1; }
