/*
 * sw_gpu.h -- C ABI of the B200 swaptions Map (libsw_gpu.so).  SURVEY.md section 8f, rank 4: the sibling
 * financial Map of P3ARSEC that sits behind the same ParallelFor::parallel_for_thid call as blackscholes.
 *
 * Reference citations are relative to /root/reference/parsec-ff/pkgs/apps/swaptions/src/.
 *
 *   reference construct                                                     replaced by
 *   ----------------------------------------------------------------------  ------------------------------
 *   ff::ParallelFor pf; pf.parallel_for_thid(0, nSwaptions, 1, PARFOR_STATIC(0), body, nThreads)
 *                                              HJM_Securities.cpp:311-323     sw_gpu_price()
 *   skepu2::Map<1>(mapFunction) over Vector<parm>
 *                                              HJM_Securities_skepu.cpp:43-57,204-214   sw_gpu_price()
 *   worker() (serial / pthreads), TBB Worker   HJM_Securities.cpp:72-95,100-136          sw_gpu_price()
 *   the Map body HJM_Swaption_Blocking(out, dStrike, dCompounding, dMaturity, dTenor, dPaymentInterval,
 *       iN, iFactors, dYears, pdYield, ppdFactors, swaption_seed + i, NUM_TRIALS, BLOCK_SIZE, tid)
 *                                              HJM_Swaption_Blocking.cpp:20-222           the kernels behind it
 *   swaptions[i].dSimSwaptionMeanPrice / dSimSwaptionStdError = pdSwaptionPrice[0] / [1]
 *                                              HJM_Securities.cpp:322-323     mean[i] / std_error[i]
 *   ROI markers                                HJM_Securities.cpp:301-303,343-345         sw_gpu_get_timing()
 *
 * The unit of work is one Monte-Carlo trial: trial t of swaption i consumes the draws
 * [t*(iN-1)*iFactors, (t+1)*(iN-1)*iFactors) of the counter stream that starts at swaption_seed + i, and the
 * reference simulates ceil(lTrials/BLOCKSIZE)*BLOCKSIZE of them (HJM_Swaption_Blocking.cpp:156) while dividing by
 * lTrials (:212-214); both are kept.  Trials are independent, so they are spread over all threads of all devices
 * and the two sums are reduced in a fixed (deterministic) tree; swaptions are sharded contiguously over devices
 * with no collective.  There is no CPU fallback.
 *
 * Conventions as in bs_gpu.h: plain C types, no exceptions, 0 or a negative status, one caller thread per context.
 */
#ifndef SW_GPU_H
#define SW_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SW_GPU_ABI_VERSION 1
#define SW_GPU_MAX_N 32       /* iN       <= 32 */
#define SW_GPU_MAX_FACTORS 8  /* iFactors <= 8  */

typedef struct sw_gpu_ctx sw_gpu_ctx;

typedef enum sw_gpu_status {
    SW_GPU_OK = 0,
    SW_GPU_ERR_INVALID = -1,   /* bad argument, or a swaption whose time indices run off the HJM path (the
                                  reference would index past its heap vectors there)                          */
    SW_GPU_ERR_NO_DEVICE = -2,
    SW_GPU_ERR_CUDA = -3,
    SW_GPU_ERR_NOMEM = -4
} sw_gpu_status;

/* The scalar fields of `parm` (PARSEC HJM_type.h) that the Map body receives, in the order of the call at
 * HJM_Securities.cpp:314-318.  pdYield and ppdFactors travel as flat arrays (see sw_gpu_price). */
typedef struct sw_gpu_swaption {
    double dStrike;
    double dCompounding;
    double dMaturity;
    double dTenor;
    double dPaymentInterval;
    double dYears;
} sw_gpu_swaption;

/* sw_gpu_price flags */
#define SW_GPU_FLAG_IEEE 1u  /* evaluate with libdevice exp/log, IEEE divides and the reference's operation order
                                (explicitly rounded, no FMA contraction) -- the shape-generic kernel; default for
                                any shape other than iN = 11, iFactors = 3                                         */
#define SW_GPU_FLAG_LEAN 2u  /* simulate only what the price depends on: time steps 1..iSwapStartTimeIndex and the
                                discount factors up to the last swap payment.  The reference simulates the whole
                                iN x iN path and every discount factor although HJM_Swaption_Blocking.cpp:167-198
                                reads only column 0 down to the swap start and the start row; prices are identical
                                whenever all intermediate values are finite.  Off by default: the default kernel
                                does the reference's full amount of work per trial                                 */

#define SW_GPU_FLAG_BATCHED 4u /* always simulate all swaptions in one kernel launch (tables in shared memory).  By default the
                                full-work kernel gives a swaption with >= 262144 trials a launch of its own, with its factor /
                                drift / yield tables in the kernel-parameter constant bank, which is 10 % faster (DESIGN.md
                                9.3); results are identical either way up to the order of the final summation        */

typedef struct sw_gpu_timing {
    double roi_ms;       /* last sw_gpu_price: device time of the kernels, max over devices (CUDA events)          */
    double wall_ms;      /* host wall clock of the last sw_gpu_price call, copies included                         */
    unsigned long long kernel_launches;
    unsigned long long trials_simulated; /* sum over swaptions of ceil(lTrials/BLOCKSIZE)*BLOCKSIZE               */
    unsigned long long h2d_bytes, d2h_bytes;
} sw_gpu_timing;

int sw_gpu_device_count(void);

/* One context per process: `num_gpus` devices (0..num_gpus-1), at most `max_swaptions` per call, fixed path shape
 * (iN time points, iFactors factors: HJM_Securities.cpp:56,58). */
int sw_gpu_init(sw_gpu_ctx **ctx, int num_gpus, int max_swaptions, int iN, int iFactors);
/* Same on explicit device ordinals (one rank per GPU under torchrun: devices = {LOCAL_RANK}, num_gpus = 1). */
int sw_gpu_init_devices(sw_gpu_ctx **ctx, const int *devices, int num_gpus, int max_swaptions, int iN, int iFactors);

/* The Map.  swaptions: nSwaptions scalar records; pdYield: nSwaptions x iN (HJM_Securities.cpp:287-290);
 * ppdFactors: nSwaptions x iFactors x (iN-1), row-major (:292-295).  Swaption i is simulated with the counter
 * stream starting at swaption_seed + i (:319).  mean[i], std_error[i] receive what the reference stores at
 * :322-323.  Blocking. */
int sw_gpu_price(sw_gpu_ctx *ctx, int nSwaptions, const sw_gpu_swaption *swaptions, const double *pdYield,
                 const double *ppdFactors, long swaption_seed, long lTrials, int BLOCKSIZE, unsigned flags,
                 double *mean, double *std_error);

/* Launch geometry override for tuning: CTAs per SM (0 = default) and trials per thread per work item (0 = auto). */
int sw_gpu_set_geometry(sw_gpu_ctx *ctx, int ctas_per_sm, int trials_per_thread);

int sw_gpu_get_timing(sw_gpu_ctx *ctx, sw_gpu_timing *out);
int sw_gpu_num_shards(sw_gpu_ctx *ctx);
/* Shard g of the last sw_gpu_price call: device ordinal, first swaption, swaption count. */
int sw_gpu_shard(sw_gpu_ctx *ctx, int g, int *device, int *first, int *count);
const char *sw_gpu_last_error(sw_gpu_ctx *ctx);
const char *sw_gpu_status_string(int status);
int sw_gpu_abi_version(void);
void sw_gpu_fini(sw_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SW_GPU_H */
