/*
 * bs_gpu.h -- C ABI of the B200 blackscholes Map (libbs_gpu.so).
 *
 * This is the in-process seam that replaces the reference's host parallel-for around the per-option
 * BlkSchlsEqEuroNoDiv + CNDF evaluation.  All reference citations are relative to
 * /root/reference/parsec-ff/pkgs/apps/blackscholes/src/ unless a longer path is given.
 *
 *   reference construct                                         replaced by
 *   ----------------------------------------------------------  ---------------------------------
 *   `map m; m.run(); m.wait();`          blackscholes.c:825-827   bs_gpu_price()
 *     = ff_Map<int>::svc, NUM_RUNS x parallel_for_thid(...)       (one real kernel launch per run)
 *                                        blackscholes.c:315-348
 *   bs_thread() (serial/pthreads/OpenMP) blackscholes.c:607-658   bs_gpu_price()
 *   skepu2::Map<7>(mapFunction) + run loop
 *                              blackscholes_skepu.cpp:324-334     bs_gpu_price()
 *   SoA malloc + AoS->SoA staging        blackscholes.c:747-767   bs_gpu_host_buffer() (pinned) +
 *                                                                 the H2D copies inside bs_gpu_price()
 *   `prices = malloc(...)`               blackscholes.c:725       bs_gpu_host_buffer(BS_BUF_PRICES)
 *   ERR_CHK body + "Num Errors" total    blackscholes.c:333-340,  err_chk=1 + *num_errors,
 *                                        :949-951                 bs_gpu_errors()
 *   static contiguous partition over workers                      contiguous shards over GPUs,
 *     parsec-ff/pkgs/libs/fastflow/ff/parallel_for_internals.hpp:498-518   bs_gpu_shard()
 *   ROI markers __parsec_roi_begin/end   blackscholes.c:781-783,  bs_gpu_get_timing()
 *                                        :912-914
 *
 * Conventions: plain C types only; no exceptions cross the boundary (the reference builds with
 * -fno-exceptions, parsec-ff/config/gcc.bldconf:75); every function returns BS_GPU_OK (0) or a
 * negative bs_gpu_status.  One caller thread per context.  There is no CPU fallback: without a
 * usable CUDA device bs_gpu_init() fails with BS_GPU_ERR_NO_DEVICE.
 */
#ifndef BS_GPU_H
#define BS_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BS_GPU_ABI_VERSION 2

typedef struct bs_gpu_ctx bs_gpu_ctx;

typedef enum bs_gpu_status {
    BS_GPU_OK = 0,
    BS_GPU_ERR_INVALID = -1,   /* bad argument                                     */
    BS_GPU_ERR_NO_DEVICE = -2, /* no CUDA device / fewer devices than requested    */
    BS_GPU_ERR_CUDA = -3,      /* a CUDA call failed; see bs_gpu_last_error()      */
    BS_GPU_ERR_NOMEM = -4,     /* host or device allocation failed                 */
    BS_GPU_ERR_STATE = -5      /* call not valid in the context's current state    */
} bs_gpu_status;

/* The SoA streams, named after the reference globals (blackscholes.c:102-111).  SPTPRICE..OTIME and
 * PRICES/DGREFVAL hold `fptype` (float or double, chosen at init); OTYPE holds int32 (1 = put,
 * 0 = call, blackscholes.c:761).  DGREFVAL is the ERR_CHK reference column (OptionData.DGrefval,
 * blackscholes.c:99), kept as a compact stream instead of a 36-byte-stride AoS field. */
typedef enum bs_gpu_buffer {
    BS_BUF_SPTPRICE = 0,
    BS_BUF_STRIKE = 1,
    BS_BUF_RATE = 2,
    BS_BUF_VOLATILITY = 3,
    BS_BUF_OTIME = 4,
    BS_BUF_OTYPE = 5,
    BS_BUF_PRICES = 6,
    BS_BUF_DGREFVAL = 7,
    BS_BUF_COUNT = 8
} bs_gpu_buffer;

/* How exp/log/sqrt/divide are evaluated.  BS_MATH_DEFAULT resolves to BS_MATH_FAST for BOTH precisions -- this
 * is what bs_gpu_init() (the entry the reference patch calls) uses; ask for BS_MATH_IEEE or BS_MATH_REFERENCE
 * through bs_gpu_init_ex() to get the reference's operation order.
 *
 * Accuracy contract (measured worst cases: DESIGN.md section 4 and profiles/r02_fp32_adversarial.json):
 *   fp32, operands in the PARSEC inputgen range (spot, strike <= 128; 0.05 <= v <= 0.65; 0.05 <= t <= 1):
 *         every mode: |price - reference CPU price| <= 1e-4, the reference's own ERR_CHK threshold
 *         (blackscholes.c:335).  BS_MATH_REFERENCE reproduces the reference CPU output bit for bit.
 *   fp32, larger operands: BS_MATH_REFERENCE keeps the flat 1e-4 bound; BS_MATH_FAST / BS_MATH_IEEE keep the same
 *         number of ulps, i.e. the bound scales as 1e-4 * max(1, max(spot, strike) / 128): a price near 1000 is
 *         itself quantised to 6e-5 in fp32 and the reference's own fp32 build is 1.7e-4 from its fp64 build there.
 *   fp64: |delta| <= 1e-9 * |reference| + 1e-12 in every mode (FAST: ~2 ulp building blocks; IEEE: only the last
 *         ulp of exp()/log() can differ from the fp64 CPU build; REFERENCE: bit-identical to it). */
typedef enum bs_gpu_math {
    BS_MATH_DEFAULT = 0,  /* the library's default: BS_MATH_FAST (fp32 and fp64)                                  */
    BS_MATH_IEEE = 1,     /* libdevice exp/log/sqrt and IEEE-rounded divides, reference operation order; fp32 is
                             pure fp32 (the reference's double-literal promotions are not imitated)               */
    BS_MATH_FAST = 2,     /* fp32: MUFU ex2/lg2/sqrt/rcp with folded constants, Horner CNDF, branch-free;
                             fp64: table-driven exp/log, MUFU-seeded reciprocals (bs_math_f64.h), with the IEEE
                             path taken for operands outside its range (t = 0, v = 0, denormals, overflow)       */
    BS_MATH_REFERENCE = 3 /* fp32: the reference's fp32 build as compiled -- same operation order AND its double
                             promotions (blackscholes.c:154-158,164-175,232,252-253), every operation individually
                             rounded, expf/logf = glibc 2.39's algorithms (csrc/bs_libm_f32.h, pinned against the
                             host libm over all 2^32 floats): the reference CPU prices, bit for bit.  The validation
                             mode: several times slower than FAST.  fp64: the operation order of BS_MATH_IEEE
                             with exp/log = glibc 2.39's double algorithms (csrc/bs_libm_f64.h, pinned against the
                             host libm on 2 x 10^9 arguments per function): the fp64 CPU prices, bit for bit      */
} bs_gpu_math;

/* bs_gpu_config.flags */
#define BS_GPU_FLAG_NO_HOST_STAGING 1u /* no pinned host buffers: device-resident data only (1B set) */
#define BS_GPU_FLAG_WITH_DGREFVAL 2u   /* allocate the DGREFVAL stream (needed for err_chk=1)       */
#define BS_GPU_FLAG_NO_GRAPH 4u        /* launch runs as plain stream launches, not a CUDA graph     */
#define BS_GPU_FLAG_NO_SUBSHARDS 32u   /* bs_gpu_price(): never cut a shard into sub-shards (see bs_gpu_price)        */
#define BS_GPU_FLAG_PDL 16u            /* opt-in: programmatic dependent launch between the runs (run j+1 is scheduled
                                          and issues its first loads while run j drains; stores stay ordered).
                                          Measured neutral to harmful on B200, hence off by default */
#define BS_GPU_FLAG_ASYNC_DISCOVERY 8u /* bs_gpu_init_ex does not touch CUDA at all: driver initialisation and device
                                          discovery run in the background too, num_gpus is an upper bound that is
                                          clamped to the devices (and options) present, and "no device" is reported
                                          by the first call that needs one.  For drivers that parse input meanwhile. */

typedef struct bs_gpu_config {
    size_t struct_size;  /* = sizeof(bs_gpu_config), for ABI growth                                  */
    size_t num_options;  /* N, as read from the input header (blackscholes.c:701)                    */
    int fp_bytes;        /* sizeof(fptype): 4 or 8 (blackscholes.c:85)                               */
    int num_gpus;        /* G >= 1                                                                   */
    const int *devices;  /* G CUDA device ordinals, or NULL for 0..G-1                                */
    unsigned flags;      /* BS_GPU_FLAG_*                                                            */
    int math;            /* bs_gpu_math                                                              */
    int threads_per_block; /* 0 = default; else 32..256, multiple of 32                              */
    int blocks_per_sm;     /* 0 = default (all the CTAs the SM can hold)                             */
    int unroll;            /* 0 = default; else 1, 2 or 4 independent 16-byte groups per thread-trip */
    int variant;           /* 0 = default; bit 0: software-pipelined loads; bit 1: DIAGNOSTIC traffic probe
                              (no pricing, same streams); bit 2: TMA variant (inputs moved by cp.async.bulk
                              into a shared-memory ring) -- see DESIGN.md; bit 3: DIAGNOSTIC fault injection
                              (every Map launch requests 1 MiB of shared memory and is rejected by the runtime:
                              the calls must then return BS_GPU_ERR_CUDA with text, never prices)             */
} bs_gpu_config;

typedef struct bs_gpu_timing {
    double h2d_ms;   /* last upload: max over devices, CUDA events on each device's stream           */
    double roi_ms;   /* last run:   max over devices, CUDA events bracketing the NUM_RUNS launches    */
    double d2h_ms;   /* last download                                                                */
    double wall_ms;  /* host wall clock of the last bs_gpu_price()/upload/run/download call          */
    double pipeline_ms; /* last bs_gpu_price(): device time from its first copy/launch to its last copy,
                           max over devices (H2D, launches and D2H overlap, so this is NOT the sum of the
                           three figures above, which then are: time until the last input chunk landed,
                           first-to-last kernel, and last kernel to last price chunk)                  */
    unsigned long long kernel_launches; /* pricing kernels launched by the last run, all devices     */
    unsigned long long h2d_bytes;       /* bytes copied by the last upload                           */
    unsigned long long d2h_bytes;       /* bytes copied by the last download                         */
} bs_gpu_timing;

/* Number of usable CUDA devices (>= 0), or a negative bs_gpu_status. */
int bs_gpu_device_count(void);

/* Optional, and only effective BEFORE the process's first CUDA call (so before bs_gpu_device_count / bs_gpu_init):
 * hide all but the first `max_gpus` devices from this process (CUDA_VISIBLE_DEVICES; an existing list is cut to its
 * first max_gpus entries).  Driver initialisation and process exit cost time per VISIBLE device -- measured on an
 * 8 x B200 box: cuInit 4.4-5.4 s with eight devices visible, 0.36-0.49 s with one -- so a one-shot driver that will use
 * G devices (<nthreads>, blackscholes.c:694) should say so first. */
int bs_gpu_limit_devices(int max_gpus);

/* Create a context pricing `num_options` options of `fp_bytes`-wide fptype on the first `num_gpus`
 * devices.  Spawns one host thread per device (each owns that device's primary context, stream and
 * CUDA graphs), splits [0,N) into contiguous shards (first N%G shards get one extra option, as
 * ff's static partitioner does), allocates the device SoA arena + prices, and host staging for every
 * stream including DGREFVAL.  Replaces blackscholes.c:747-758 (+ :725).
 * Returns as soon as the staging buffers exist: the device contexts and arenas come up in the
 * background (so a loader can parse while CUDA initialises) and the device threads pin the staging
 * buffers (cudaHostRegister) right before the first copy.  A device-side setup failure is reported by the
 * first later call that needs the devices. */
int bs_gpu_init(bs_gpu_ctx **ctx, int num_gpus, size_t num_options, int fp_bytes);

/* Same, with explicit device list / flags / math mode / launch geometry. */
int bs_gpu_init_ex(bs_gpu_ctx **ctx, const bs_gpu_config *cfg);

/* Pinned host array for one stream (N elements), written by the loader / read by the writer in place
 * of the reference's malloc'ed sptprice/strike/rate/volatility/otime/otype/prices.  NULL when the
 * context was created with BS_GPU_FLAG_NO_HOST_STAGING or `which` is out of range.  Writing to an
 * input buffer after a bs_gpu_price() call requires bs_gpu_mark_dirty(). */
void *bs_gpu_host_buffer(bs_gpu_ctx *ctx, int which);
int bs_gpu_mark_dirty(bs_gpu_ctx *ctx);

/* The hot path.  H2D of the input streams if they are dirty, then `num_runs` real launches of the
 * pricing kernel per device (every run re-reads all inputs and rewrites all prices; nothing is
 * cached or hoisted across runs), then D2H of the prices into the pinned PRICES buffer.  Blocking.
 * Copies and launches are pipelined.  A shard whose streams exceed 256 MiB is cut into contiguous sub-shards,
 * each larger than the L2 cache (so every run over it still streams its inputs from HBM): sub-shard i runs all
 * its num_runs launches while sub-shard i+1 is still on the PCIe bus and the prices of sub-shard i-1 travel
 * back -- the time-multiplexed twin of sharding over several GPUs.  Smaller shards keep run order: run 0 is
 * launched chunk by chunk as the input chunks land, runs 1..num_runs-2 are whole-shard launches, and the last
 * run is launched chunk by chunk with each price chunk copied back as soon as it is complete.
 * With err_chk != 0 every run also evaluates |DGrefval - price| >= 1e-4 (blackscholes.c:335) and
 * *num_errors receives the total over all runs, i.e. the number the reference prints as
 * "Num Errors" (blackscholes.c:950).  num_errors may be NULL. */
int bs_gpu_price(bs_gpu_ctx *ctx, int num_runs, int err_chk, unsigned long long *num_errors);

/* The three phases of bs_gpu_price() on their own, for measurement and for device-resident sets. */
int bs_gpu_upload(bs_gpu_ctx *ctx);
int bs_gpu_run(bs_gpu_ctx *ctx, int num_runs, int err_chk, unsigned long long *num_errors);
int bs_gpu_download(bs_gpu_ctx *ctx);

/* The CAF Map's entry point (blackscholes.c:482-570, driver :771-778,:874-885): the reference's CAF_V3 build sends
 * its actors ONE message holding the options as an array of
 *     struct DataCont { int otype; float sptprice, strike, rate, volatility, otime; };      (24 bytes, :482-489)
 * and receives a vector of prices (OutVec, :492).  `records` points at `num_records` such records in ordinary
 * (pageable) host memory -- e.g. data_vec.data() -- and must equal the context's num_options; the context must have
 * fp_bytes = 4.  The records are copied H2D as they are and gathered into the SoA streams on the device, then
 * `num_runs` real launches, then the prices are copied into out_prices[0..num_records).  Shards over several GPUs
 * exactly like bs_gpu_price().  (The reference's message holds 2N records, the first N zero-initialised, :772-777;
 * a caller that reproduces that passes num_records = 2N and gets NaN for the zero records, as the reference does.) */
int bs_gpu_price_aos(bs_gpu_ctx *ctx, const void *records, size_t num_records, float *out_prices, int num_runs);

/* Fill the DEVICE input streams (and DGREFVAL when allocated) with the synthetic inputgen sequence:
 * global option i = table[(first_index + i) % 1000] (p3arsec_b200/csrc/bs_option_table.h).  Used for
 * sets that cannot pass through a text file (1B options).  Marks the device copy current. */
int bs_gpu_fill_synthetic(bs_gpu_ctx *ctx, unsigned long long first_index);

/* Copy `count` elements of a device stream, starting at global option index `first`, to `dst`
 * (pageable or pinned).  For spot checks of device-resident sets. */
int bs_gpu_read_device(bs_gpu_ctx *ctx, int which, size_t first, size_t count, void *dst);

/* Indices (ascending, at most `cap`) of the options that failed ERR_CHK in the LAST run of the most
 * recent err_chk call; returns how many were written, or a negative status.  At most
 * BS_GPU_MAX_ERROR_LIST offenders per device are recorded. */
#define BS_GPU_MAX_ERROR_LIST 65536
long long bs_gpu_errors(bs_gpu_ctx *ctx, long long *idx, size_t cap);

/* Shard g of the context: CUDA device ordinal, first global option index and option count. */
int bs_gpu_num_shards(bs_gpu_ctx *ctx);
int bs_gpu_shard(bs_gpu_ctx *ctx, int g, int *device, size_t *first, size_t *count);

int bs_gpu_get_timing(bs_gpu_ctx *ctx, bs_gpu_timing *out);

/* The math mode / launch geometry actually in use (after defaults were resolved). */
int bs_gpu_get_launch(bs_gpu_ctx *ctx, int *math, int *threads_per_block, int *blocks);

/* Text of the most recent failure on this context ("" if none); valid until the next call. */
const char *bs_gpu_last_error(bs_gpu_ctx *ctx);
const char *bs_gpu_status_string(int status);
int bs_gpu_abi_version(void);

/* Join the device threads and free all device and pinned memory.  NULL is a no-op. */
void bs_gpu_fini(bs_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* BS_GPU_H */
