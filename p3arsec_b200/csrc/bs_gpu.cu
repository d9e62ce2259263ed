/*
 * bs_gpu.cu -- implementation of the C ABI in include/bs_gpu.h (libbs_gpu.so).
 *
 * Host-side structure (north_star item 3): one host thread per device.  Each thread binds its device
 * once, owns that device's stream, events, CUDA graphs and its contiguous shard of the option range,
 * and executes commands (upload / run / download / fill) posted by the caller thread; a command is
 * complete when every device thread has finished it.  There is no inter-device traffic: shards are
 * independent, exactly like the worker ranges of the reference's static parallel-for
 * (/root/reference/parsec-ff/pkgs/libs/fastflow/ff/parallel_for_internals.hpp:498-518).
 *
 * Device memory layout (per shard, count = options in the shard): ONE cudaMalloc arena holding
 *   [sptprice | strike | rate | volatility | otime | otype | prices | DGrefval]
 * each stream padded to a multiple of 256 bytes so that every stream base is 256-byte aligned and all
 * 128-bit vector accesses of the kernel are aligned whatever `count` is.  (The reference packs the five
 * fp streams back to back without padding, blackscholes.c:750-755, which would misalign them for
 * count % 4 != 0.)
 */
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

#include "../../include/bs_gpu.h"
#include "bs_kernels.cuh"
#include "bs_option_table.h"

namespace {

enum Cmd { CMD_NONE = 0, CMD_SETUP, CMD_PIN, CMD_UPLOAD, CMD_RUN, CMD_DOWNLOAD, CMD_PRICE, CMD_FILL, CMD_AOS_IN, CMD_DOWNLOAD_TO, CMD_TEARDOWN, CMD_EXIT };

enum { PIPE_CHUNKS = 8 };  // chunks of the first / last run of a pipelined bs_gpu_price()

struct GraphKey {
    int num_runs, err_chk;  // err_chk: 0 = off, 1 = on and the last run records offenders, 2 = on, nothing recorded
    size_t first, count;    // option range of the shard the launches cover
    bool operator<(const GraphKey &o) const
    {
        if (num_runs != o.num_runs) return num_runs < o.num_runs;
        if (err_chk != o.err_chk) return err_chk < o.err_chk;
        if (first != o.first) return first < o.first;
        return count < o.count;
    }
};

struct Shard {
    int index = 0;
    int device = 0;
    size_t first = 0, count = 0;
    int sm_count = 0;
    // device memory
    char *arena = nullptr;
    size_t arena_bytes = 0;
    void *d[BS_BUF_COUNT] = {nullptr};
    unsigned long long *d_err_count = nullptr;
    unsigned int *d_list_count = nullptr;
    long long *d_list = nullptr;
    char *d_table = nullptr;  // synthetic base table, built on first fill
    char *d_aos = nullptr;    // CAF DataCont records of the shard (bs_gpu_price_aos), allocated on first use
    size_t d_aos_bytes = 0;
    // execution
    cudaStream_t stream = nullptr;       // launches (and the un-pipelined copies)
    cudaStream_t copy_stream = nullptr;  // H2D (and, in the chunked scheme, D2H) of a pipelined bs_gpu_price()
    cudaStream_t d2h_stream = nullptr;   // D2H of the sub-shard scheme: the two copy engines work at the same time
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;  // timing marks of the pipeline
    cudaEvent_t ev_in[PIPE_CHUNKS] = {nullptr}, ev_out[PIPE_CHUNKS] = {nullptr};  // chunk landed / chunk priced
    float pipeline_ms = 0;
    std::map<GraphKey, cudaGraphExec_t> graphs;
    int threads = 0, blocks = 0;
    int tma_blocks = 0;  // grid of the TMA variant: SMs x CTAs that fit (one or two, 150 KB of shared memory each)
    // results of the last command
    int status = BS_GPU_OK;
    std::string err;
    float h2d_ms = 0, roi_ms = 0, d2h_ms = 0;
    unsigned long long err_total = 0;
    unsigned int list_n = 0;
    std::vector<long long> list;
    bool refval_on_device = false;
    std::thread worker;
};

}  // namespace

struct bs_gpu_ctx {
    size_t n = 0;
    int fp_bytes = 4;
    unsigned flags = 0;
    int math = BS_MATH_FAST;
    int cfg_threads = 0, cfg_blocks_per_sm = 0, unroll = 0, variant = 0;
    int tma_shape = 0;  // fp64 TMA kernel: 0 = 16 consumer warps / 4 stages; 1 = 24 / 3; 2 = 16 warps, two groups per thread, 2 stages
    std::vector<Shard> shards;
    void *host[BS_BUF_COUNT] = {nullptr};  // page-aligned anonymous mappings, pinned lazily by the device threads
    size_t host_bytes[BS_BUF_COUNT] = {0};
    bool host_registered[BS_BUF_COUNT] = {false};  // written by the device thread that owns buffer b (b % G)
    bool host_pinned = false;                      // CMD_PIN has run
    bool setup_pending = false;  // CMD_SETUP was posted by bs_gpu_init and has not been waited for yet
    int setup_status = BS_GPU_OK;
    // BS_GPU_FLAG_ASYNC_DISCOVERY: even device discovery (cuInit) runs in the background, on this thread
    std::thread bootstrap;
    int want_gpus = 0;
    std::vector<int> want_devices;
    bool inputs_dirty = true;   // host inputs newer than device copy
    bool device_valid = false;  // device inputs hold something meaningful
    // command mailbox
    std::mutex mu;
    std::condition_variable cv_cmd, cv_done;
    unsigned long long epoch = 0;
    int cmd = CMD_NONE;
    int pending = 0;
    // command arguments
    int arg_num_runs = 0, arg_err_chk = 0, arg_upload_what = 0;
    unsigned long long arg_first_index = 0;
    const void *arg_aos = nullptr;  // bs_gpu_price_aos: the caller's record array
    void *arg_out = nullptr;        // bs_gpu_price_aos: the caller's price array
    // reporting
    bs_gpu_timing timing;
    std::string err;
    bs_gpu_ctx() { memset(&timing, 0, sizeof(timing)); }
};

namespace {

size_t elem_bytes(const bs_gpu_ctx *c, int which) { return which == BS_BUF_OTYPE ? sizeof(int) : (size_t)c->fp_bytes; }
size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

void set_err(Shard &s, int status, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    s.status = status;
    s.err = buf;
}

#define SH_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_err(s, e_ == cudaErrorMemoryAllocation ? BS_GPU_ERR_NOMEM : BS_GPU_ERR_CUDA,            \
                    "device %d: %s failed: %s", s.device, #call, cudaGetErrorString(e_));               \
            return;                                                                                     \
        }                                                                                               \
    } while (0)

// ---- kernel dispatch -----------------------------------------------------------------------------
typedef void (*KernelF32)(bsk::StreamsF32, size_t, bsk::ErrChk);
typedef void (*KernelF64)(bsk::StreamsF64, size_t, bsk::ErrChk);

// bs_gpu_config.variant bits
enum {
    VARIANT_PIPE = 1,   // software-pipelined loads (next trip in flight during the math)
    VARIANT_PROBE = 2,  // DIAGNOSTIC ONLY: no pricing, same seven streams (sum of the inputs is written): the
                        // bandwidth ceiling of this traffic pattern.  Never selected by default.
    VARIANT_TMA = 4,    // inputs moved by cp.async.bulk into a shared-memory ring (bs_map_tma); ERR_CHK runs use
                        // the plain kernel
    VARIANT_FAULT = 8   // DIAGNOSTIC ONLY (fault injection for the tests): every Map launch asks for 1 MiB of dynamic
                        // shared memory, which the runtime rejects -- proves a failed launch surfaces as BS_GPU_ERR_CUDA
};

template <typename FP, int MATH, bool CHK, bool PIPE>
void (*pick_unroll(int unroll))(bsk::Streams<FP>, size_t, bsk::ErrChk)
{
    switch (unroll) {
    case 1: return bsk::bs_map<FP, MATH, 1, CHK, PIPE>;
    case 4: return bsk::bs_map<FP, MATH, 4, CHK, PIPE>;
    default: return bsk::bs_map<FP, MATH, 2, CHK, PIPE>;
    }
}
template <typename FP>
void (*pick_kernel(int math, int unroll, bool chk, bool pipe))(bsk::Streams<FP>, size_t, bsk::ErrChk)
{
    if (math == bsk::MATH_PROBE) return unroll == 1 ? bsk::bs_map<FP, bsk::MATH_PROBE, 1, false, false> : bsk::bs_map<FP, bsk::MATH_PROBE, 2, false, false>;
    if (math == BS_MATH_REFERENCE)  // validation mode: one geometry (one group per trip, plain loads)
        return chk ? bsk::bs_map<FP, bsk::MATH_REFERENCE, 1, true, false> : bsk::bs_map<FP, bsk::MATH_REFERENCE, 1, false, false>;
    if (math == BS_MATH_IEEE) {
        if (chk) return pipe ? pick_unroll<FP, bsk::MATH_IEEE, true, true>(unroll) : pick_unroll<FP, bsk::MATH_IEEE, true, false>(unroll);
        return pipe ? pick_unroll<FP, bsk::MATH_IEEE, false, true>(unroll) : pick_unroll<FP, bsk::MATH_IEEE, false, false>(unroll);
    }
    if (chk) return pipe ? pick_unroll<FP, bsk::MATH_FAST, true, true>(unroll) : pick_unroll<FP, bsk::MATH_FAST, true, false>(unroll);
    return pipe ? pick_unroll<FP, bsk::MATH_FAST, false, true>(unroll) : pick_unroll<FP, bsk::MATH_FAST, false, false>(unroll);
}
int kernel_math(const bs_gpu_ctx *c)
{
    if (c->variant & VARIANT_PROBE) return (int)bsk::MATH_PROBE;
    return c->math;
}
bool use_tma(const bs_gpu_ctx *c, bool chk) { return (c->variant & VARIANT_TMA) && !chk && c->math != BS_MATH_REFERENCE; }
template <typename FP, int SHAPE> const void *tma_kernel(int math)
{
    if (math == bsk::MATH_PROBE) return (const void *)bsk::bs_map_tma<FP, bsk::MATH_PROBE, SHAPE>;
    if (math == BS_MATH_IEEE) return (const void *)bsk::bs_map_tma<FP, bsk::MATH_IEEE, SHAPE>;
    return (const void *)bsk::bs_map_tma<FP, bsk::MATH_FAST, SHAPE>;
}
// shape of the TMA kernel: fp32 has one; fp64 has three (tma_shape: 16 warps / 4 stages, 24 / 3, 16 warps x two groups / 2)
const void *tma_kernel_ptr(const bs_gpu_ctx *c)
{
    if (c->fp_bytes == 4) return tma_kernel<float, 0>(kernel_math(c));
    return c->tma_shape == 2 ? tma_kernel<double, 2>(kernel_math(c)) : c->tma_shape == 1 ? tma_kernel<double, 1>(kernel_math(c)) : tma_kernel<double, 0>(kernel_math(c));
}
size_t tma_smem(const bs_gpu_ctx *c)
{
    if (c->fp_bytes == 4) return bsk::tma_smem_bytes<float, 0>();
    return c->tma_shape == 2 ? bsk::tma_smem_bytes<double, 2>() : c->tma_shape == 1 ? bsk::tma_smem_bytes<double, 1>() : bsk::tma_smem_bytes<double, 0>();
}
int tma_threads(const bs_gpu_ctx *c)
{
    if (c->fp_bytes == 4) return bsk::tma_threads<float, 0>();
    return c->tma_shape == 2 ? bsk::tma_threads<double, 2>() : c->tma_shape == 1 ? bsk::tma_threads<double, 1>() : bsk::tma_threads<double, 0>();
}
KernelF32 pick_f32(const bs_gpu_ctx *c, bool chk) { return pick_kernel<float>(kernel_math(c), c->unroll, chk, (c->variant & VARIANT_PIPE) != 0); }
KernelF64 pick_f64(const bs_gpu_ctx *c, bool chk) { return pick_kernel<double>(kernel_math(c), c->unroll, chk, (c->variant & VARIANT_PIPE) != 0); }
const void *kernel_ptr(const bs_gpu_ctx *c, bool chk)
{
    return c->fp_bytes == 4 ? (const void *)pick_f32(c, chk) : (const void *)pick_f64(c, chk);
}

// One launch of the Map kernel on the shard's stream.  With BS_GPU_FLAG_PDL the launch carries the
// programmatic-stream-serialization attribute: behind another Map launch it may begin (and issue its first loads)
// while that one drains; the kernel itself orders its stores after the predecessor (griddepcontrol.wait).  Behind
// a copy or memset the attribute has no effect.  Captured into the runs graph as a programmatic edge.
cudaError_t launch_kernel(bs_gpu_ctx *c, Shard &s, const void *fn, int blocks, void *streams, size_t *count, bsk::ErrChk *ec, bool allow_pdl,
                          int threads = 0, size_t smem = 0)
{
    if (c->variant & VARIANT_FAULT) smem = (size_t)1 << 20;
    void *args[3] = {streams, count, ec};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)(threads ? threads : s.threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (allow_pdl && (c->flags & BS_GPU_FLAG_PDL)) ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

// Launch the Map over options [first, first+count) of the shard (first must be a multiple of 4).
cudaError_t launch_map_range(bs_gpu_ctx *c, Shard &s, bool chk, int record, size_t first, size_t count, bool allow_pdl = false)
{
    if (count == 0) return cudaSuccess;
    bsk::ErrChk ec;
    ec.count = s.d_err_count;
    ec.list_count = s.d_list_count;
    ec.list = s.d_list;
    ec.list_cap = BS_GPU_MAX_ERROR_LIST;
    ec.record = record;
    ec.base = (long long)first;
    // whole-shard launches use the persistent grid chosen at setup; partial ones a proportional share of it
    int blocks = s.blocks;
    if (count != s.count) {
        const size_t per_group = c->fp_bytes == 4 ? 4 : 2;
        const size_t needed = (count / per_group + s.threads - 1) / s.threads;
        blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)s.blocks, needed));
    }
    if (c->fp_bytes == 4) {
        bsk::StreamsF32 a;
        a.spt = (const float *)s.d[BS_BUF_SPTPRICE] + first;
        a.strike = (const float *)s.d[BS_BUF_STRIKE] + first;
        a.rate = (const float *)s.d[BS_BUF_RATE] + first;
        a.vol = (const float *)s.d[BS_BUF_VOLATILITY] + first;
        a.otime = (const float *)s.d[BS_BUF_OTIME] + first;
        a.otype = (const int *)s.d[BS_BUF_OTYPE] + first;
        a.prices = (float *)s.d[BS_BUF_PRICES] + first;
        a.refval = s.d[BS_BUF_DGREFVAL] ? (const float *)s.d[BS_BUF_DGREFVAL] + first : nullptr;
        if (use_tma(c, chk)) return launch_kernel(c, s, tma_kernel_ptr(c), s.tma_blocks, &a, &count, &ec, allow_pdl, tma_threads(c), tma_smem(c));
        return launch_kernel(c, s, (const void *)pick_f32(c, chk), blocks, &a, &count, &ec, allow_pdl);
    } else {
        bsk::StreamsF64 a;
        a.spt = (const double *)s.d[BS_BUF_SPTPRICE] + first;
        a.strike = (const double *)s.d[BS_BUF_STRIKE] + first;
        a.rate = (const double *)s.d[BS_BUF_RATE] + first;
        a.vol = (const double *)s.d[BS_BUF_VOLATILITY] + first;
        a.otime = (const double *)s.d[BS_BUF_OTIME] + first;
        a.otype = (const int *)s.d[BS_BUF_OTYPE] + first;
        a.prices = (double *)s.d[BS_BUF_PRICES] + first;
        a.refval = s.d[BS_BUF_DGREFVAL] ? (const double *)s.d[BS_BUF_DGREFVAL] + first : nullptr;
        if (use_tma(c, chk)) return launch_kernel(c, s, tma_kernel_ptr(c), s.tma_blocks, &a, &count, &ec, allow_pdl, tma_threads(c), tma_smem(c));
        return launch_kernel(c, s, (const void *)pick_f64(c, chk), blocks, &a, &count, &ec, allow_pdl);
    }
}


// Pinning of the staging buffers (cudaHostRegister, portable).  Each buffer is registered WHOLE -- a copy whose host
// range straddles two separate registrations is rejected by the runtime -- and the eight buffers are dealt out over
// the device threads (buffer b goes to thread b % G) so the page pinning runs in parallel.  It happens once, as its
// own broadcast command right before the first copy, i.e. after the loader has filled (and faulted in) the pages.
// A failure is not fatal: copies from pageable memory still work, only slower.
void do_pin(bs_gpu_ctx *c, Shard &s)
{
    const int G = (int)c->shards.size();
    for (int b = s.index; b < BS_BUF_COUNT; b += G) {
        if (!c->host[b] || c->host_registered[b]) continue;
        const cudaError_t e = cudaHostRegister(c->host[b], c->host_bytes[b], cudaHostRegisterPortable);
        if (e == cudaSuccess) {
            c->host_registered[b] = true;
        } else {
            cudaGetLastError();
            s.err = std::string("cudaHostRegister: ") + cudaGetErrorString(e) + " (continuing with pageable copies)";
        }
    }
}

void do_unpin(bs_gpu_ctx *c, Shard &s)
{
    const int G = (int)c->shards.size();
    for (int b = s.index; b < BS_BUF_COUNT; b += G)
        if (c->host[b] && c->host_registered[b]) {
            if (cudaHostUnregister(c->host[b]) != cudaSuccess) cudaGetLastError();
            c->host_registered[b] = false;
        }
}

// ---- per-device commands (run on the device's own thread) ------------------------------------------
void do_setup(bs_gpu_ctx *c, Shard &s)
{
    SH_CUDA(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    SH_CUDA(cudaGetDeviceProperties(&prop, s.device));
    s.sm_count = prop.multiProcessorCount;
    SH_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    SH_CUDA(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    SH_CUDA(cudaStreamCreateWithFlags(&s.d2h_stream, cudaStreamNonBlocking));
    SH_CUDA(cudaEventCreate(&s.ev0));
    SH_CUDA(cudaEventCreate(&s.ev1));
    SH_CUDA(cudaEventCreate(&s.ev_h2d));
    SH_CUDA(cudaEventCreate(&s.ev_k0));
    SH_CUDA(cudaEventCreate(&s.ev_k1));
    for (int i = 0; i < PIPE_CHUNKS; i++) {
        SH_CUDA(cudaEventCreateWithFlags(&s.ev_in[i], cudaEventDisableTiming));
        SH_CUDA(cudaEventCreateWithFlags(&s.ev_out[i], cudaEventDisableTiming));
    }

    // arena: eight padded streams (DGrefval only when requested)
    size_t off[BS_BUF_COUNT], total = 0;
    for (int b = 0; b < BS_BUF_COUNT; b++) {
        off[b] = total;
        if (b == BS_BUF_DGREFVAL && !(c->flags & BS_GPU_FLAG_WITH_DGREFVAL)) continue;
        total += pad256(std::max<size_t>(s.count, 1) * elem_bytes(c, b));
    }
    s.arena_bytes = total;
    SH_CUDA(cudaMalloc((void **)&s.arena, total));
    for (int b = 0; b < BS_BUF_COUNT; b++) {
        if (b == BS_BUF_DGREFVAL && !(c->flags & BS_GPU_FLAG_WITH_DGREFVAL)) { s.d[b] = nullptr; continue; }
        s.d[b] = s.arena + off[b];
    }
    SH_CUDA(cudaMalloc((void **)&s.d_err_count, sizeof(unsigned long long)));
    SH_CUDA(cudaMalloc((void **)&s.d_list_count, sizeof(unsigned int)));
    SH_CUDA(cudaMalloc((void **)&s.d_list, sizeof(long long) * BS_GPU_MAX_ERROR_LIST));
    SH_CUDA(cudaMemsetAsync(s.d_err_count, 0, sizeof(unsigned long long), s.stream));
    SH_CUDA(cudaMemsetAsync(s.d_list_count, 0, sizeof(unsigned int), s.stream));

    // launch geometry: persistent grid = SMs x resident CTAs, clipped to the work available
    const size_t per_group = c->fp_bytes == 4 ? 4 : 2;
    const size_t groups = std::max<size_t>(s.count / per_group, 1);
    int threads = c->cfg_threads ? c->cfg_threads : 256;
    if (!c->cfg_threads && groups < (size_t)s.sm_count * 256) threads = 64;  // tiny inputs: spread over SMs
    int resident = 0;
    SH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel_ptr(c, false), threads, 0));
    if (resident < 1) resident = 1;
    if (c->cfg_blocks_per_sm > 0) resident = c->cfg_blocks_per_sm;
    size_t blocks = (size_t)s.sm_count * resident;
    const size_t needed = (groups + threads - 1) / threads;
    if (blocks > needed) blocks = needed;
    if (blocks < 1) blocks = 1;
    s.threads = threads;
    s.blocks = (int)blocks;
    if (c->variant & VARIANT_TMA) {
        SH_CUDA(cudaFuncSetAttribute(tma_kernel_ptr(c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem(c)));
        int fit = 0;
        SH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, tma_kernel_ptr(c), tma_threads(c), tma_smem(c)));
        if (fit < 1) fit = 1;
        if (c->cfg_blocks_per_sm > 0) fit = std::min(fit, c->cfg_blocks_per_sm);
        s.tma_blocks = s.sm_count * fit;
    }
    SH_CUDA(cudaStreamSynchronize(s.stream));
}

enum { UP_INPUTS = 1, UP_REFVAL = 2 };

void do_upload(bs_gpu_ctx *c, Shard &s, int what)
{
    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    if (what & UP_INPUTS)
        for (int b = BS_BUF_SPTPRICE; b <= BS_BUF_OTYPE; b++) {
            const size_t eb = elem_bytes(c, b);
            SH_CUDA(cudaMemcpyAsync(s.d[b], (const char *)c->host[b] + s.first * eb, s.count * eb, cudaMemcpyHostToDevice, s.stream));
        }
    if ((what & UP_REFVAL) && s.d[BS_BUF_DGREFVAL]) {
        const size_t eb = elem_bytes(c, BS_BUF_DGREFVAL);
        SH_CUDA(cudaMemcpyAsync(s.d[BS_BUF_DGREFVAL], (const char *)c->host[BS_BUF_DGREFVAL] + s.first * eb, s.count * eb,
                                cudaMemcpyHostToDevice, s.stream));
        s.refval_on_device = true;
    }
    SH_CUDA(cudaEventRecord(s.ev1, s.stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.h2d_ms, s.ev0, s.ev1));
}

// NUM_RUNS real launches (blackscholes.c:318): every run re-reads all inputs and rewrites all prices
cudaError_t enqueue_runs(bs_gpu_ctx *c, Shard &s, int num_runs, bool chk, bool record_last, size_t first, size_t count)
{
    for (int j = 0; j < num_runs; j++) {
        const cudaError_t e = launch_map_range(c, s, chk, chk && record_last && j == num_runs - 1, first, count, true);
        if (e != cudaSuccess) return e;  // the first failed launch ends the sequence: nothing after it would be valid
    }
    return cudaSuccess;
}
cudaError_t enqueue_runs(bs_gpu_ctx *c, Shard &s, int num_runs, bool chk, bool record_last) { return enqueue_runs(c, s, num_runs, chk, record_last, 0, s.count); }

cudaError_t reset_err_counters(Shard &s)
{
    const cudaError_t e = cudaMemsetAsync(s.d_err_count, 0, sizeof(unsigned long long), s.stream);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(s.d_list_count, 0, sizeof(unsigned int), s.stream);
}

// `num_runs` launches over [first, first+count) as one cached CUDA graph (or nullptr when graphs are disabled)
cudaGraphExec_t runs_graph(bs_gpu_ctx *c, Shard &s, int num_runs, bool chk, bool record_last, size_t first, size_t count)
{
    if ((c->flags & BS_GPU_FLAG_NO_GRAPH) || num_runs <= 0) return nullptr;
    GraphKey key = {num_runs, chk ? (record_last ? 1 : 2) : 0, first, count};
    auto it = s.graphs.find(key);
    if (it != s.graphs.end()) return it->second;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    // Any failure here returns nullptr and the caller launches the runs directly, where the same failure (if it is
    // one of the launch, not of the capture machinery) is reported with its CUDA error text.
    if (cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const cudaError_t queued = enqueue_runs(c, s, num_runs, chk, record_last, first, count);
    const cudaError_t ended = cudaStreamEndCapture(s.stream, &graph);
    if (queued != cudaSuccess || ended != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return nullptr;
    }
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); cudaGetLastError(); return nullptr; }
    cudaGraphDestroy(graph);
    if (cudaGraphUpload(exec, s.stream) != cudaSuccess) cudaGetLastError();  // an optimisation only: the first launch uploads otherwise
    s.graphs[key] = exec;
    return exec;
}
cudaGraphExec_t runs_graph(bs_gpu_ctx *c, Shard &s, int num_runs, bool chk, bool record_last)
{
    return runs_graph(c, s, num_runs, chk, record_last, 0, s.count);
}

void fetch_err_results(bs_gpu_ctx *c, Shard &s, bool chk);

void do_run(bs_gpu_ctx *c, Shard &s)
{
    const int num_runs = c->arg_num_runs;
    const bool chk = c->arg_err_chk != 0;
    if (s.count == 0) { s.roi_ms = 0; s.err_total = 0; s.list_n = 0; s.list.clear(); return; }
    cudaGraphExec_t exec = runs_graph(c, s, num_runs, chk, true);
    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    if (chk) SH_CUDA(reset_err_counters(s));
    if (exec) {
        SH_CUDA(cudaGraphLaunch(exec, s.stream));
    } else {
        SH_CUDA(enqueue_runs(c, s, num_runs, chk, true));
        SH_CUDA(cudaGetLastError());
    }
    SH_CUDA(cudaEventRecord(s.ev1, s.stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.roi_ms, s.ev0, s.ev1));
    fetch_err_results(c, s, chk);
}

void fetch_err_results(bs_gpu_ctx *c, Shard &s, bool chk)
{
    (void)c;
    if (chk) {
        SH_CUDA(cudaMemcpyAsync(&s.err_total, s.d_err_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        SH_CUDA(cudaMemcpyAsync(&s.list_n, s.d_list_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
        SH_CUDA(cudaStreamSynchronize(s.stream));
        const unsigned int keep = std::min<unsigned int>(s.list_n, BS_GPU_MAX_ERROR_LIST);
        s.list.resize(keep);
        if (keep) {
            SH_CUDA(cudaMemcpyAsync(s.list.data(), s.d_list, keep * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
            SH_CUDA(cudaStreamSynchronize(s.stream));
            std::sort(s.list.begin(), s.list.end());
        }
    } else {
        s.err_total = 0;
        s.list_n = 0;
        s.list.clear();
    }
}

// bs_gpu_price() on one device: copies and launches pipelined IN RUN ORDER.
//   copy stream   : H2D chunk 0..7 ........................................ D2H chunk 0..7 (as each is priced)
//   launch stream : run 0 chunk 0..7 (each after its inputs landed) | runs 1..R-2 whole-shard | run R-1 chunk 0..7
// Every run still reads every input from HBM and writes every price; only the first and the last run are
// cut into PIPE_CHUNKS launches so that PCIe traffic hides behind them.
// Number of sub-shards a bs_gpu_price() call cuts this shard into.  A sub-shard's streams must not fit the L2
// (126 MB): every run over it then still streams every input from HBM -- ncu on a 5M-option (140 MB) sub-shard:
// dram__bytes_read = 120.0 MB = its algorithmic reads, L2 hit rate 0.04 % (profiles/r01_half_shard_probe.txt) -- so
// the split changes the SCHEDULE of the NUM_RUNS x options evaluations, never what a run reads.
int price_subshards(const bs_gpu_ctx *c, const Shard &s)
{
    const size_t bytes = s.count * (6 * (size_t)c->fp_bytes + 4);  // six fptype streams (5 in, 1 out) + otype
    const size_t min_bytes = (size_t)128 << 20;
    return (int)std::max<size_t>(1, std::min<size_t>(PIPE_CHUNKS, bytes / min_bytes));
}

// bs_gpu_price() on one device, sub-shard scheme (shards of >= 256 MiB): the shard is cut into S contiguous
// sub-shards, each larger than the L2, and sub-shard i runs ALL its NUM_RUNS launches as soon as its inputs have
// landed, while sub-shard i+1 is still on the PCIe bus and the prices of sub-shard i-1 travel back:
//   H2D engine    : in 0 | in 1 | in 2 ...
//   launch stream :        R runs on 0 | R runs on 1 | ...
//   D2H engine    :                      out 0       | out 1 ...
// Independent options make this legal: it is the time-multiplexed twin of sharding over several GPUs, where every
// device also runs all NUM_RUNS over its own range without waiting for the others.
void do_price_subshards(bs_gpu_ctx *c, Shard &s, int S)
{
    const int R = c->arg_num_runs;
    const bool chk = c->arg_err_chk != 0;
    const int what = c->arg_upload_what;
    size_t lo[PIPE_CHUNKS + 1];
    const size_t per = ((s.count + S - 1) / S + 1023) & ~(size_t)1023;
    for (int k = 0; k <= S; k++) lo[k] = std::min(s.count, per * (size_t)k);
    lo[S] = s.count;
    const size_t pe = elem_bytes(c, BS_BUF_PRICES);

    // graphs first (instantiation must not sit between the timing marks)
    cudaGraphExec_t exec[PIPE_CHUNKS];
    for (int k = 0; k < S; k++) exec[k] = runs_graph(c, s, R, chk, true, lo[k], lo[k + 1] - lo[k]);

    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    SH_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev0, 0));
    SH_CUDA(cudaStreamWaitEvent(s.d2h_stream, s.ev0, 0));
    if (chk) SH_CUDA(reset_err_counters(s));
    if (what) {
        for (int k = 0; k < S; k++) {
            const size_t n = lo[k + 1] - lo[k];
            if (what & UP_INPUTS)
                for (int b = BS_BUF_SPTPRICE; b <= BS_BUF_OTYPE; b++) {
                    const size_t eb = elem_bytes(c, b);
                    SH_CUDA(cudaMemcpyAsync((char *)s.d[b] + lo[k] * eb, (const char *)c->host[b] + (s.first + lo[k]) * eb, n * eb,
                                            cudaMemcpyHostToDevice, s.copy_stream));
                }
            if ((what & UP_REFVAL) && s.d[BS_BUF_DGREFVAL]) {
                const size_t eb = elem_bytes(c, BS_BUF_DGREFVAL);
                SH_CUDA(cudaMemcpyAsync((char *)s.d[BS_BUF_DGREFVAL] + lo[k] * eb, (const char *)c->host[BS_BUF_DGREFVAL] + (s.first + lo[k]) * eb,
                                        n * eb, cudaMemcpyHostToDevice, s.copy_stream));
            }
            SH_CUDA(cudaEventRecord(s.ev_in[k], s.copy_stream));
        }
        if ((what & UP_REFVAL) && s.d[BS_BUF_DGREFVAL]) s.refval_on_device = true;
    }
    SH_CUDA(cudaEventRecord(s.ev_h2d, s.copy_stream));
    for (int k = 0; k < S; k++) {
        if (what) SH_CUDA(cudaStreamWaitEvent(s.stream, s.ev_in[k], 0));
        if (k == 0) SH_CUDA(cudaEventRecord(s.ev_k0, s.stream));
        if (exec[k]) SH_CUDA(cudaGraphLaunch(exec[k], s.stream));
        else SH_CUDA(enqueue_runs(c, s, R, chk, true, lo[k], lo[k + 1] - lo[k]));
        SH_CUDA(cudaEventRecord(s.ev_out[k], s.stream));
    }
    SH_CUDA(cudaGetLastError());
    SH_CUDA(cudaEventRecord(s.ev_k1, s.stream));
    for (int k = 0; k < S; k++) {
        const size_t n = lo[k + 1] - lo[k];
        SH_CUDA(cudaStreamWaitEvent(s.d2h_stream, s.ev_out[k], 0));
        if (n)
            SH_CUDA(cudaMemcpyAsync((char *)c->host[BS_BUF_PRICES] + (s.first + lo[k]) * pe, (const char *)s.d[BS_BUF_PRICES] + lo[k] * pe, n * pe,
                                    cudaMemcpyDeviceToHost, s.d2h_stream));
    }
    SH_CUDA(cudaEventRecord(s.ev1, s.d2h_stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaStreamSynchronize(s.stream));
    SH_CUDA(cudaStreamSynchronize(s.copy_stream));
    SH_CUDA(cudaEventElapsedTime(&s.pipeline_ms, s.ev0, s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.h2d_ms, s.ev0, s.ev_h2d));
    SH_CUDA(cudaEventElapsedTime(&s.roi_ms, s.ev_k0, s.ev_k1));
    SH_CUDA(cudaEventElapsedTime(&s.d2h_ms, s.ev_k1, s.ev1));
    fetch_err_results(c, s, chk);
}

void do_price(bs_gpu_ctx *c, Shard &s)
{
    const int R = c->arg_num_runs;
    const bool chk = c->arg_err_chk != 0;
    const int what = c->arg_upload_what;
    s.h2d_ms = s.roi_ms = s.d2h_ms = s.pipeline_ms = 0;
    if (s.count == 0) { s.err_total = 0; s.list_n = 0; s.list.clear(); return; }
    {
        const int S = price_subshards(c, s);
        if (S > 1 && R >= 1 && !(c->flags & BS_GPU_FLAG_NO_SUBSHARDS)) { do_price_subshards(c, s, S); return; }
    }

    // chunk boundaries: multiples of 1024 options keep every stream 16-byte aligned.  Shards whose copies are bound by
    // latency, not by bandwidth (under 4 MiB of streams: simsmall is 112 KB), go in ONE chunk: eight chunks of six
    // streams each would be 48 copies of a few KB and a dozen extra launches around 100 launches of a microsecond each.
    const int NCH = s.count * (5 * elem_bytes(c, BS_BUF_SPTPRICE) + 4 + elem_bytes(c, BS_BUF_PRICES)) < ((size_t)4 << 20) ? 1 : PIPE_CHUNKS;
    size_t lo[PIPE_CHUNKS + 1];
    const size_t per = ((s.count + NCH - 1) / NCH + 1023) & ~(size_t)1023;
    for (int k = 0; k <= PIPE_CHUNKS; k++) lo[k] = std::min(s.count, per * (size_t)k);
    const size_t pe = elem_bytes(c, BS_BUF_PRICES);

    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    SH_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev0, 0));
    if (chk) SH_CUDA(reset_err_counters(s));

    // ---- inputs: chunked H2D on the copy stream
    if (what) {
        for (int k = 0; k < NCH; k++) {
            const size_t n = lo[k + 1] - lo[k];
            if (n) {
                if (what & UP_INPUTS)
                    for (int b = BS_BUF_SPTPRICE; b <= BS_BUF_OTYPE; b++) {
                        const size_t eb = elem_bytes(c, b);
                        SH_CUDA(cudaMemcpyAsync((char *)s.d[b] + lo[k] * eb, (const char *)c->host[b] + (s.first + lo[k]) * eb, n * eb,
                                                cudaMemcpyHostToDevice, s.copy_stream));
                    }
                if ((what & UP_REFVAL) && s.d[BS_BUF_DGREFVAL]) {
                    const size_t eb = elem_bytes(c, BS_BUF_DGREFVAL);
                    SH_CUDA(cudaMemcpyAsync((char *)s.d[BS_BUF_DGREFVAL] + lo[k] * eb, (const char *)c->host[BS_BUF_DGREFVAL] + (s.first + lo[k]) * eb,
                                            n * eb, cudaMemcpyHostToDevice, s.copy_stream));
                }
            }
            SH_CUDA(cudaEventRecord(s.ev_in[k], s.copy_stream));
        }
        if ((what & UP_REFVAL) && s.d[BS_BUF_DGREFVAL]) s.refval_on_device = true;
    }
    SH_CUDA(cudaEventRecord(s.ev_h2d, s.copy_stream));

    // ---- runs
    int done = 0;
    bool k0_marked = false;
    auto mark_k0 = [&]() -> cudaError_t {
        if (k0_marked) return cudaSuccess;
        k0_marked = true;
        return cudaEventRecord(s.ev_k0, s.stream);
    };
    if (R >= 1 && what) {  // run 0 follows the input chunks (it is also the last run when R == 1)
        const bool last = (R == 1);
        for (int k = 0; k < NCH; k++) {
            SH_CUDA(cudaStreamWaitEvent(s.stream, s.ev_in[k], 0));
            SH_CUDA(mark_k0());
            SH_CUDA(launch_map_range(c, s, chk, chk && last, lo[k], lo[k + 1] - lo[k]));
            if (last) SH_CUDA(cudaEventRecord(s.ev_out[k], s.stream));
        }
        done = 1;
    } else if (what) {
        SH_CUDA(cudaStreamWaitEvent(s.stream, s.ev_h2d, 0));  // R == 0: nothing to overlap with
    }
    SH_CUDA(mark_k0());
    const int middle = std::max(0, R - done - 1);  // whole-shard runs between the pipelined first and last
    if (middle > 0) {
        cudaGraphExec_t exec = runs_graph(c, s, middle, chk, false);
        if (exec) SH_CUDA(cudaGraphLaunch(exec, s.stream));
        else SH_CUDA(enqueue_runs(c, s, middle, chk, false));
        done += middle;
    }
    bool out_marked = (R == 1 && what);
    if (done < R) {  // the last run, chunk by chunk, each chunk's prices leaving as soon as they exist
        for (int k = 0; k < NCH; k++) {
            SH_CUDA(launch_map_range(c, s, chk, chk, lo[k], lo[k + 1] - lo[k]));
            SH_CUDA(cudaEventRecord(s.ev_out[k], s.stream));
        }
        out_marked = true;
    }
    SH_CUDA(cudaGetLastError());
    SH_CUDA(cudaEventRecord(s.ev_k1, s.stream));

    // ---- prices: chunked D2H on the copy stream
    for (int k = 0; k < NCH; k++) {
        const size_t n = lo[k + 1] - lo[k];
        if (out_marked) SH_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev_out[k], 0));
        else if (k == 0) SH_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev_k1, 0));
        if (n)
            SH_CUDA(cudaMemcpyAsync((char *)c->host[BS_BUF_PRICES] + (s.first + lo[k]) * pe, (const char *)s.d[BS_BUF_PRICES] + lo[k] * pe, n * pe,
                                    cudaMemcpyDeviceToHost, s.copy_stream));
    }
    SH_CUDA(cudaEventRecord(s.ev1, s.copy_stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaStreamSynchronize(s.stream));
    SH_CUDA(cudaEventElapsedTime(&s.pipeline_ms, s.ev0, s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.h2d_ms, s.ev0, s.ev_h2d));
    SH_CUDA(cudaEventElapsedTime(&s.roi_ms, s.ev_k0, s.ev_k1));
    SH_CUDA(cudaEventElapsedTime(&s.d2h_ms, s.ev_k1, s.ev1));
    fetch_err_results(c, s, chk);
}

void do_download(bs_gpu_ctx *c, Shard &s)
{
    const size_t eb = elem_bytes(c, BS_BUF_PRICES);
    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    SH_CUDA(cudaMemcpyAsync((char *)c->host[BS_BUF_PRICES] + s.first * eb, s.d[BS_BUF_PRICES], s.count * eb, cudaMemcpyDeviceToHost, s.stream));
    SH_CUDA(cudaEventRecord(s.ev1, s.stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.d2h_ms, s.ev0, s.ev1));
}

// bs_gpu_price_aos, input half: the shard's slice of the caller's DataCont records (24 bytes each, pageable memory)
// goes H2D as it is and is turned into the SoA streams on the device (bsk::bs_aos_to_soa).
void do_aos_in(bs_gpu_ctx *c, Shard &s)
{
    s.h2d_ms = 0;
    if (s.count == 0) return;
    const size_t rec = sizeof(int) + 5 * sizeof(float);
    const size_t tiles = (s.count + bsk::AOS_TILE - 1) / bsk::AOS_TILE;
    const size_t need = tiles * bsk::AOS_TILE * rec;  // whole tiles: the gather reads 16-byte vectors up to the tile end
    if (s.d_aos_bytes < need) {
        if (s.d_aos) SH_CUDA(cudaFree(s.d_aos));
        s.d_aos = nullptr;
        s.d_aos_bytes = 0;
        SH_CUDA(cudaMalloc((void **)&s.d_aos, need));
        s.d_aos_bytes = need;
    }
    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    SH_CUDA(cudaMemcpyAsync(s.d_aos, (const char *)c->arg_aos + s.first * rec, s.count * rec, cudaMemcpyHostToDevice, s.stream));
    const int blocks = (int)std::min<size_t>(tiles, (size_t)s.sm_count * 8);
    bsk::bs_aos_to_soa<<<blocks, bsk::AOS_TILE, 0, s.stream>>>((const uint4 *)s.d_aos, s.count, (float *)s.d[BS_BUF_SPTPRICE],
                                                              (float *)s.d[BS_BUF_STRIKE], (float *)s.d[BS_BUF_RATE],
                                                              (float *)s.d[BS_BUF_VOLATILITY], (float *)s.d[BS_BUF_OTIME],
                                                              (int *)s.d[BS_BUF_OTYPE]);
    SH_CUDA(cudaGetLastError());
    SH_CUDA(cudaEventRecord(s.ev1, s.stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.h2d_ms, s.ev0, s.ev1));
    s.refval_on_device = false;
}

// bs_gpu_price_aos, output half: the shard's prices into the caller's (pageable) array.
void do_download_to(bs_gpu_ctx *c, Shard &s)
{
    s.d2h_ms = 0;
    if (s.count == 0) return;
    const size_t eb = elem_bytes(c, BS_BUF_PRICES);
    SH_CUDA(cudaEventRecord(s.ev0, s.stream));
    SH_CUDA(cudaMemcpyAsync((char *)c->arg_out + s.first * eb, s.d[BS_BUF_PRICES], s.count * eb, cudaMemcpyDeviceToHost, s.stream));
    SH_CUDA(cudaEventRecord(s.ev1, s.stream));
    SH_CUDA(cudaEventSynchronize(s.ev1));
    SH_CUDA(cudaEventElapsedTime(&s.d2h_ms, s.ev0, s.ev1));
}

// Table values exactly as a reader of the inputgen text would obtain them (text round trip), so that a
// device-filled set is bit-identical to loading the equivalent file.
template <typename FP> FP through_text(double v, const char *fmt)
{
    char buf[64];
    snprintf(buf, sizeof(buf), fmt, v);
    return sizeof(FP) == 4 ? (FP)strtof(buf, nullptr) : (FP)strtod(buf, nullptr);
}

template <typename FP> void do_fill_typed(bs_gpu_ctx *c, Shard &s)
{
    const int R = BS_TABLE_ROWS;
    const size_t fp_block = pad256(sizeof(FP) * R);
    const size_t table_bytes = 6 * fp_block + pad256(sizeof(int) * R);
    if (!s.d_table) {
        std::vector<char> h(table_bytes, 0);
        FP *hs = (FP *)(h.data() + 0 * fp_block), *hk = (FP *)(h.data() + 1 * fp_block), *hr = (FP *)(h.data() + 2 * fp_block);
        FP *hv = (FP *)(h.data() + 3 * fp_block), *ht = (FP *)(h.data() + 4 * fp_block), *hd = (FP *)(h.data() + 5 * fp_block);
        int *ho = (int *)(h.data() + 6 * fp_block);
        for (int i = 0; i < R; i++) {
            const bs_table_row &row = bs_option_table[i];
            hs[i] = through_text<FP>(row.s, "%.2f");
            hk[i] = through_text<FP>(row.strike, "%.2f");
            hr[i] = through_text<FP>(row.r, "%.4f");
            hv[i] = through_text<FP>(row.v, "%.2f");
            ht[i] = through_text<FP>(row.t, "%.2f");
            hd[i] = through_text<FP>(row.dgrefval, "%.18f");
            ho[i] = (row.option_type == 'P') ? 1 : 0;  // blackscholes.c:761
        }
        SH_CUDA(cudaMalloc((void **)&s.d_table, table_bytes));
        SH_CUDA(cudaMemcpyAsync(s.d_table, h.data(), table_bytes, cudaMemcpyHostToDevice, s.stream));
        SH_CUDA(cudaStreamSynchronize(s.stream));
    }
    bsk::TableDev<FP> tab;
    tab.spt = (const FP *)(s.d_table + 0 * fp_block);
    tab.strike = (const FP *)(s.d_table + 1 * fp_block);
    tab.rate = (const FP *)(s.d_table + 2 * fp_block);
    tab.vol = (const FP *)(s.d_table + 3 * fp_block);
    tab.otime = (const FP *)(s.d_table + 4 * fp_block);
    tab.refval = (const FP *)(s.d_table + 5 * fp_block);
    tab.otype = (const int *)(s.d_table + 6 * fp_block);
    tab.rows = R;
    if (s.count) {
        const int blocks = (int)std::min<size_t>((s.count + 255) / 256, (size_t)s.sm_count * 8);
        bsk::bs_fill_synthetic<FP><<<blocks, 256, 0, s.stream>>>(
            tab, c->arg_first_index + s.first, s.count, (FP *)s.d[BS_BUF_SPTPRICE], (FP *)s.d[BS_BUF_STRIKE],
            (FP *)s.d[BS_BUF_RATE], (FP *)s.d[BS_BUF_VOLATILITY], (FP *)s.d[BS_BUF_OTIME], (int *)s.d[BS_BUF_OTYPE],
            (FP *)s.d[BS_BUF_DGREFVAL]);
        SH_CUDA(cudaGetLastError());
    }
    SH_CUDA(cudaStreamSynchronize(s.stream));
    s.refval_on_device = s.d[BS_BUF_DGREFVAL] != nullptr;
}

void do_teardown(bs_gpu_ctx *c, Shard &s)
{
    (void)c;
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.copy_stream) cudaStreamSynchronize(s.copy_stream);
    if (s.d2h_stream) cudaStreamSynchronize(s.d2h_stream);
    do_unpin(c, s);
    for (auto &kv : s.graphs) cudaGraphExecDestroy(kv.second);
    s.graphs.clear();
    if (s.d_table) cudaFree(s.d_table);
    if (s.d_aos) cudaFree(s.d_aos);
    s.d_aos = nullptr;
    s.d_aos_bytes = 0;
    if (s.d_list) cudaFree(s.d_list);
    if (s.d_list_count) cudaFree(s.d_list_count);
    if (s.d_err_count) cudaFree(s.d_err_count);
    if (s.arena) cudaFree(s.arena);
    for (cudaEvent_t e : {s.ev0, s.ev1, s.ev_h2d, s.ev_k0, s.ev_k1})
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < PIPE_CHUNKS; i++) {
        if (s.ev_in[i]) cudaEventDestroy(s.ev_in[i]);
        if (s.ev_out[i]) cudaEventDestroy(s.ev_out[i]);
    }
    if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
    if (s.d2h_stream) cudaStreamDestroy(s.d2h_stream);
    if (s.stream) cudaStreamDestroy(s.stream);
    s.copy_stream = s.d2h_stream = nullptr;
    s.d_table = nullptr; s.d_list = nullptr; s.d_list_count = nullptr; s.d_err_count = nullptr;
    s.arena = nullptr; s.ev0 = s.ev1 = nullptr; s.stream = nullptr;
}

void device_thread(bs_gpu_ctx *c, int g)
{
    Shard &s = c->shards[g];
    unsigned long long seen = 0;
    for (;;) {
        int cmd;
        {
            std::unique_lock<std::mutex> lk(c->mu);
            c->cv_cmd.wait(lk, [&] { return c->epoch != seen; });
            seen = c->epoch;
            cmd = c->cmd;
        }
        s.status = BS_GPU_OK;
        s.err.clear();
        switch (cmd) {
        case CMD_SETUP: do_setup(c, s); break;
        case CMD_PIN: do_pin(c, s); break;
        case CMD_UPLOAD: do_upload(c, s, c->arg_upload_what); break;
        case CMD_RUN: do_run(c, s); break;
        case CMD_DOWNLOAD: do_download(c, s); break;
        case CMD_PRICE: do_price(c, s); break;
        case CMD_FILL:
            if (c->fp_bytes == 4) do_fill_typed<float>(c, s); else do_fill_typed<double>(c, s);
            break;
        case CMD_AOS_IN: do_aos_in(c, s); break;
        case CMD_DOWNLOAD_TO: do_download_to(c, s); break;
        case CMD_TEARDOWN: do_teardown(c, s); break;
        default: break;
        }
        {
            std::lock_guard<std::mutex> lk(c->mu);
            if (--c->pending == 0) c->cv_done.notify_all();
        }
        if (cmd == CMD_EXIT) return;
    }
}

// Post one command to every device thread (no waiting).
void post(bs_gpu_ctx *c, int cmd)
{
    std::unique_lock<std::mutex> lk(c->mu);
    c->cmd = cmd;
    c->pending = (int)c->shards.size();
    c->epoch++;
    c->cv_cmd.notify_all();
}

// Wait until every device thread has finished the posted command.  Returns the first failure.
int wait_all(bs_gpu_ctx *c)
{
    {
        std::unique_lock<std::mutex> lk(c->mu);
        c->cv_done.wait(lk, [&] { return c->pending == 0; });
    }
    for (auto &s : c->shards)
        if (s.status != BS_GPU_OK) {
            c->err = s.err;
            return s.status;
        }
    return BS_GPU_OK;
}

void device_thread(bs_gpu_ctx *c, int g);

// Lay the shards over `G` devices (contiguous; the first N % G shards take one extra option -- the ff static
// partitioner rule), start one thread per device and post CMD_SETUP.
void start_devices(bs_gpu_ctx *c, int G)
{
    c->shards.resize(G);
    const size_t q = c->n / G, r = c->n % G;
    size_t first = 0;
    for (int g = 0; g < G; g++) {
        Shard &s = c->shards[g];
        s.index = g;
        s.device = c->want_devices.empty() ? g : c->want_devices[g];
        s.first = first;
        s.count = q + ((size_t)g < r ? 1 : 0);
        first += s.count;
    }
    for (int g = 0; g < G; g++) c->shards[g].worker = std::thread(device_thread, c, g);
    post(c, CMD_SETUP);
}

int count_devices()
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? 0 : BS_GPU_ERR_CUDA;
    }
    return n;
}

// Background half of an ASYNC_DISCOVERY init: cuInit, clamp the GPU count to what exists, bring the devices up.
void bootstrap_thread(bs_gpu_ctx *c)
{
    const int have = count_devices();
    if (have <= 0) {
        c->setup_status = have < 0 ? have : BS_GPU_ERR_NO_DEVICE;
        c->err = "no usable CUDA device";
        return;
    }
    int G = std::min(c->want_gpus, have);
    for (int g = 0; g < (int)c->want_devices.size() && g < G; g++)
        if (c->want_devices[g] < 0 || c->want_devices[g] >= have) {
            c->setup_status = BS_GPU_ERR_NO_DEVICE;
            c->err = "device ordinal out of range";
            return;
        }
    if ((size_t)G > std::max<size_t>(c->n, 1)) G = (int)std::max<size_t>(c->n, 1);
    start_devices(c, G);
    c->setup_status = wait_all(c);
}

// bs_gpu_init posts CMD_SETUP and returns: the device contexts come up in the background while the caller
// (the loader) fills the staging buffers.  The first later call collects the outcome; a failed setup is sticky.
int finish_setup(bs_gpu_ctx *c)
{
    if (c->bootstrap.joinable()) {
        c->bootstrap.join();  // sets setup_status
        c->setup_pending = false;
    }
    if (c->setup_pending) {
        c->setup_status = wait_all(c);
        c->setup_pending = false;
    }
    return c->setup_status;
}

// Post one command to every device thread and wait for all of them.  Returns the first failure.
int broadcast(bs_gpu_ctx *c, int cmd)
{
    if (cmd != CMD_TEARDOWN && cmd != CMD_EXIT) {
        const int st = finish_setup(c);
        if (st != BS_GPU_OK) return st;
    } else if (c->setup_pending) {
        wait_all(c);
        c->setup_pending = false;
    }
    post(c, cmd);
    return wait_all(c);
}

// Pin the staging buffers once, right before the first command that copies from / to them.
int pin_staging(bs_gpu_ctx *c)
{
    if (c->host_pinned || (c->flags & BS_GPU_FLAG_NO_HOST_STAGING)) return BS_GPU_OK;
    const int st = broadcast(c, CMD_PIN);
    if (st == BS_GPU_OK) c->host_pinned = true;
    return st;
}

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(bs_gpu_ctx *c, int status, const char *msg)
{
    if (c) c->err = msg;
    return status;
}

}  // namespace

// ====================================================================================================
// C ABI
// ====================================================================================================
extern "C" {

int bs_gpu_abi_version(void) { return BS_GPU_ABI_VERSION; }

const char *bs_gpu_status_string(int status)
{
    switch (status) {
    case BS_GPU_OK: return "ok";
    case BS_GPU_ERR_INVALID: return "invalid argument";
    case BS_GPU_ERR_NO_DEVICE: return "no usable CUDA device";
    case BS_GPU_ERR_CUDA: return "CUDA call failed";
    case BS_GPU_ERR_NOMEM: return "out of memory";
    case BS_GPU_ERR_STATE: return "invalid state for this call";
    default: return "unknown status";
    }
}

int bs_gpu_device_count(void) { return count_devices(); }

/* Must run before the process's first CUDA call to have an effect: the driver initialises every VISIBLE device in
 * cuInit -- 4.4-5.4 s on an 8-GPU B200 box against 0.36-0.49 s with one device visible (profiles/r02_cuinit_8gpu.jsonl) --
 * and tears all of them down again at process exit. */
int bs_gpu_limit_devices(int max_gpus)
{
    if (max_gpus < 1) return BS_GPU_ERR_INVALID;
    std::string list;
    const char *vis = getenv("CUDA_VISIBLE_DEVICES");
    if (vis && !*vis) return BS_GPU_OK;  // set and empty: the caller hid every device; leave it so
    if (vis) {  // keep the first max_gpus entries of the caller's own list
        int kept = 0;
        const char *p = vis;
        while (*p && kept < max_gpus) {
            const char *q = strchr(p, ',');
            const size_t len = q ? (size_t)(q - p) : strlen(p);
            if (kept) list += ",";
            list.append(p, len);
            kept++;
            if (!q) break;
            p = q + 1;
        }
    } else {
        for (int g = 0; g < max_gpus; g++) list += (g ? "," : "") + std::to_string(g);
    }
    return setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1) == 0 ? BS_GPU_OK : BS_GPU_ERR_NOMEM;
}

int bs_gpu_init_ex(bs_gpu_ctx **out, const bs_gpu_config *cfg)
{
    if (!out) return BS_GPU_ERR_INVALID;
    *out = nullptr;
    if (!cfg || cfg->struct_size != sizeof(bs_gpu_config)) return BS_GPU_ERR_INVALID;
    if (cfg->fp_bytes != 4 && cfg->fp_bytes != 8) return BS_GPU_ERR_INVALID;
    if (cfg->num_gpus < 1) return BS_GPU_ERR_INVALID;
    if (cfg->math != BS_MATH_DEFAULT && cfg->math != BS_MATH_IEEE && cfg->math != BS_MATH_FAST && cfg->math != BS_MATH_REFERENCE)
        return BS_GPU_ERR_INVALID;
    if (cfg->unroll != 0 && cfg->unroll != 1 && cfg->unroll != 2 && cfg->unroll != 4) return BS_GPU_ERR_INVALID;
    if (cfg->threads_per_block != 0 && (cfg->threads_per_block < 32 || cfg->threads_per_block > 256 || cfg->threads_per_block % 32))
        return BS_GPU_ERR_INVALID;
    if (cfg->blocks_per_sm < 0 || cfg->blocks_per_sm > 32) return BS_GPU_ERR_INVALID;
    if (cfg->variant < 0 || cfg->variant > 15) return BS_GPU_ERR_INVALID;
    if ((cfg->variant & VARIANT_PIPE) && cfg->unroll == 4) return BS_GPU_ERR_INVALID;  // would spill: not built for use
    if (cfg->num_options > 2147483647ull) return BS_GPU_ERR_INVALID;  // the reference's `int numOptions`

    const bool async_discovery = (cfg->flags & BS_GPU_FLAG_ASYNC_DISCOVERY) != 0;
    if (!async_discovery) {
        const int have = bs_gpu_device_count();
        if (have <= 0) return BS_GPU_ERR_NO_DEVICE;  // no CPU fallback, by design
        if (cfg->num_gpus > have) return BS_GPU_ERR_NO_DEVICE;
        for (int g = 0; g < cfg->num_gpus; g++) {
            const int dev = cfg->devices ? cfg->devices[g] : g;
            if (dev < 0 || dev >= have) return BS_GPU_ERR_NO_DEVICE;
        }
    }

    bs_gpu_ctx *c = new (std::nothrow) bs_gpu_ctx();
    if (!c) return BS_GPU_ERR_NOMEM;
    c->n = cfg->num_options;
    c->fp_bytes = cfg->fp_bytes;
    c->flags = cfg->flags;
    c->math = cfg->math == BS_MATH_DEFAULT ? BS_MATH_FAST : cfg->math;
    c->cfg_threads = cfg->threads_per_block;
    c->cfg_blocks_per_sm = cfg->blocks_per_sm;
    // Defaults from the round-1 sweep on B200 (profiles/r01_tune_sweep.txt): one 16-byte group per thread-trip
    // and, for the MUFU-math fp32 kernel, 4 x 256 resident threads per SM (~98 KB of loads in flight per SM)
    // sit at the top of the curve; more resident warps only add DRAM page conflicts.
    // Sets whose single pass already lasts milliseconds (>= 64M options per device: the 1B-option set) are
    // priced for seconds at a time; the board then sits at its 1000 W power cap, where two groups per trip with
    // all resident CTAs sustain ~3 % more (6.2 TB/s, the same cap the traffic-only probe hits; profiles/
    // r01_sustained_power_cap.txt).
    const bool sustained = c->fp_bytes == 4 && c->math == BS_MATH_FAST && c->n / (size_t)cfg->num_gpus >= ((size_t)64 << 20);
    c->unroll = cfg->unroll ? cfg->unroll : (sustained ? 2 : 1);
    if (!c->cfg_blocks_per_sm && !c->cfg_threads && c->fp_bytes == 4 && c->math == BS_MATH_FAST && !sustained) c->cfg_blocks_per_sm = 4;
    c->variant = cfg->variant;
    // Sets that a launch finishes in a microsecond or two (simsmall: 4,096 options) are bound by the gap between the runs'
    // launches, not by anything the kernel does: programmatic dependent launch lets run j+1 be scheduled while run j
    // drains (measured on B200, tools/simsmall_pdl.py -> profiles/r02_simsmall_pdl.txt: 1.64 -> 1.43 us per run at 4K
    // options, 1.98 -> 1.69 at 64K; from 256K options on it costs time, as it does for the large sets).
    if (c->n / (size_t)cfg->num_gpus <= ((size_t)128 << 10) && !cfg->variant) c->flags |= BS_GPU_FLAG_PDL;
    {
        // Shape of the fp64 TMA kernel: 0 (16 consumer warps, one group per thread and tile, four stages).  While the kernel
        // executed 162 instructions per option, shape 2 (two groups per consumer thread and tile: barrier wait, addresses and
        // loop paid once per four options) was worth 1-3 % in the sustained, power-capped regime every bench line lives in
        // and was the default for shards of a million options and more (profiles/r02_tune_fp64_gtab.txt); at 141 instructions
        // per option the order is the other way round -- 118.6-118.9 (shape 0) / 117.1-118.2 (1) / 116.1-117.4 (2) G options/s,
        // three alternating bench lines on one board, profiles/r02_tune_fp64_shapes_final.txt.
        // BS_GPU_TMA_WIDE = 0 | 1 | 2 forces a shape (measurements, tests).
        const char *w = getenv("BS_GPU_TMA_WIDE");
        if (c->fp_bytes == 8) c->tma_shape = (w && *w >= '0' && *w <= '2') ? *w - '0' : 0;
    }
    // fp64 (unless the caller chose a geometry/variant explicitly): the fast-math kernel takes its inputs through the
    // bulk-copy ring (bs_map_tma: 81.0 us per 10M options = 6.42 TB/s against 85.9 us with software-pipelined LDG.128
    // and 84.3 us for the LDG traffic probe itself, profiles/r02_tune_fp64_tma.txt); ERR_CHK runs, which the TMA kernel
    // does not implement, use the software-pipelined LDG kernel.
    if (c->fp_bytes == 8 && !cfg->variant && !cfg->unroll && !cfg->threads_per_block && !cfg->blocks_per_sm && c->math == BS_MATH_FAST)
        c->variant = VARIANT_PIPE | VARIANT_TMA;

    c->want_gpus = cfg->num_gpus;
    if (cfg->devices) c->want_devices.assign(cfg->devices, cfg->devices + cfg->num_gpus);

    // Host staging (north_star item 1): page-aligned anonymous memory the loader writes SoA straight into.  It is
    // available immediately -- no CUDA context is needed to allocate it -- and the device threads pin it
    // (cudaHostRegister, portable; see do_pin) right before the first copy, so the context creation overlaps the
    // file parse.
    // Buffers of 2 MiB and more are aligned to 2 MiB and advised as transparent huge pages: the loader then takes one
    // page fault per 2 MiB instead of per 4 KiB, and cudaHostRegister pins 512x fewer pages (measured on the B200
    // boxes, tools/micro/h2d_ceiling.cu: 0.05 s instead of 0.27-0.8 s to fault in and pin 2 x 250 MB; the copy rate
    // itself, 55 GB/s per GPU, is the same for every pinned kind).  Where THP is off the advice is a no-op.
    if (!(c->flags & BS_GPU_FLAG_NO_HOST_STAGING)) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t HUGE = (size_t)2 << 20;
        for (int b = 0; b < BS_BUF_COUNT; b++) {
            size_t bytes = (std::max<size_t>(c->n, 1) * elem_bytes(c, b) + page - 1) / page * page;
            const bool huge = bytes >= HUGE;
            if (huge) bytes = (bytes + HUGE - 1) / HUGE * HUGE;
            const size_t span = huge ? bytes + HUGE : bytes;  // room to align the start
            void *m = mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (m == MAP_FAILED) {
                for (int k = 0; k < b; k++) munmap(c->host[k], c->host_bytes[k]);
                delete c;
                return BS_GPU_ERR_NOMEM;
            }
            if (huge) {
                char *base = (char *)m;
                char *aligned = (char *)(((uintptr_t)base + HUGE - 1) & ~(uintptr_t)(HUGE - 1));
                if (aligned > base) munmap(base, (size_t)(aligned - base));
                const size_t tail = (size_t)((base + span) - (aligned + bytes));
                if (tail) munmap(aligned + bytes, tail);
                madvise(aligned, bytes, MADV_HUGEPAGE);
                m = aligned;
            }
            c->host[b] = m;
            c->host_bytes[b] = bytes;
        }
    }

    if (async_discovery) {
        c->bootstrap = std::thread(bootstrap_thread, c);  // cuInit + device bring-up entirely off the caller's thread
    } else {
        start_devices(c, cfg->num_gpus);  // asynchronous from here on: see finish_setup()
        c->setup_pending = true;
    }
    *out = c;
    return BS_GPU_OK;
}

int bs_gpu_init(bs_gpu_ctx **ctx, int num_gpus, size_t num_options, int fp_bytes)
{
    bs_gpu_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.num_options = num_options;
    cfg.fp_bytes = fp_bytes;
    cfg.num_gpus = num_gpus;
    cfg.flags = BS_GPU_FLAG_WITH_DGREFVAL;
    return bs_gpu_init_ex(ctx, &cfg);
}

void *bs_gpu_host_buffer(bs_gpu_ctx *c, int which)
{
    if (!c || which < 0 || which >= BS_BUF_COUNT) return nullptr;
    return c->host[which];
}

int bs_gpu_mark_dirty(bs_gpu_ctx *c)
{
    if (!c) return BS_GPU_ERR_INVALID;
    finish_setup(c);
    c->inputs_dirty = true;
    for (auto &s : c->shards) s.refval_on_device = false;
    return BS_GPU_OK;
}

static int upload_impl(bs_gpu_ctx *c, int what)
{
    if (c->flags & BS_GPU_FLAG_NO_HOST_STAGING) return fail(c, BS_GPU_ERR_STATE, "context has no host staging buffers");
    if ((what & UP_REFVAL) && !(c->flags & BS_GPU_FLAG_WITH_DGREFVAL))
        return fail(c, BS_GPU_ERR_STATE, "context was created without the DGREFVAL stream");
    c->arg_upload_what = what;
    {
        const int pinned = pin_staging(c);
        if (pinned != BS_GPU_OK) return pinned;
    }
    const double t0 = now_ms();
    const int st = broadcast(c, CMD_UPLOAD);
    c->timing.wall_ms = now_ms() - t0;
    if (st != BS_GPU_OK) return st;
    double mx = 0;
    for (auto &s : c->shards) mx = std::max<double>(mx, s.h2d_ms);
    c->timing.h2d_ms = mx;
    c->timing.h2d_bytes = (unsigned long long)c->n * (((what & UP_INPUTS) ? 5ull * c->fp_bytes + 4ull : 0ull) +
                                                      ((what & UP_REFVAL) ? (unsigned long long)c->fp_bytes : 0ull));
    if (what & UP_INPUTS) {
        c->inputs_dirty = false;
        c->device_valid = true;
    }
    return BS_GPU_OK;
}

static bool refval_missing(bs_gpu_ctx *c)
{
    for (auto &s : c->shards)
        if (!s.refval_on_device) return true;
    return false;
}

/* Copies the six input streams H2D.  DGREFVAL is not an input of the Map: it follows lazily, the first
 * time a run asks for err_chk. */
int bs_gpu_upload(bs_gpu_ctx *c)
{
    if (!c) return BS_GPU_ERR_INVALID;
    return upload_impl(c, UP_INPUTS);
}

int bs_gpu_run(bs_gpu_ctx *c, int num_runs, int err_chk, unsigned long long *num_errors)
{
    if (!c || num_runs < 0) return BS_GPU_ERR_INVALID;
    {
        const int ready = finish_setup(c);
        if (ready != BS_GPU_OK) return ready;
    }
    if (!c->device_valid) return fail(c, BS_GPU_ERR_STATE, "no inputs on the device: call bs_gpu_upload / bs_gpu_fill_synthetic first");
    if (err_chk) {
        if (!(c->flags & BS_GPU_FLAG_WITH_DGREFVAL)) return fail(c, BS_GPU_ERR_STATE, "err_chk needs BS_GPU_FLAG_WITH_DGREFVAL");
        if (refval_missing(c)) {
            if (c->flags & BS_GPU_FLAG_NO_HOST_STAGING) return fail(c, BS_GPU_ERR_STATE, "DGREFVAL is not on the device");
            const int up = upload_impl(c, UP_REFVAL);
            if (up != BS_GPU_OK) return up;
        }
    }
    c->arg_num_runs = num_runs;
    c->arg_err_chk = err_chk;
    const double t0 = now_ms();
    const int st = broadcast(c, CMD_RUN);
    c->timing.wall_ms = now_ms() - t0;
    if (st != BS_GPU_OK) return st;
    double mx = 0;
    unsigned long long total = 0, launches = 0;
    for (auto &s : c->shards) {
        mx = std::max<double>(mx, s.roi_ms);
        total += s.err_total;
        if (s.count) launches += (unsigned long long)num_runs;
    }
    c->timing.roi_ms = mx;
    c->timing.kernel_launches = launches;
    if (num_errors) *num_errors = total;
    return BS_GPU_OK;
}

int bs_gpu_download(bs_gpu_ctx *c)
{
    if (!c) return BS_GPU_ERR_INVALID;
    if (c->flags & BS_GPU_FLAG_NO_HOST_STAGING) return fail(c, BS_GPU_ERR_STATE, "context has no host staging buffers");
    {
        const int pinned = pin_staging(c);
        if (pinned != BS_GPU_OK) return pinned;
    }
    const double t0 = now_ms();
    const int st = broadcast(c, CMD_DOWNLOAD);
    c->timing.wall_ms = now_ms() - t0;
    if (st != BS_GPU_OK) return st;
    double mx = 0;
    for (auto &s : c->shards) mx = std::max<double>(mx, s.d2h_ms);
    c->timing.d2h_ms = mx;
    c->timing.d2h_bytes = (unsigned long long)c->n * c->fp_bytes;
    return BS_GPU_OK;
}

int bs_gpu_price(bs_gpu_ctx *c, int num_runs, int err_chk, unsigned long long *num_errors)
{
    if (!c || num_runs < 0) return BS_GPU_ERR_INVALID;
    {
        const int ready = finish_setup(c);
        if (ready != BS_GPU_OK) return ready;
    }
    if (c->flags & BS_GPU_FLAG_NO_HOST_STAGING) return fail(c, BS_GPU_ERR_STATE, "bs_gpu_price needs host staging; use bs_gpu_run");
    if (err_chk && !(c->flags & BS_GPU_FLAG_WITH_DGREFVAL)) return fail(c, BS_GPU_ERR_STATE, "err_chk needs BS_GPU_FLAG_WITH_DGREFVAL");
    const bool need_refval = err_chk && refval_missing(c);
    const int what = (c->inputs_dirty ? UP_INPUTS : 0) | (need_refval ? UP_REFVAL : 0);
    if (!c->inputs_dirty && !c->device_valid) return fail(c, BS_GPU_ERR_STATE, "no inputs on the device");
    c->arg_num_runs = num_runs;
    c->arg_err_chk = err_chk;
    c->arg_upload_what = what;
    const double t0 = now_ms();  // the one-time pinning of the staging buffers is part of the first call's wall time
    {
        const int pinned = pin_staging(c);
        if (pinned != BS_GPU_OK) return pinned;
    }
    const int st = broadcast(c, CMD_PRICE);
    c->timing.wall_ms = now_ms() - t0;
    if (st != BS_GPU_OK) return st;
    unsigned long long total = 0, launches = 0;
    c->timing.h2d_ms = c->timing.roi_ms = c->timing.d2h_ms = c->timing.pipeline_ms = 0;
    for (auto &s : c->shards) {
        c->timing.h2d_ms = std::max<double>(c->timing.h2d_ms, s.h2d_ms);
        c->timing.roi_ms = std::max<double>(c->timing.roi_ms, s.roi_ms);
        c->timing.d2h_ms = std::max<double>(c->timing.d2h_ms, s.d2h_ms);
        c->timing.pipeline_ms = std::max<double>(c->timing.pipeline_ms, s.pipeline_ms);
        total += s.err_total;
        if (s.count) launches += (unsigned long long)num_runs;
    }
    // "launches" counts runs (one logical launch of the Map per run and device), not the chunk launches
    c->timing.kernel_launches = launches;
    c->timing.h2d_bytes = (unsigned long long)c->n * (((what & UP_INPUTS) ? 5ull * c->fp_bytes + 4ull : 0ull) +
                                                      ((what & UP_REFVAL) ? (unsigned long long)c->fp_bytes : 0ull));
    c->timing.d2h_bytes = (unsigned long long)c->n * c->fp_bytes;
    if (what & UP_INPUTS) {
        c->inputs_dirty = false;
        c->device_valid = true;
    }
    if (num_errors) *num_errors = total;
    return BS_GPU_OK;
}

/* The CAF Map's entry (blackscholes.c:482-570): the message is a vector of 24-byte DataCont records. */
int bs_gpu_price_aos(bs_gpu_ctx *c, const void *records, size_t num_records, float *out_prices, int num_runs)
{
    if (!c || num_runs < 0 || (num_records && (!records || !out_prices))) return BS_GPU_ERR_INVALID;
    {
        const int ready = finish_setup(c);
        if (ready != BS_GPU_OK) return ready;
    }
    if (c->fp_bytes != 4) return fail(c, BS_GPU_ERR_STATE, "bs_gpu_price_aos: DataCont holds floats; the context must be created with fp_bytes = 4");
    if (num_records != c->n) return fail(c, BS_GPU_ERR_INVALID, "bs_gpu_price_aos: num_records differs from the context's num_options");
    const double t0 = now_ms();
    c->arg_aos = records;
    c->arg_out = out_prices;
    int st = broadcast(c, CMD_AOS_IN);
    if (st != BS_GPU_OK) return st;
    c->device_valid = true;
    c->inputs_dirty = false;  // the device now holds the message's options, not the staging buffers' (bs_gpu_mark_dirty to go back)
    double h2d = 0;
    for (auto &s : c->shards) h2d = std::max<double>(h2d, s.h2d_ms);
    c->arg_num_runs = num_runs;
    c->arg_err_chk = 0;
    st = broadcast(c, CMD_RUN);
    if (st != BS_GPU_OK) return st;
    st = broadcast(c, CMD_DOWNLOAD_TO);
    if (st != BS_GPU_OK) return st;
    c->timing.wall_ms = now_ms() - t0;
    c->timing.h2d_ms = h2d;
    c->timing.roi_ms = c->timing.d2h_ms = c->timing.pipeline_ms = 0;
    unsigned long long launches = 0;
    for (auto &s : c->shards) {
        c->timing.roi_ms = std::max<double>(c->timing.roi_ms, s.roi_ms);
        c->timing.d2h_ms = std::max<double>(c->timing.d2h_ms, s.d2h_ms);
        if (s.count) launches += (unsigned long long)num_runs;
    }
    c->timing.kernel_launches = launches;
    c->timing.h2d_bytes = (unsigned long long)c->n * 24ull;
    c->timing.d2h_bytes = (unsigned long long)c->n * 4ull;
    return BS_GPU_OK;
}

int bs_gpu_fill_synthetic(bs_gpu_ctx *c, unsigned long long first_index)
{
    if (!c) return BS_GPU_ERR_INVALID;
    c->arg_first_index = first_index;
    const int st = broadcast(c, CMD_FILL);
    if (st != BS_GPU_OK) return st;
    c->device_valid = true;
    c->inputs_dirty = false;
    return BS_GPU_OK;
}

int bs_gpu_read_device(bs_gpu_ctx *c, int which, size_t first, size_t count, void *dst)
{
    if (!c || which < 0 || which >= BS_BUF_COUNT || (!dst && count)) return BS_GPU_ERR_INVALID;
    if (first > c->n || count > c->n - first) return BS_GPU_ERR_INVALID;
    const int ready = finish_setup(c);
    if (ready != BS_GPU_OK) return ready;
    const size_t eb = elem_bytes(c, which);
    for (auto &s : c->shards) {
        const size_t lo = std::max(first, s.first), hi = std::min(first + count, s.first + s.count);
        if (lo >= hi) continue;
        if (!s.d[which]) return fail(c, BS_GPU_ERR_STATE, "stream not allocated on the device");
        cudaError_t e = cudaSetDevice(s.device);
        if (e == cudaSuccess)
            e = cudaMemcpy((char *)dst + (lo - first) * eb, (const char *)s.d[which] + (lo - s.first) * eb, (hi - lo) * eb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            c->err = std::string("bs_gpu_read_device: ") + cudaGetErrorString(e);
            return BS_GPU_ERR_CUDA;
        }
    }
    return BS_GPU_OK;
}

long long bs_gpu_errors(bs_gpu_ctx *c, long long *idx, size_t cap)
{
    if (!c || (!idx && cap)) return BS_GPU_ERR_INVALID;
    finish_setup(c);
    size_t w = 0;
    for (auto &s : c->shards)
        for (long long local : s.list) {
            if (w >= cap) return (long long)w;
            idx[w++] = local + (long long)s.first;
        }
    return (long long)w;
}

int bs_gpu_num_shards(bs_gpu_ctx *c)
{
    if (!c) return BS_GPU_ERR_INVALID;
    const int ready = finish_setup(c);  // with ASYNC_DISCOVERY the shard count is known only after discovery
    return ready != BS_GPU_OK ? ready : (int)c->shards.size();
}

int bs_gpu_shard(bs_gpu_ctx *c, int g, int *device, size_t *first, size_t *count)
{
    if (!c) return BS_GPU_ERR_INVALID;
    finish_setup(c);
    if (g < 0 || g >= (int)c->shards.size()) return BS_GPU_ERR_INVALID;
    if (device) *device = c->shards[g].device;
    if (first) *first = c->shards[g].first;
    if (count) *count = c->shards[g].count;
    return BS_GPU_OK;
}

int bs_gpu_get_timing(bs_gpu_ctx *c, bs_gpu_timing *out)
{
    if (!c || !out) return BS_GPU_ERR_INVALID;
    *out = c->timing;
    return BS_GPU_OK;
}

int bs_gpu_get_launch(bs_gpu_ctx *c, int *math, int *threads_per_block, int *blocks)
{
    if (!c) return BS_GPU_ERR_INVALID;
    const int ready = finish_setup(c);  // the grid is chosen by the device threads during setup
    if (ready != BS_GPU_OK) return ready;
    if (c->shards.empty()) return BS_GPU_ERR_STATE;
    if (math) *math = c->math;
    const bool tma = use_tma(c, false);  // the geometry of the kernel that plain (non-ERR_CHK) runs launch
    if (threads_per_block) *threads_per_block = tma ? tma_threads(c) : c->shards[0].threads;
    if (blocks) *blocks = tma ? c->shards[0].tma_blocks : c->shards[0].blocks;
    return BS_GPU_OK;
}

const char *bs_gpu_last_error(bs_gpu_ctx *c) { return c ? c->err.c_str() : ""; }

void bs_gpu_fini(bs_gpu_ctx *c)
{
    if (!c) return;
    if (c->bootstrap.joinable()) c->bootstrap.join();
    bool any = false;
    for (auto &s : c->shards) any = any || s.worker.joinable();
    if (any) {
        broadcast(c, CMD_TEARDOWN);
        broadcast(c, CMD_EXIT);
        for (auto &s : c->shards)
            if (s.worker.joinable()) s.worker.join();
    }
    for (int b = 0; b < BS_BUF_COUNT; b++)
        if (c->host[b]) munmap(c->host[b], c->host_bytes[b]);
    delete c;
}

}  // extern "C"
