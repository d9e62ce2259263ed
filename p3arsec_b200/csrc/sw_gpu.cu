/*
 * sw_gpu.cu -- implementation of the C ABI in include/sw_gpu.h (libsw_gpu.so): the swaptions Map on B200.
 *
 * Reference citations are relative to /root/reference/parsec-ff/pkgs/apps/swaptions/src/ (HSB: =
 * HJM_Swaption_Blocking.cpp, SEC: = HJM_Securities.cpp).
 *
 * Host side of one sw_gpu_price() call:
 *   1. per swaption, everything HJM_Swaption_Blocking computes before its trial loop (HSB:48-61,116-148: time
 *      indices, the swap payoff vector, the forward curve and the drifts) is evaluated here with the reference's own
 *      expressions and packed into one swk::SwParams record (~1.2 KB) in pinned memory;
 *   2. swaptions are split contiguously over the devices (the static partition of SEC:312 / worker() SEC:100-118);
 *      each device receives its records (H2D), runs the simulation kernel over (swaption, trial-chunk) work items
 *      and the per-swaption finalize kernel, and returns mean / standard error (D2H) -- all asynchronous on the
 *      device's own stream, so the devices work at the same time under one caller thread;
 *   3. the call returns when every device has finished.  Device time is taken with CUDA events on those streams.
 * No inter-device traffic, no collective, no CPU fallback.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/sw_gpu.h"
#include "sw_kernels.cuh"
#include "sw_prepare.h"

namespace {

enum { AUX_STREAMS = 3 };
// A swaption gets a kernel launch of its own (tables in the constant bank: sw_sim_one) from this many simulated trials
// on; below it one kernel walks over all (swaption, chunk) items with the tables in shared memory (sw_sim_fast).
const long long ONE_SWAPTION_MIN_TRIALS = 262144;
// tuning knob: SW_GPU_ONE_MIN_TRIALS=<n> in the environment overrides the threshold (read at every call)
long long one_swaption_min_trials()
{
    const char *e = getenv("SW_GPU_ONE_MIN_TRIALS");
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? v : ONE_SWAPTION_MIN_TRIALS;
}

struct Dev {
    int device = 0;
    int sm_count = 0;
    int first = 0, count = 0;  // shard of the last call
    cudaStream_t stream = nullptr;
    cudaStream_t aux[AUX_STREAMS] = {nullptr};     // one-swaption launches rotate over stream + aux so that their tails overlap
    cudaEvent_t aux_done[AUX_STREAMS] = {nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    swk::SwParams *d_params = nullptr;
    double2 *d_partials = nullptr;
    size_t partial_cap = 0;
    double *d_out = nullptr;  // [mean(max) | err(max)]
    double *d_tables = nullptr;  // [tail table | exp/log tables], expanded once by sw_fill_tables
    int occ_fast = 0, occ_lean = 0, occ_generic = 0;
    float roi_ms = 0;
};

}  // namespace

struct sw_gpu_ctx {
    int max_swaptions = 0, iN = 0, iFactors = 0;
    std::vector<Dev> devs;
    swk::SwParams *h_params = nullptr;  // pinned, max_swaptions records
    double *h_out = nullptr;            // pinned, 2 x max_swaptions
    int cfg_ctas_per_sm = 0, cfg_tpt = 0;
    int shards_used = 0;
    sw_gpu_timing timing;
    std::string err;
    sw_gpu_ctx() { memset(&timing, 0, sizeof(timing)); }
};

namespace {

int fail(sw_gpu_ctx *c, int status, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return status;
}

#define SW_CUDA(c, call)                                                                                      \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail((c), e_ == cudaErrorMemoryAllocation ? SW_GPU_ERR_NOMEM : SW_GPU_ERR_CUDA, "%s: %s", \
                        #call, cudaGetErrorString(e_));                                                       \
    } while (0)

enum Kind { K_FAST = 0, K_LEAN = 1, K_GENERIC = 2 };

}  // namespace

extern "C" {

int sw_gpu_abi_version(void) { return SW_GPU_ABI_VERSION; }

const char *sw_gpu_status_string(int status)
{
    switch (status) {
        case SW_GPU_OK: return "ok";
        case SW_GPU_ERR_INVALID: return "invalid argument";
        case SW_GPU_ERR_NO_DEVICE: return "no CUDA device";
        case SW_GPU_ERR_CUDA: return "CUDA error";
        case SW_GPU_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

int sw_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *sw_gpu_last_error(sw_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void sw_gpu_fini(sw_gpu_ctx *ctx)
{
    if (!ctx) return;
    for (Dev &d : ctx->devs) {
        if (cudaSetDevice(d.device) != cudaSuccess) continue;
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.d_params) cudaFree(d.d_params);
        if (d.d_partials) cudaFree(d.d_partials);
        if (d.d_out) cudaFree(d.d_out);
        if (d.d_tables) cudaFree(d.d_tables);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        for (int k = 0; k < AUX_STREAMS; ++k) {
            if (d.aux_done[k]) cudaEventDestroy(d.aux_done[k]);
            if (d.aux[k]) cudaStreamDestroy(d.aux[k]);
        }
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    delete ctx;
}

static int init_impl(sw_gpu_ctx *c, const int *devices, int num_gpus)
{
    int visible = sw_gpu_device_count();
    if (visible <= 0) return fail(c, SW_GPU_ERR_NO_DEVICE, "no CUDA device visible (libsw_gpu has no CPU fallback)");
    if (!devices && num_gpus > visible) num_gpus = visible;  // like the blackscholes driver: clamp to what the box has
    c->devs.resize(num_gpus);
    const size_t n = (size_t)c->max_swaptions;
    for (int g = 0; g < num_gpus; ++g) {
        Dev &d = c->devs[g];
        d.device = devices ? devices[g] : g;
        if (d.device < 0 || d.device >= visible) return fail(c, SW_GPU_ERR_NO_DEVICE, "device ordinal %d not present (%d visible)", d.device, visible);
        SW_CUDA(c, cudaSetDevice(d.device));
        cudaDeviceProp prop;
        SW_CUDA(c, cudaGetDeviceProperties(&prop, d.device));
        if (prop.major < 10) return fail(c, SW_GPU_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", d.device, prop.major, prop.minor);
        d.sm_count = prop.multiProcessorCount;
        SW_CUDA(c, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        for (int k = 0; k < AUX_STREAMS; ++k) {
            SW_CUDA(c, cudaStreamCreateWithFlags(&d.aux[k], cudaStreamNonBlocking));
            SW_CUDA(c, cudaEventCreateWithFlags(&d.aux_done[k], cudaEventDisableTiming));
        }
        SW_CUDA(c, cudaFuncSetAttribute(swk::sw_sim_one<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)swk::one_shared_bytes(swk::FD)));
        SW_CUDA(c, cudaFuncSetAttribute(swk::sw_sim_one<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)swk::one_shared_bytes(swk::FD)));
        SW_CUDA(c, cudaEventCreate(&d.ev0));
        SW_CUDA(c, cudaEventCreate(&d.ev1));
        SW_CUDA(c, cudaMalloc(&d.d_params, n * sizeof(swk::SwParams)));
        SW_CUDA(c, cudaMalloc(&d.d_out, 2 * n * sizeof(double)));
        SW_CUDA(c, cudaMalloc(&d.d_tables, swk::TABLE_DOUBLES * sizeof(double)));
        swk::sw_fill_tables<<<1, 128, 0, d.stream>>>(d.d_tables);
        SW_CUDA(c, cudaGetLastError());
        SW_CUDA(c, cudaStreamSynchronize(d.stream));
        const size_t full_smem = swk::fast_shared_bytes(swk::FD);
        SW_CUDA(c, cudaFuncSetAttribute(swk::sw_sim_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full_smem));
        SW_CUDA(c, cudaFuncSetAttribute(swk::sw_sim_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full_smem));
        SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_fast, swk::sw_sim_fast<false>, swk::THREADS, full_smem));
        SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_lean, swk::sw_sim_fast<true>, swk::THREADS, full_smem));
        SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_generic, swk::sw_sim_generic, swk::THREADS, 0));
        if (d.occ_fast < 1 || d.occ_lean < 1 || d.occ_generic < 1) return fail(c, SW_GPU_ERR_CUDA, "a kernel does not fit on device %d", d.device);
    }
    SW_CUDA(c, cudaHostAlloc(&c->h_params, n * sizeof(swk::SwParams), cudaHostAllocPortable));
    SW_CUDA(c, cudaHostAlloc(&c->h_out, 2 * n * sizeof(double), cudaHostAllocPortable));
    return SW_GPU_OK;
}

int sw_gpu_init(sw_gpu_ctx **out, int num_gpus, int max_swaptions, int iN, int iFactors)
{
    return sw_gpu_init_devices(out, nullptr, num_gpus, max_swaptions, iN, iFactors);
}

int sw_gpu_init_devices(sw_gpu_ctx **out, const int *devices, int num_gpus, int max_swaptions, int iN, int iFactors)
{
    if (!out) return SW_GPU_ERR_INVALID;
    *out = nullptr;
    if (num_gpus < 1 || max_swaptions < 1 || iN < 2 || iN > SW_GPU_MAX_N || iFactors < 1 || iFactors > SW_GPU_MAX_FACTORS)
        return SW_GPU_ERR_INVALID;
    sw_gpu_ctx *c = new (std::nothrow) sw_gpu_ctx();
    if (!c) return SW_GPU_ERR_NOMEM;
    c->max_swaptions = max_swaptions;
    c->iN = iN;
    c->iFactors = iFactors;
    int st = init_impl(c, devices, num_gpus);
    if (st != SW_GPU_OK) {
        fprintf(stderr, "sw_gpu_init: %s\n", c->err.c_str());
        sw_gpu_fini(c);
        return st;
    }
    *out = c;
    return SW_GPU_OK;
}

int sw_gpu_set_geometry(sw_gpu_ctx *c, int ctas_per_sm, int trials_per_thread)
{
    if (!c || ctas_per_sm < 0 || trials_per_thread < 0) return SW_GPU_ERR_INVALID;
    c->cfg_ctas_per_sm = ctas_per_sm;
    c->cfg_tpt = trials_per_thread;
    return SW_GPU_OK;
}

}  // extern "C"

namespace {

// Every device's streams drained: nothing launched by a failed sw_gpu_price() may still read the pinned parameter
// records or write the pinned results when the call returns (a later call, or sw_gpu_fini's cudaFreeHost, would race).
void quiesce(sw_gpu_ctx *c)
{
    for (auto &d : c->devs) {
        if (cudaSetDevice(d.device) != cudaSuccess) { cudaGetLastError(); continue; }
        if (d.stream) cudaStreamSynchronize(d.stream);
        for (int k = 0; k < AUX_STREAMS; ++k)
            if (d.aux[k]) cudaStreamSynchronize(d.aux[k]);
        cudaGetLastError();
    }
}

int price_impl(sw_gpu_ctx *c, int nSwaptions, const sw_gpu_swaption *swaptions, const double *pdYield,
               const double *ppdFactors, long swaption_seed, long lTrials, int BLOCKSIZE, unsigned flags, double *mean,
               double *std_error);

}  // namespace

extern "C" {

int sw_gpu_price(sw_gpu_ctx *c, int nSwaptions, const sw_gpu_swaption *swaptions, const double *pdYield,
                 const double *ppdFactors, long swaption_seed, long lTrials, int BLOCKSIZE, unsigned flags, double *mean,
                 double *std_error)
{
    if (!c) return SW_GPU_ERR_INVALID;
    const int st = price_impl(c, nSwaptions, swaptions, pdYield, ppdFactors, swaption_seed, lTrials, BLOCKSIZE, flags, mean, std_error);
    if (st != SW_GPU_OK) {  // one cleanup path for every failure, wherever in the per-device loop it happened
        quiesce(c);
        memset(&c->timing, 0, sizeof(c->timing));
    }
    return st;
}

}  // extern "C"

namespace {

int price_impl(sw_gpu_ctx *c, int nSwaptions, const sw_gpu_swaption *swaptions, const double *pdYield,
               const double *ppdFactors, long swaption_seed, long lTrials, int BLOCKSIZE, unsigned flags, double *mean,
               double *std_error)
{
    if (nSwaptions < 0 || nSwaptions > c->max_swaptions || (nSwaptions > 0 && (!swaptions || !pdYield || !ppdFactors || !mean || !std_error)))
        return fail(c, SW_GPU_ERR_INVALID, "sw_gpu_price: bad arguments (nSwaptions = %d, capacity %d)", nSwaptions, c->max_swaptions);
    if (BLOCKSIZE < 1 || lTrials < 0) return fail(c, SW_GPU_ERR_INVALID, "sw_gpu_price: BLOCKSIZE %d, lTrials %ld", BLOCKSIZE, lTrials);
    if (flags & ~(SW_GPU_FLAG_IEEE | SW_GPU_FLAG_LEAN | SW_GPU_FLAG_BATCHED)) return fail(c, SW_GPU_ERR_INVALID, "sw_gpu_price: unknown flags 0x%x", flags);
    const auto w0 = std::chrono::steady_clock::now();
    memset(&c->timing, 0, sizeof(c->timing));
    c->shards_used = 0;
    if (nSwaptions == 0) return SW_GPU_OK;

    const int iN = c->iN, nF = c->iFactors;
    const long long draws = (long long)(iN - 1) * nF;
    const long long sims = lTrials <= 0 ? 0 : ((lTrials + BLOCKSIZE - 1) / BLOCKSIZE) * (long long)BLOCKSIZE;  // HSB:156
    // the fast kernels: the reference drivers' shape, counters inside the range of the 32-bit residue arithmetic
    Kind kind = K_GENERIC;
    if (!(flags & SW_GPU_FLAG_IEEE) && iN == swk::FN && nF == swk::FF && swaption_seed >= 0 &&
        (double)swaption_seed + (double)nSwaptions + (double)sims * (double)draws < 1099511627776.0 /* 2^40 */)
        kind = (flags & SW_GPU_FLAG_LEAN) ? K_LEAN : K_FAST;

    const int G = std::min<int>((int)c->devs.size(), nSwaptions);
    const int q = nSwaptions / G, r = nSwaptions % G;
    int first = 0;
    for (int g = 0; g < G; ++g) {
        Dev &d = c->devs[g];
        d.first = first;
        d.count = q + (g < r ? 1 : 0);
        first += d.count;

        // this shard's parameter records are prepared while the devices before it already simulate
        for (int i = d.first; i < d.first + d.count; ++i) {
            if (!swk::prepare(c->h_params[i], swaptions[i], iN, nF, pdYield + (size_t)i * iN, ppdFactors + (size_t)i * nF * (iN - 1),
                         swaption_seed + i, lTrials, BLOCKSIZE)) {
                return fail(c, SW_GPU_ERR_INVALID, "swaption %d: dYears/dMaturity/dTenor/dPaymentInterval put a time index outside the %d-point HJM path", i, iN);
            }
        }

        // shared memory of the fast kernels: the lean one keeps only the draws of the time steps it simulates
        // (full kernel only: a lean launch of one swaption is too short -- 20 us per million trials -- and measured 33 % slower)
        const bool one = kind == K_FAST && !(flags & SW_GPU_FLAG_BATCHED) && sims >= one_swaption_min_trials();
        int z_rows = swk::FD;
        if (kind == K_LEAN) {
            int max_start = 1;
            for (int i = d.first; i < d.first + d.count; ++i) max_start = std::max(max_start, c->h_params[i].start);
            z_rows = swk::FF * max_start;
        }
        const size_t smem = kind == K_GENERIC ? 0 : one ? swk::one_shared_bytes(z_rows) : swk::fast_shared_bytes(z_rows);
        SW_CUDA(c, cudaSetDevice(d.device));
        int occ = d.occ_generic;
        if (kind == K_FAST) {
            if (one) SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, swk::sw_sim_one<false>, swk::THREADS, smem));
            else occ = d.occ_fast;
        } else if (kind == K_LEAN) {
            if (one) SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, swk::sw_sim_one<true>, swk::THREADS, smem));
            else SW_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, swk::sw_sim_fast<true>, swk::THREADS, smem));
        }
        if (occ < 1) return fail(c, SW_GPU_ERR_CUDA, "kernel does not fit on device %d", d.device);
        const int per_sm = c->cfg_ctas_per_sm > 0 ? std::min(c->cfg_ctas_per_sm, occ) : occ;
        const long long grid_ctas = (long long)d.sm_count * per_sm;
        const long long grid_threads = grid_ctas * swk::THREADS;
        // trials per thread per item.  Batched kernel: enough items for ~8 per resident CTA (static round-robin balance),
        // capped.  One swaption per launch: one item per CTA, the swaption's trials spread over the whole grid.
        int tpt = c->cfg_tpt;
        if (tpt <= 0) {
            const long long want = one ? (sims + grid_threads - 1) / grid_threads : (sims * d.count) / (grid_threads * 8);
            tpt = (int)std::max<long long>(1, std::min<long long>(one ? 4096 : 64, want));
        }
        swk::Geom geo;
        geo.iN = iN;
        geo.iFactors = nF;
        geo.tpt = tpt;
        geo.chunk_trials = (long long)swk::THREADS * tpt;
        const long long chunks = std::max<long long>(1, (sims + geo.chunk_trials - 1) / geo.chunk_trials);
        if (chunks * d.count > 0x7fffffffLL) return fail(c, SW_GPU_ERR_INVALID, "too many work items");
        geo.chunks = (int)chunks;
        geo.items = (int)(chunks * d.count);
        const int blocks = (int)std::min<long long>(one ? chunks : (long long)geo.items, grid_ctas);

        if ((size_t)geo.items > d.partial_cap) {
            if (d.d_partials) SW_CUDA(c, cudaFree(d.d_partials));
            d.d_partials = nullptr;
            d.partial_cap = 0;
            SW_CUDA(c, cudaMalloc(&d.d_partials, (size_t)geo.items * sizeof(double2)));
            d.partial_cap = (size_t)geo.items;
        }
        SW_CUDA(c, cudaMemcpyAsync(d.d_params, c->h_params + d.first, (size_t)d.count * sizeof(swk::SwParams), cudaMemcpyHostToDevice, d.stream));
        SW_CUDA(c, cudaEventRecord(d.ev0, d.stream));
        if (one) {
            for (int k = 0; k < AUX_STREAMS; ++k) SW_CUDA(c, cudaStreamWaitEvent(d.aux[k], d.ev0, 0));
            for (int i = 0; i < d.count; ++i) {
                const swk::SwParams &H = c->h_params[d.first + i];
                swk::OneSwaption P;
                swk::to_one_swaption(P, H);
                P.chunk_trials = geo.chunk_trials;
                P.tpt = tpt;
                P.chunks = geo.chunks;
                P.partial_base = i * geo.chunks;
                P.sw_index = i;
                cudaStream_t st = (i % (AUX_STREAMS + 1)) == 0 ? d.stream : d.aux[(i % (AUX_STREAMS + 1)) - 1];
                if (kind == K_FAST) swk::sw_sim_one<false><<<blocks, swk::THREADS, smem, st>>>(P, d.d_params, d.d_partials, d.d_tables);
                else swk::sw_sim_one<true><<<blocks, swk::THREADS, smem, st>>>(P, d.d_params, d.d_partials, d.d_tables);
            }
            SW_CUDA(c, cudaGetLastError());
            for (int k = 0; k < AUX_STREAMS; ++k) {
                SW_CUDA(c, cudaEventRecord(d.aux_done[k], d.aux[k]));
                SW_CUDA(c, cudaStreamWaitEvent(d.stream, d.aux_done[k], 0));
            }
            c->timing.kernel_launches += (unsigned long long)d.count - 1;
        } else if (kind == K_FAST)
            swk::sw_sim_fast<false><<<blocks, swk::THREADS, smem, d.stream>>>(d.d_params, geo, d.d_partials, d.d_tables);
        else if (kind == K_LEAN)
            swk::sw_sim_fast<true><<<blocks, swk::THREADS, smem, d.stream>>>(d.d_params, geo, d.d_partials, d.d_tables);
        else
            swk::sw_sim_generic<<<blocks, swk::THREADS, 0, d.stream>>>(d.d_params, geo, d.d_partials);
        SW_CUDA(c, cudaGetLastError());
        swk::sw_finalize<<<d.count, swk::THREADS, 0, d.stream>>>(d.d_params, geo, d.d_partials, d.d_out, d.d_out + c->max_swaptions);
        SW_CUDA(c, cudaGetLastError());
        SW_CUDA(c, cudaEventRecord(d.ev1, d.stream));
        SW_CUDA(c, cudaMemcpyAsync(c->h_out + d.first, d.d_out, (size_t)d.count * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        SW_CUDA(c, cudaMemcpyAsync(c->h_out + c->max_swaptions + d.first, d.d_out + c->max_swaptions, (size_t)d.count * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        c->timing.kernel_launches += 2;
        c->timing.h2d_bytes += (unsigned long long)d.count * sizeof(swk::SwParams);
        c->timing.d2h_bytes += (unsigned long long)d.count * 2 * sizeof(double);
    }
    c->shards_used = G;
    double roi = 0;
    for (int g = 0; g < G; ++g) {
        Dev &d = c->devs[g];
        SW_CUDA(c, cudaSetDevice(d.device));
        SW_CUDA(c, cudaStreamSynchronize(d.stream));
        SW_CUDA(c, cudaEventElapsedTime(&d.roi_ms, d.ev0, d.ev1));
        roi = std::max(roi, (double)d.roi_ms);
    }
    memcpy(mean, c->h_out, (size_t)nSwaptions * sizeof(double));
    memcpy(std_error, c->h_out + c->max_swaptions, (size_t)nSwaptions * sizeof(double));
    c->timing.roi_ms = roi;
    c->timing.trials_simulated = (unsigned long long)sims * (unsigned long long)nSwaptions;
    c->timing.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return SW_GPU_OK;
}

}  // namespace

extern "C" {

int sw_gpu_get_timing(sw_gpu_ctx *c, sw_gpu_timing *out)
{
    if (!c || !out) return SW_GPU_ERR_INVALID;
    *out = c->timing;
    return SW_GPU_OK;
}

int sw_gpu_num_shards(sw_gpu_ctx *c) { return c ? c->shards_used : SW_GPU_ERR_INVALID; }

int sw_gpu_shard(sw_gpu_ctx *c, int g, int *device, int *first, int *count)
{
    if (!c || g < 0 || g >= c->shards_used) return SW_GPU_ERR_INVALID;
    if (device) *device = c->devs[g].device;
    if (first) *first = c->devs[g].first;
    if (count) *count = c->devs[g].count;
    return SW_GPU_OK;
}

}  // extern "C"
