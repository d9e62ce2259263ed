// h2d_ceiling.cu -- what can this box copy host->device (and back) when G GPUs copy at the same time?
//
// bs_gpu_price() moves, per step and GPU, six input streams H2D (240 MB for the native set) and one price stream D2H
// (40 MB).  Round 1 saw the per-GPU H2D rate of the 8-rank bench fall from ~55 to ~20 GB/s; this tool measures the
// ceiling of the box itself so that the staging scheme can be judged against it ("Next" 2):
//
//     h2d_ceiling [--mb 240] [--reps 8] [--gpus 1,2,4,8]
//
// For every staging kind x GPU count it starts one host thread per GPU (each bound to its device, own stream, own
// staging buffers, first-touched by that thread), releases them together, and has every thread copy its six buffers
// `reps` times back to back; the rate is bytes / (last thread's finish - common start), on the host clock around
// stream synchronisation, plus the per-GPU device-event rate.  Staging kinds:
//     hostalloc      cudaHostAlloc(portable)                      -- the runtime's own pinned allocation
//     hostalloc_wc   cudaHostAlloc(portable | write-combined)
//     register       anonymous mmap (4 KiB pages) + cudaHostRegister        -- what round 1 shipped
//     register_thp   2 MiB-aligned anonymous mmap + MADV_HUGEPAGE + cudaHostRegister
//     register_htlb  MAP_HUGETLB mmap + cudaHostRegister (only if the box has huge pages reserved)
//     pageable       plain malloc (the driver stages through its own bounce buffers)
// Output: one JSON object per line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o h2d_ceiling h2d_ceiling.cu -lpthread
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

enum Kind { HOSTALLOC, HOSTALLOC_WC, REGISTER, REGISTER_THP, REGISTER_HTLB, PAGEABLE, KINDS };
const char *kind_name[KINDS] = {"hostalloc", "hostalloc_wc", "register", "register_thp", "register_htlb", "pageable"};

struct Buf {
    void *p = nullptr, *base = nullptr;
    size_t bytes = 0, mapped = 0;
    bool registered = false, cuda_alloc = false, mmaped = false;
};

bool alloc_buf(Buf &b, size_t bytes, Kind kind)
{
    b.bytes = bytes;
    const size_t HUGE = (size_t)2 << 20;
    switch (kind) {
    case HOSTALLOC:
    case HOSTALLOC_WC:
        if (cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable | (kind == HOSTALLOC_WC ? cudaHostAllocWriteCombined : 0)) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        b.cuda_alloc = true;
        break;
    case REGISTER:
    case REGISTER_THP:
    case REGISTER_HTLB: {
        b.mapped = (bytes + HUGE - 1) / HUGE * HUGE + (kind == REGISTER_THP ? HUGE : 0);
        const int flags = MAP_PRIVATE | MAP_ANONYMOUS | (kind == REGISTER_HTLB ? MAP_HUGETLB : 0);
        void *m = mmap(nullptr, b.mapped, PROT_READ | PROT_WRITE, flags, -1, 0);
        if (m == MAP_FAILED) return false;
        b.mmaped = true;
        b.base = m;
        char *base = (char *)m;
        if (kind == REGISTER_THP) {
            char *aligned = (char *)(((uintptr_t)base + HUGE - 1) & ~(uintptr_t)(HUGE - 1));
            madvise(aligned, (bytes + HUGE - 1) / HUGE * HUGE, MADV_HUGEPAGE);
            b.p = aligned;
        } else {
            b.p = base;
        }
        break;
    }
    case PAGEABLE:
        b.p = malloc(bytes);
        if (!b.p) return false;
        break;
    default: return false;
    }
    memset(b.p, 1, bytes);  // first touch by the owning thread; faults every page in
    if (kind == REGISTER || kind == REGISTER_THP || kind == REGISTER_HTLB) {
        if (cudaHostRegister(b.p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        b.registered = true;
    }
    return true;
}

void free_buf(Buf &b)
{
    if (b.registered) cudaHostUnregister(b.p);
    if (b.cuda_alloc) cudaFreeHost(b.p);
    else if (b.mmaped) munmap(b.base, b.mapped);
    else free(b.p);
    b = Buf();
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Result {
    bool ok = true;
    double t_end = 0;   // host clock when this GPU's copies were done
    float dev_ms = 0;   // device events around this GPU's copies
    double alloc_s = 0; // allocation + first touch + pinning of the staging buffers
};

long anon_huge_kb()
{
    FILE *f = fopen("/proc/meminfo", "r");
    if (!f) return -1;
    char line[256];
    long v = -1;
    while (fgets(line, sizeof line, f))
        if (sscanf(line, "AnonHugePages: %ld kB", &v) == 1) break;
    fclose(f);
    return v;
}

// dir: 0 = H2D only, 1 = D2H only, 2 = both at once (two streams: the two copy engines)
void measure(Kind kind, int G, size_t total_bytes, int reps, int dir)
{
    const int NB = 6;
    const size_t per = (total_bytes / NB + 4095) & ~(size_t)4095;
    std::vector<Result> res(G);
    std::atomic<int> ready(0), go(0);
    double t_start = 0;
    std::vector<std::thread> th;
    const long huge_before = anon_huge_kb();
    std::atomic<long> huge_peak(huge_before);
    for (int g = 0; g < G; g++) {
        th.emplace_back([&, g] {
            Result &r = res[g];
            if (cudaSetDevice(g) != cudaSuccess) { r.ok = false; ready++; return; }
            cudaStream_t s_in, s_out;
            cudaEvent_t e0, e1, e2;
            cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
            cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
            Buf hb[NB];
            char *d = nullptr;
            const double a0 = now_s();
            for (int b = 0; b < NB && r.ok; b++) r.ok = alloc_buf(hb[b], per, kind);
            r.alloc_s = now_s() - a0;
            if (r.ok && cudaMalloc((void **)&d, per * NB) != cudaSuccess) r.ok = false;
            if (r.ok) {  // warm-up copy
                cudaMemcpyAsync(d, hb[0].p, per, cudaMemcpyHostToDevice, s_in);
                cudaStreamSynchronize(s_in);
            }
            long h = anon_huge_kb();
            if (h > huge_peak) huge_peak = h;
            ready++;
            while (!go.load()) std::this_thread::yield();
            if (r.ok) {
                cudaEventRecord(e0, s_in);
                cudaStreamWaitEvent(s_out, e0, 0);
                for (int it = 0; it < reps; it++)
                    for (int b = 0; b < NB; b++) {
                        if (dir == 0 || dir == 2) cudaMemcpyAsync(d + (size_t)b * per, hb[b].p, per, cudaMemcpyHostToDevice, s_in);
                        if (dir == 1) cudaMemcpyAsync(hb[b].p, d + (size_t)b * per, per, cudaMemcpyDeviceToHost, s_out);
                        if (dir == 2 && b == 0) cudaMemcpyAsync(hb[NB - 1].p, d + (size_t)(NB - 1) * per, per, cudaMemcpyDeviceToHost, s_out);
                    }
                cudaEventRecord(e2, s_out);
                cudaStreamWaitEvent(s_in, e2, 0);
                cudaEventRecord(e1, s_in);
                if (cudaStreamSynchronize(s_in) != cudaSuccess) r.ok = false;
                cudaStreamSynchronize(s_out);
                r.t_end = now_s();
                cudaEventElapsedTime(&r.dev_ms, e0, e1);
            }
            for (int b = 0; b < NB; b++)
                if (hb[b].p) free_buf(hb[b]);
            if (d) cudaFree(d);
            cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
            cudaStreamDestroy(s_in); cudaStreamDestroy(s_out);
        });
    }
    while (ready.load() < G) std::this_thread::yield();
    t_start = now_s();
    go = 1;
    for (auto &t : th) t.join();
    bool ok = true;
    double t_end = 0, alloc = 0;
    for (auto &r : res) { ok = ok && r.ok; t_end = r.t_end > t_end ? r.t_end : t_end; alloc = r.alloc_s > alloc ? r.alloc_s : alloc; }
    // bytes moved per GPU in the measured direction(s)
    double h2d_bytes = (dir == 0 || dir == 2) ? (double)per * NB * reps : 0;
    double d2h_bytes = dir == 1 ? (double)per * NB * reps : dir == 2 ? (double)per * reps : 0;
    printf("{\"staging\": \"%s\", \"gpus\": %d, \"direction\": \"%s\", \"ok\": %s, \"mb_per_copy_set\": %.1f, \"reps\": %d", kind_name[kind], G,
           dir == 0 ? "h2d" : dir == 1 ? "d2h" : "h2d+d2h", ok ? "true" : "false", per * NB / 1e6, reps);
    if (ok) {
        const double wall = t_end - t_start;
        printf(", \"aggregate_gbs\": %.2f, \"per_gpu_gbs_wall\": %.2f, \"per_gpu_gbs_device\": [", (h2d_bytes + d2h_bytes) * G / wall / 1e9,
               (h2d_bytes + d2h_bytes) / wall / 1e9);
        for (int g = 0; g < G; g++) printf("%s%.2f", g ? ", " : "", (h2d_bytes + d2h_bytes) / (res[g].dev_ms * 1e-3) / 1e9);
        printf("], \"alloc_pin_s\": %.3f, \"anon_huge_mb_gained\": %.0f", alloc, (huge_peak.load() - huge_before) / 1024.0);
    }
    printf("}\n");
    fflush(stdout);
}

}  // namespace

int main(int argc, char **argv)
{
    size_t mb = 240;
    int reps = 8;
    std::vector<int> gpus;
    std::vector<int> kinds;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--mb") && i + 1 < argc) mb = (size_t)atol(argv[++i]);
        else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) {
            for (char *tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) gpus.push_back(atoi(tok));
        } else if (!strcmp(argv[i], "--kinds") && i + 1 < argc) {
            for (char *tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ","))
                for (int k = 0; k < KINDS; k++)
                    if (!strcmp(tok, kind_name[k])) kinds.push_back(k);
        } else {
            fprintf(stderr, "usage: h2d_ceiling [--mb 240] [--reps 8] [--gpus 1,2,4,8] [--kinds hostalloc,register,...]\n");
            return 2;
        }
    }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) {
        fprintf(stderr, "h2d_ceiling: no CUDA device\n");
        return 1;
    }
    if (gpus.empty())
        for (int g = 1; g <= have; g *= 2) gpus.push_back(g);
    if (kinds.empty())
        for (int k = 0; k < KINDS; k++) kinds.push_back(k);
    for (int g = 0; g < have; g++) {  // contexts up before anything is timed
        cudaSetDevice(g);
        cudaFree(0);
    }
    {
        FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
        char line[128] = "?";
        if (f) { if (!fgets(line, sizeof line, f)) strcpy(line, "?"); fclose(f); }
        for (char *p = line; *p; p++) if (*p == '\n') *p = 0;
        printf("{\"info\": \"box\", \"cuda_devices\": %d, \"host_threads\": %u, \"thp_enabled\": \"%s\"}\n", have, std::thread::hardware_concurrency(), line);
    }
    for (int G : gpus) {
        if (G > have) continue;
        for (int k : kinds) {
            measure((Kind)k, G, mb << 20, reps, 0);
            if (k == HOSTALLOC || k == REGISTER || k == REGISTER_THP) {
                measure((Kind)k, G, mb << 20, reps, 1);
                measure((Kind)k, G, mb << 20, reps, 2);
            }
        }
    }
    return 0;
}
