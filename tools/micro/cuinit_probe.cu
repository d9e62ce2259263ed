// cuinit_probe.cu -- where does the start-up time of a one-shot pricing process go?
//   cuinit_probe [ndev]     prints one JSON line: driver initialisation (first runtime call), then the primary context
//                           of devices 0..ndev-1 created one after the other, then the same in parallel threads.
// Run it under different environments (CUDA_VISIBLE_DEVICES=0, CUDA_MODULE_LOADING=LAZY|EAGER, CUDA_DEVICE_ORDER=...)
// to see what a drop-in driver can do about the start-up cost on a multi-GPU box (VERDICT r1 "Next" 6).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cuinit_probe cuinit_probe.cu -lpthread
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    const double t0 = now_s();
    int have = 0;
    const cudaError_t e = cudaGetDeviceCount(&have);  // cuInit happens here
    const double t_init = now_s() - t0;
    if (e != cudaSuccess || have < 1) {
        printf("{\"error\": \"%s\", \"cuinit_s\": %.3f}\n", cudaGetErrorString(e), t_init);
        return 1;
    }
    int want = argc > 1 ? atoi(argv[1]) : have;
    if (want > have) want = have;
    const bool parallel = argc > 2 && atoi(argv[2]) != 0;
    std::vector<double> ctx(want, 0.0);
    const double t1 = now_s();
    if (!parallel) {
        for (int g = 0; g < want; g++) {
            const double a = now_s();
            cudaSetDevice(g);
            cudaFree(0);
            ctx[g] = now_s() - a;
        }
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < want; g++)
            th.emplace_back([&, g] {
                const double a = now_s();
                cudaSetDevice(g);
                cudaFree(0);
                ctx[g] = now_s() - a;
            });
        for (auto &t : th) t.join();
    }
    const double t_ctx = now_s() - t1;
    const char *vis = getenv("CUDA_VISIBLE_DEVICES");
    const char *ml = getenv("CUDA_MODULE_LOADING");
    printf("{\"visible_devices\": %d, \"CUDA_VISIBLE_DEVICES\": \"%s\", \"CUDA_MODULE_LOADING\": \"%s\", \"cuinit_s\": %.3f, \"contexts\": %d, "
           "\"parallel\": %s, \"contexts_total_s\": %.3f, \"per_context_s\": [",
           have, vis ? vis : "", ml ? ml : "", t_init, want, parallel ? "true" : "false", t_ctx);
    for (int g = 0; g < want; g++) printf("%s%.3f", g ? ", " : "", ctx[g]);
    printf("], \"total_s\": %.3f}\n", now_s() - t0);
    return 0;
}
