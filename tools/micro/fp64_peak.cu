// fp64_peak.cu -- measured FP64 (DFMA) issue rate of a B200, the denominator of the swaptions roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
// Each thread runs CHAINS independent DFMA chains for ITERS trips; variants: pure DFMA, DFMA + the same number of
// integer instructions (does an FP64 instruction take one issue slot or two?), and warps per SM from 4 to 32.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int INTS>
__global__ void k(double *out, int iters, double a, double b, unsigned ia)
{
    double x[CHAINS];
    unsigned y[INTS > 0 ? INTS : 1];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < (INTS > 0 ? INTS : 1); ++i) y[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
        for (int i = 0; i < INTS; ++i) y[i] = (y[i] ^ ia) + (y[i] >> 3);  // 2-3 ALU instructions each
    }
    double s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < (INTS > 0 ? INTS : 1); ++i) t += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int CHAINS, int INTS>
void run(const char *name, int sms, int warps_per_sm, double *d)
{
    const int threads = 128, iters = 20000;
    const int blocks = sms * warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<CHAINS, INTS><<<blocks, threads>>>(d, 100, 0.999999, 1e-7, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<CHAINS, INTS><<<blocks, threads>>>(d, iters, 0.999999, 1e-7, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)blocks * threads * iters * CHAINS;
    printf("%-28s warps/SM %2d chains %d : %.3f ms  %.2f T DFMA/s = %.1f TFLOP/s  (%.2f DFMA thread-instr/clk/SM at 1965 MHz)\n", name,
           warps_per_sm, CHAINS, ms, dfma / ms / 1e9, 2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / sms / 1.965e9);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 64 * 32 * sizeof(double));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32, 64}) run<8, 0>("pure DFMA", p.multiProcessorCount, w, d);
    for (int w : {16, 32}) run<4, 0>("pure DFMA", p.multiProcessorCount, w, d);
    for (int w : {16, 32}) run<2, 0>("pure DFMA", p.multiProcessorCount, w, d);
    for (int w : {16, 32}) run<1, 0>("pure DFMA (latency)", p.multiProcessorCount, w, d);
    for (int w : {16, 32}) run<8, 4>("DFMA + ~1.25x int ops", p.multiProcessorCount, w, d);
    for (int w : {16, 32}) run<8, 8>("DFMA + ~2.5x int ops", p.multiProcessorCount, w, d);
    return 0;
}
