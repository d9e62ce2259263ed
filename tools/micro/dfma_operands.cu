// dfma_operands.cu -- does the DFMA rate depend on where its operands come from?  16 warps/SM, 8 independent chains.
//   (a) x = fma(x, a, b)      a, b kernel parameters (constant bank / uniform registers)
//   (b) x = fma(x, y, z)      three distinct register pairs
//   (c) x = fma(y, z, x)      accumulator last
//   (d) x = fma(x, y, c[k])   Horner step: one register, one __constant__ coefficient
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double COEF[8] = {0.999, 0.998, 0.997, 0.996, 0.995, 0.994, 0.993, 0.992};

template <int V>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = threadIdx.x * 1e-9 + i;
        y[i] = 0.999999 + threadIdx.x * 1e-12 + i * 1e-9;
        z[i] = 1e-7 + i * 1e-9 + threadIdx.x * 1e-13;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (V == 0) x[i] = fma(x[i], a, b);
            if (V == 1) x[i] = fma(x[i], y[i], z[i]);
            if (V == 2) x[i] = fma(y[i], z[i], x[i]);
            if (V == 3) x[i] = fma(x[i], y[i], COEF[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char *name, int sms, double *d)
{
    const int threads = 128, iters = 20000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<V><<<blocks, threads>>>(d, 100, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %.2f scheduler cycles per DFMA warp-instruction, %.2f T DFMA/s\n", name, ms * 1e-3 * 1.965e9 / iters / 4.0 / 8.0,
           (double)blocks * threads * iters * 8 / ms / 1e9);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 128 * sizeof(double));
    printf("%s\n", p.name);
    run<0>("(a) fma(x, param, param)", p.multiProcessorCount, d);
    run<1>("(b) fma(x, y, z): three register pairs", p.multiProcessorCount, d);
    run<2>("(c) fma(y, z, x): accumulate", p.multiProcessorCount, d);
    run<3>("(d) fma(x, y, __constant__)", p.multiProcessorCount, d);
    return 0;
}
