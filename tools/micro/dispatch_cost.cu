// dispatch_cost.cu -- what one non-FP64 instruction costs next to a stream of DFMAs on a B200 SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dispatch_cost dispatch_cost.cu && ./dispatch_cost
// Every thread runs 8 independent DFMA chains per trip plus K instructions of one kind (inline PTX, volatile, results
// unused); 16 warps per SM (4 per scheduler), as in the swaptions kernel.  Reported: cycles per warp-trip on one
// scheduler and the marginal cost of the added instruction in scheduler cycles.
#include <cstdio>
#include <cuda_runtime.h>

enum Op { NONE, I2F64, RCP64H, IMADWIDE, LOP, SHF, LDS64, STS64, FSEL32, IMAD32, DADDX, F2F };

// One instruction of kind OP whose result is kept alive by xor-ing one of its words into the integer accumulator `a`
// (one LOP3, whose own cost the LOP row measures; LOP itself uses a = a ^ b ^ rot only).
template <int OP>
__device__ __forceinline__ void one(unsigned &a, unsigned b, double &d, double *sm)
{
    unsigned lo = 0, hi = 0;
    if (OP == I2F64) { double r; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(r) : "r"(a)); asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "d"(r)); a ^= hi; }
    if (OP == RCP64H) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(d)); }  // chain on d (four independent chains)
    if (OP == IMADWIDE) { unsigned long long r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r)); a ^= hi; }
    if (OP == LOP) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a) : "r"(a), "r"(b), "r"(a >> 0)); }
    if (OP == SHF) { unsigned r; asm volatile("shf.r.wrap.b32 %0, %1, %2, 7;" : "=r"(r) : "r"(a), "r"(b)); a ^= r; }
    if (OP == LDS64) { double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + a) & 127)))); asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "d"(r)); a ^= hi; }
    if (OP == STS64) { asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(sm + threadIdx.x)), "d"(d)); }
    if (OP == FSEL32) { unsigned r; asm volatile("{ .reg .pred p; setp.gt.u32 p, %1, %2; selp.b32 %0, %1, %2, p; }" : "=r"(r) : "r"(a), "r"(b)); a ^= r; }
    if (OP == IMAD32) { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(a)); a ^= r; }
    if (OP == DADDX) { asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(d), "d"(d)); }
    if (OP == F2F) { float r; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(d)); a ^= __float_as_uint(r); }
}

template <int OP, int K>
__global__ void k(double *out, int iters, double a, double b, unsigned ia)
{
    __shared__ double sm[256];
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    unsigned u[4] = {threadIdx.x + ia, threadIdx.x * 3 + ia, threadIdx.x * 5 + ia, threadIdx.x * 7 + ia};
    double y[4] = {1.5 + threadIdx.x, 2.5 + threadIdx.x, 3.5 + threadIdx.x, 4.5 + threadIdx.x};
    sm[threadIdx.x] = 1.0;
    if (threadIdx.x < 128) sm[threadIdx.x + 128] = 2.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = fma(x[i], a, b);
            if (i < K) one<OP>(u[i & 3], ia, y[i & 3], sm);
        }
        if (K > 8) {
#pragma unroll
            for (int i = 8; i < K; ++i) one<OP>(u[i & 3], ia, y[i & 3], sm);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (u[0] ^ u[1] ^ u[2] ^ u[3]) + y[0] + y[1] + y[2] + y[3];
}

template <int OP, int K>
double run(int sms, double *d)
{
    const int threads = 128, iters = 20000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP, K><<<blocks, threads>>>(d, 100, 0.999999, 1e-7, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP, K><<<blocks, threads>>>(d, iters, 0.999999, 1e-7, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3 * 1.965e9 / iters / 4.0;  // scheduler cycles per warp-trip (4 warps per scheduler)
}

template <int OP>
void report(const char *name, int sms, double *d, double base)
{
    const double t4 = run<OP, 4>(sms, d), t8 = run<OP, 8>(sms, d), t16 = run<OP, 16>(sms, d);
    printf("%-22s 8 DFMA + 4: %6.2f  + 8: %6.2f  + 16: %6.2f cycles/warp-trip   marginal %5.2f / %5.2f / %5.2f cycles per added instruction\n", name, t4,
           t8, t16, (t4 - base) / 4, (t8 - base) / 8, (t16 - base) / 16);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 128 * sizeof(double));
    const int sms = p.multiProcessorCount;
    const double base = run<NONE, 0>(sms, d);
    printf("%s: 8 DFMA alone: %.2f scheduler cycles per warp-trip (%.2f per DFMA), 16 warps/SM\n", p.name, base, base / 8);
    report<DADDX>("DADD (FP64 pipe)", sms, d, base);
    report<I2F64>("I2F.F64.S32 + LOP3", sms, d, base);
    report<RCP64H>("MUFU.RCP64H", sms, d, base);
    report<IMADWIDE>("IMAD.WIDE.U32 + LOP3", sms, d, base);
    report<IMAD32>("IMAD (32-bit) + LOP3", sms, d, base);
    report<LOP>("LOP3", sms, d, base);
    report<SHF>("SHF + LOP3", sms, d, base);
    report<FSEL32>("ISETP + SEL + LOP3", sms, d, base);
    report<LDS64>("LDS.64 + 2 ALU", sms, d, base);
    report<STS64>("STS.64", sms, d, base);
    return 0;
}
